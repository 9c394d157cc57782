"""CPU oracle (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py) for the replay-buffer construction of SURVEY 8f N3:
a numpy restatement of `preprocess_train_data_incl_act_rew` (recommenders/data_utils/preprocessing.py:199-320) with its
helpers `get_state` (:5-29) and `get_next_state` (:143-170) -- the sliding-window part that the IKEA preprocessing
(ikea/data_utils/preprocessing.py:385-489) shares through `get_state_col` / `get_next_state_col`.

Input: one event log sorted by session (session ids, item ids[, rewards]).  Output columns, one row per event:
state / next_state [n, L] padded with `pad_id` at the end ("end") or the beginning ("beg"), action, true_state_len =
clip(n_items_bef, 1, L), true_next_state_len = min(n_items_bef + 1, L), is_end (last event of its session).

Pinned: `python -m oracle.make_golden_preprocess` runs the unmodified reference function on a seeded log and asserts
equality with this restatement; the log and the reference's output are committed as tests/golden/preprocess_rr.npz."""

import numpy as np


def session_offsets(session_ids):
    """CSR offsets [S+1] of a log whose events of one session are contiguous (pandas groupby order = first appearance)."""
    sid = np.asarray(session_ids)
    n = len(sid)
    if n == 0:
        return np.zeros(1, dtype=np.int64)
    starts = np.flatnonzero(np.concatenate(([True], sid[1:] != sid[:-1])))
    return np.concatenate((starts, [n])).astype(np.int64)


def build_replay_rows(session_ids, item_ids, state_len, pad_id, pad_pos="end", rewards=None):
    items = np.asarray(item_ids, dtype=np.int64)
    off = session_offsets(session_ids)
    n, L = len(items), int(state_len)
    state = np.full((n, L), pad_id, dtype=np.int64)
    nxt = np.full((n, L), pad_id, dtype=np.int64)
    ln = np.zeros(n, dtype=np.int64)
    nln = np.zeros(n, dtype=np.int64)
    end = np.zeros(n, dtype=bool)
    for s in range(len(off) - 1):
        lo, hi = int(off[s]), int(off[s + 1])
        hist = items[lo:hi]
        for i in range(lo, hi):
            nb = i - lo  # n_items_bef (preprocessing.py:225)
            # get_state (:5-29)
            if nb >= L:
                state[i] = hist[nb - L:nb]
            elif pad_pos == "end":
                state[i, :nb] = hist[:nb]
            else:
                state[i, L - nb:] = hist[:nb]
            # get_next_state (:143-170)
            if nb + 1 >= L:
                nxt[i] = hist[nb - L + 1:nb + 1]
            elif pad_pos == "end":
                nxt[i, :nb + 1] = hist[:nb + 1]
            else:
                nxt[i, L - nb - 1:] = hist[:nb + 1]
            ln[i] = min(max(nb, 1), L)       # :252-262 (first state artificially 1, capped at state_len)
            nln[i] = min(nb + 1, L)          # :264-268
            end[i] = i == hi - 1             # :230-233
    out = dict(state=state, action=items.copy(), next_state=nxt, true_state_len=ln, true_next_state_len=nln, is_end=end)
    if rewards is not None:
        out["r_act"] = np.asarray(rewards, dtype=np.float32)
    return out
