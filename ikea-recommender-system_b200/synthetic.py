"""Synthetic session data shaped like the reference's replay buffer / evaluation set.

Follows the output contract of `preprocess_train_data`
(recommenders/data_utils/preprocessing.py:385-489) and the `ReplayBuffer.__getitem__`
tuple layout (recommenders/ikea/data_utils/replay_buffer.py:65-74):
    (state[L], action, r_act, next_state[L], true_state_len, true_next_state_len, is_end)
with "end" padding by pad id N, `true_state_len = clip(n_before, 1, L)`,
`true_next_state_len = clip(n_before + 1, 1, L)`, `is_end` on each session's last row.
Item popularity is Zipf-like, session lengths geometric clipped to [2, 50]
(SURVEY.md section 8d).  Pure numpy; used by tests, bench.py and smoke().
"""

from __future__ import annotations

import numpy as np
import torch


def zipf_items(rng: np.random.Generator, n_items: int, size: int, a: float = 1.05) -> np.ndarray:
    """Item ids in [0, n_items) with a Zipf(a)-like head/tail profile (inverse-CDF on ranks)."""
    ranks = np.arange(1, n_items + 1, dtype=np.float64)
    cdf = np.cumsum(ranks ** (-a))
    cdf /= cdf[-1]
    u = rng.random(size)
    r = np.searchsorted(cdf, u, side="left")
    # scatter popularity ranks over the id space so that hot items are not the lowest ids
    perm_mul = 2654435761 % n_items
    if np.gcd(perm_mul, n_items) != 1:
        perm_mul = 1
    return ((r.astype(np.int64) * perm_mul) % n_items).astype(np.int64)


def make_replay_rows(n_rows: int, n_items: int, state_len: int, seed: int = 0, pad_id=None,
                     pad_pos: str = "end", zipf_a: float = 1.05):
    """Returns dict of numpy arrays with >= n_rows replay-buffer rows (trimmed to n_rows)."""
    rng = np.random.default_rng(seed)
    pad = n_items if pad_id is None else pad_id
    L = state_len
    states, nxt, acts, lens, nlens, ends = [], [], [], [], [], []
    total = 0
    while total < n_rows:
        m = int(np.clip(rng.geometric(0.18) + 1, 2, 50))
        items = zipf_items(rng, n_items, m, zipf_a)
        for i in range(m):
            hist = items[max(0, i - L):i]
            nh = items[max(0, i + 1 - L):i + 1]
            st = np.full(L, pad, dtype=np.int64)
            ns = np.full(L, pad, dtype=np.int64)
            if pad_pos == "end":
                st[:len(hist)] = hist
                ns[:len(nh)] = nh
            else:
                if len(hist):
                    st[L - len(hist):] = hist
                ns[L - len(nh):] = nh
            states.append(st)
            nxt.append(ns)
            acts.append(items[i])
            lens.append(min(max(i, 1), L))
            nlens.append(min(i + 1, L))
            ends.append(i == m - 1)
        total += m
    out = dict(
        state=np.stack(states)[:n_rows],
        action=np.asarray(acts, dtype=np.int64)[:n_rows],
        next_state=np.stack(nxt)[:n_rows],
        true_state_len=np.asarray(lens, dtype=np.int64)[:n_rows],
        true_next_state_len=np.asarray(nlens, dtype=np.int64)[:n_rows],
        is_end=np.asarray(ends, dtype=bool)[:n_rows],
    )
    out["r_act"] = np.where(rng.random(n_rows) < 0.04, 1.0, 0.2).astype(np.float32)
    return out


def make_replay_rows_fast(n_rows: int, n_items: int, state_len: int, seed: int = 0, pad_id=None,
                          zipf_a: float = 1.05):
    """Vectorised variant for large benches ("end" padding only): same contract, no python loop."""
    rng = np.random.default_rng(seed)
    pad = n_items if pad_id is None else pad_id
    L = state_len
    n_sess = max(8, int(n_rows / 4))
    m = np.clip(rng.geometric(0.18, size=n_sess) + 1, 2, 50)
    while m.sum() < n_rows:
        m = np.concatenate([m, np.clip(rng.geometric(0.18, size=n_sess) + 1, 2, 50)])
    starts = np.concatenate([[0], np.cumsum(m)[:-1]])
    total = int(m.sum())
    items = zipf_items(rng, n_items, total, zipf_a)
    sess = np.repeat(np.arange(len(m)), m)
    pos = np.arange(total) - starts[sess]  # n_items_bef
    padded = np.concatenate([items, np.full(L + 1, pad, dtype=np.int64)])
    ar = np.arange(L)[None, :]
    # state: items[i-len .. i) left-aligned, len = min(pos, L)
    ln = np.minimum(pos, L)
    base = (np.arange(total) - ln)[:, None] + ar
    st = np.where(ar < ln[:, None], padded[np.clip(base, 0, total + L)], pad)
    ln2 = np.minimum(pos + 1, L)
    base2 = (np.arange(total) + 1 - ln2)[:, None] + ar
    ns = np.where(ar < ln2[:, None], padded[np.clip(base2, 0, total + L)], pad)
    out = dict(
        state=st[:n_rows].astype(np.int64),
        action=items[:n_rows],
        next_state=ns[:n_rows].astype(np.int64),
        true_state_len=np.clip(pos, 1, L)[:n_rows].astype(np.int64),
        true_next_state_len=ln2[:n_rows].astype(np.int64),
        is_end=(pos == m[sess] - 1)[:n_rows],
    )
    out["r_act"] = np.where(rng.random(n_rows) < 0.04, 1.0, 0.2).astype(np.float32)
    return out


def unpopular_set_from_actions(actions: np.ndarray, quantile: float = 0.9) -> set:
    """Items whose frequency is below the `quantile` frequency quantile
    (recommenders/data_utils/item_frequency.py:8-24)."""
    ids, cnt = np.unique(actions, return_counts=True)
    thr = np.quantile(cnt, quantile)
    return set(int(i) for i in ids[cnt < thr])


def as_torch_batch(rows: dict, lo: int, hi: int):
    """(s, a, r, s_next, true_len, true_next_len, is_end) as CPU torch tensors, reference dtypes."""
    t = torch.from_numpy
    return (t(rows["state"][lo:hi]), t(rows["action"][lo:hi]), t(rows["r_act"][lo:hi]),
            t(rows["next_state"][lo:hi]), t(rows["true_state_len"][lo:hi]),
            t(rows["true_next_state_len"][lo:hi]), t(rows["is_end"][lo:hi]))
