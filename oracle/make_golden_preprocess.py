"""Pins oracle/preprocess.py against the UNMODIFIED reference (`/root/reference`, build container only):
runs `preprocess_train_data_incl_act_rew` (recommenders/data_utils/preprocessing.py:199-320) on a seeded event log for
both padding positions and writes tests/golden/preprocess_rr.npz (log + the reference's output).

    python -m oracle.make_golden_preprocess
"""
import os
import sys
import tempfile

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def make_log(seed=3, n_sessions=40, n_items=30, max_len=17):
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, max_len + 1, size=n_sessions)
    lens[:3] = [1, 2, max_len]  # singleton, pair, longest
    sid = np.repeat(np.arange(100, 100 + n_sessions), lens)
    items = rng.integers(0, n_items, size=int(lens.sum()))
    is_buy = (rng.random(len(items)) < 0.1).astype(np.int64)
    return sid, items, is_buy


def main():
    sys.path.insert(0, "/root/reference")
    from recommenders.data_utils.preprocessing import preprocess_train_data_incl_act_rew as ref_fn
    from oracle.preprocess import build_replay_rows
    sid, items, is_buy = make_log()
    L, pad = 6, 30
    out = dict(session_id=sid, item_id=items, is_buy=is_buy, state_len=np.int64(L), pad_id=np.int64(pad))
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "log.pkl")
        pd.DataFrame(dict(session_id=sid, item_id=items, is_buy=is_buy)).to_pickle(path)
        for pos in ("end", "beg"):
            # incl_reward=False: the reward branch (`Series.apply(action_to_reward, 1, ...)`, :270-273) does not run on
            # this container's pandas; the reward is an elementwise map of action_type and is passed through as is
            df = ref_fn(path, padding_id=pad, state_len=L, incl_reward=False, pad_pos=pos)
            ref = dict(state=np.stack(df["state"].values), next_state=np.stack(df["next_state"].values),
                       action=df["action"].to_numpy(), true_state_len=df["true_state_len"].to_numpy(),
                       true_next_state_len=df["true_next_state_len"].to_numpy(), is_end=df["is_end"].to_numpy())
            assert np.array_equal(df["action_type"].to_numpy(), is_buy)
            mine = build_replay_rows(sid, items, L, pad, pos)
            for k in ref:
                assert np.array_equal(np.asarray(ref[k]), np.asarray(mine[k])), (pos, k)
                out[f"{pos}_{k}"] = np.asarray(ref[k])
    np.savez_compressed(os.path.join(GOLD, "preprocess_rr.npz"), **out)
    msg = (f"preprocess_rr: reference preprocess_train_data_incl_act_rew == oracle.preprocess.build_replay_rows exactly "
           f"({len(items)} events, {len(np.unique(sid))} sessions, state_len {L}, pad 'end' and 'beg')")
    with open(os.path.join(GOLD, "VALIDATION.txt"), "a") as f:
        f.write(msg + "\n")
    print(msg)


if __name__ == "__main__":
    main()
