"""Pin the SARM oracle to the REAL reference class and write its golden fixtures.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference exists):

    python -m oracle.make_golden_sarm       # validates + writes tests/golden/sarm_*.npz, VALIDATION_SARM.txt

`SARM_trainer` (recommenders/models/SARM/sarm.py:78-158) is runnable on its own (only the stale loop
ikea/training/trainSARM.py:119 passes a `gamma=` keyword the class does not accept), so unlike the SMORL step the
whole train step is pinned BIT-EXACT: same seeded init, same losses, same parameters after every step.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

REF_ROOT = os.environ.get("REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)

import oracle  # noqa: E402
from oracle.make_golden import _sd_equal, _pack_sd, _batches, _store_rows, CFG_SMALL, CFG_64, B_SMALL, STEPS  # noqa: E402


def golden_sarm(SARM_trainer, report, name, cfg, packed, train_pad):
    kw = dict(hidden_dim=cfg["hidden_dim"], embedding_dim=cfg["embedding_dim"], train_pad_embed=train_pad,
              use_packed_seq=packed, learning_rate=0.01, item_num=cfg["item_num"], state_size=cfg["state_size"],
              action_dim=cfg["action_dim"], gru_layers=1, device="cpu", torch_rand_seed=118, python_rand_seed=999)
    r = SARM_trainer(**kw)
    rs = random.getstate()
    o = oracle.SARMTrainer(**kw)
    _sd_equal(r.network.state_dict(), o.network.state_dict())
    out = {}
    _pack_sd("init", r.network.state_dict(), out)
    rows, batches = _batches(cfg, B_SMALL, STEPS, seed=21)
    _store_rows(out, rows)
    losses = []
    for bt in batches:
        random.setstate(rs)
        l_ref = r.train_step(*bt)
        random.setstate(rs)
        l_or = o.train_step(*bt)
        rs = random.getstate()
        assert l_ref == l_or, (name, l_ref, l_or)
        losses.append(l_ref)
    _sd_equal(r.network.state_dict(), o.network.state_dict())
    _pack_sd("final", r.network.state_dict(), out)
    out["losses"] = np.asarray(losses, dtype=np.float64)  # [steps, 2] = (sup, mean q)
    out["meta"] = np.asarray([cfg["item_num"], cfg["action_dim"], cfg["embedding_dim"], cfg["hidden_dim"],
                              cfg["state_size"], B_SMALL, STEPS, int(packed), int(train_pad), 1])
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), **out)
    report.append(f"{name}: reference SARM_trainer == oracle.SARMTrainer bit-exact over {STEPS} steps; losses {losses}")


def main():
    torch.set_num_threads(1)
    sys.path.insert(0, REF_ROOT)
    from recommenders.models.SARM.sarm import SARM_trainer
    report = [f"torch {torch.__version__}; reference at {REF_ROOT}"]
    golden_sarm(SARM_trainer, report, "sarm_small", CFG_SMALL, packed=True, train_pad=True)
    golden_sarm(SARM_trainer, report, "sarm_unpacked", CFG_SMALL, packed=False, train_pad=True)
    golden_sarm(SARM_trainer, report, "sarm_64", CFG_64, packed=True, train_pad=True)
    with open(os.path.join(GOLD, "VALIDATION_SARM.txt"), "w") as f:
        f.write("\n".join(report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
