"""Shared helpers for the parity tests (oracle = checker, CUDA path = thing under test)."""
import os
import random

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-3  # north_star tolerance for loss / Q-values / gradients (fp32 accumulate)


def load_golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def sd_from_golden(g, prefix):
    return {k[len(prefix) + 1:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix + "/")}


def rows_from_golden(g):
    return {k[5:]: g[k] for k in g.files if k.startswith("rows/")}


def assert_close(a, b, rtol=RTOL, atol=1e-5, what=""):
    a = torch.as_tensor(a).detach().cpu().double()
    b = torch.as_tensor(b).detach().cpu().double()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bool(bad.any()), f"{what}: {int(bad.sum())}/{bad.numel()} off, max abs err {float(err.max()):.3e}, " \
                                f"max ratio {float((err / tol).max()):.2f}"


def assert_state_close(sd_mine, sd_ref, rtol=RTOL, atol=1e-5, skip=()):
    assert list(sd_mine.keys()) == list(sd_ref.keys())
    for k in sd_ref:
        if k in skip:
            continue
        assert_close(sd_mine[k], sd_ref[k], rtol, atol, what=k)


class synced_random:
    """Run oracle and native steps from the same python-RNG state."""

    def __init__(self):
        self.state = random.getstate()

    def replay(self):
        random.setstate(self.state)

    def advance(self):
        self.state = random.getstate()
