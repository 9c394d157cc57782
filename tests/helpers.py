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


def assert_state_close(sd_mine, sd_ref, rtol=RTOL, atol=1e-5, skip=(), outlier_frac=0.0, outlier_atol=0.0):
    """Parameter parity after Adam steps.  Adam divides by sqrt(v): an element whose gradient nearly
    cancels (|g| tiny against its summands) turns 1e-7-level summation-order noise into an O(1e-2 * lr)
    change of the update -- the CPU reference itself moves by ~4e-6 when its BLAS thread count changes.
    So besides the element-wise bound, a fraction `outlier_frac` of the elements of a tensor may exceed it,
    but never by more than `outlier_atol` (a few percent of lr * steps)."""
    assert list(sd_mine.keys()) == list(sd_ref.keys())
    worst = {}
    for k in sd_ref:
        if k in skip:
            continue
        if outlier_frac <= 0:
            assert_close(sd_mine[k], sd_ref[k], rtol, atol, what=k)
            continue
        a = torch.as_tensor(sd_mine[k]).detach().cpu().double()
        b = torch.as_tensor(sd_ref[k]).detach().cpu().double()
        assert a.shape == b.shape, (k, a.shape, b.shape)
        err = (a - b).abs()
        bad = err > atol + rtol * b.abs()
        n_bad = int(bad.sum())
        worst[k] = dict(n=bad.numel(), beyond_rtol=n_bad, max_abs_err=float(err.max()),
                        max_rel_err=float((err / (b.abs() + atol / max(rtol, 1e-12))).max()))
        assert n_bad <= max(1, int(outlier_frac * bad.numel())), f"{k}: {n_bad}/{bad.numel()} beyond rtol={rtol}"
        assert float(err.max()) <= outlier_atol, f"{k}: max abs err {float(err.max()):.3e} > {outlier_atol}"
    if worst:
        report("assert_state_close (observed worst case per tensor)", worst)


class synced_random:
    """Run oracle and native steps from the same python-RNG state."""

    def __init__(self):
        self.state = random.getstate()

    def replay(self):
        random.setstate(self.state)

    def advance(self):
        self.state = random.getstate()


# ---- well-posedness of exact-id comparisons --------------------------------------------------------------
def topk_margin(scores64, k):
    """Smallest gap between neighbours among the k+1 best scores of any row (float64 scores).

    Exact top-k / argmax ids are only a well-posed comparison where the reference's own fp32 arithmetic separates
    the candidates: two fp32 implementations of h.W + b that sum in different orders differ by a few ulp, so a
    pair of logits closer than that has no defined order.  Tests that compare ids at default-initialised
    (unscaled) weights assert `topk_margin(...) > MARGIN_MIN` on the oracle's float64 logits first, which pins the
    seeds to inputs every correct fp32 implementation must rank identically."""
    top = torch.topk(scores64, min(k + 1, scores64.shape[1]), dim=1).values
    return float((top[:, :-1] - top[:, 1:]).min())


MARGIN_MIN = 2e-7  # >= 10 fp32 ulp at |logit| <= 0.25 (default-init logits are O(0.1))


def double_copy(net):
    import copy
    return copy.deepcopy(net).double()


def report(what, payload):
    """Observed errors / margins of a parity comparison -> gpurun_out/parity_report.jsonl (merged back from the GPU
    box; copied into profiles/ per round) and stdout (visible with pytest -s)."""
    import json
    line = json.dumps({"what": what, "observed": payload}, default=float)
    print("[parity]", line)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.jsonl"), "a") as f:
            f.write(line + "\n")
    except OSError:
        pass
