"""GPU parity tests of the wide-layer kernels (E, H multiples of 128: tensor-core GRU trunk of csrc/gru_tc.cu and the
K-loop head kernels of csrc/heads_tck.cu, D = 128 .. 512) against the CPU oracle, incl. BASELINE cfg3's real shape
(BidirGRU4Rec-SQN, V = N = 250 000, L = 50, E = H = 256, B = 256; reference semantics:
models/BidirGRU4Rec/model.py:51-99 + models/SQN/sqn_gru.py:183-254).  Tolerances as in test_gpu_parity.py."""
import numpy as np
import pytest
import torch

import oracle
from helpers import assert_close, assert_state_close, synced_random, report, RTOL

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ATOL_P = 2e-5


def _syn():
    from ikea_recommender_system_b200 import synthetic
    return synthetic


@pytest.mark.parametrize("family,packed,E,H,B,L", [("sqn", True, 128, 128, 70, 9), ("bidir_sqn", True, 128, 128, 300, 12),
                                                     ("bidir_sqn", False, 128, 256, 37, 7), ("gru4rec", True, 256, 128, 130, 20)])
def test_wide_forward_matches_oracle(pkg, family, packed, E, H, B, L):
    """Final states + logits through the tensor-core GRU (ragged lengths, partial session blocks, both directions)."""
    torch.manual_seed(5)
    N = 333
    kw = dict(hidden_dim=H, embedding_dim=E, item_num=N, state_size=L, action_dim=N, gru_layers=1, use_packed_seq=packed)
    onet = oracle.SessionNet(family=family, **kw)
    with torch.no_grad():
        onet.embedding.weight.mul_(20.0)  # make the recurrence non-trivial
    if family == "gru4rec":
        net = pkg.GRU4Rec(hidden_size=H, embedding_dim=E, item_num=N, state_size=L, action_dim=N, use_packed_seq=packed)
    else:
        net = pkg.SQN_Network(hidden_dim=H, item_num=N, state_size=L, action_dim=N, gamma=0.5, gru_layers=1,
                              embedding_dim=E, use_packed_seq=packed, bidirectional=family == "bidir_sqn")
    net.load_state_dict(onet.state_dict())
    net.to(DEV).eval()
    onet.eval()
    rows = _syn().make_replay_rows(B, N, L, seed=2)
    s, a, _, _, ln, _, _ = _syn().as_torch_batch(rows, 0, B)
    with torch.no_grad():
        want = onet(s, ln)
        h_want = onet.final_state(s, ln)
    h_got = net.final_state(s, ln)
    err = (h_got.cpu().double() - h_want.double()).abs().max()
    report(f"wide forward {family} E={E} H={H} B={B} L={L}: max abs err of the final state", float(err))
    assert_close(h_got, h_want, rtol=1e-4, atol=2e-6, what="final state")
    got = net(s, ln)
    if isinstance(want, tuple):
        for g_, w_ in zip(got, want):
            assert_close(g_, w_, rtol=1e-4, atol=1e-5, what="logits")
    else:
        assert_close(got, want, rtol=1e-4, atol=1e-5, what="logits")


@pytest.mark.parametrize("bidir,packed,B", [(False, True, 96), (True, False, 200)])
def test_wide_sqn_train_steps_match_oracle(pkg, bidir, packed, B):
    """SQN steps at E = H = 128 (D = 128 / 256): tensor-core GRU forward + BPTT + weight-gradient GEMMs, K-loop heads."""
    kw = dict(hidden_dim=128, embedding_dim=128, train_pad_embed=True, use_packed_seq=packed, learning_rate=0.005,
              item_num=2500, state_size=11, action_dim=2500, gamma=0.5, gru_layers=1)
    ref = oracle.SQNTrainer(family="bidir_sqn" if bidir else "sqn", **kw)
    t = pkg.SQN_trainer(device=DEV, bidirectional=bidir, **kw)
    t.send_to_device()
    rows = _syn().make_replay_rows(4 * B, 2500, 11, seed=8)
    rng = synced_random()
    for i in range(4):
        b = _syn().as_torch_batch(rows, i * B, (i + 1) * B)
        rng.replay(); want = ref.train_step(*b)
        rng.replay(); got = t.train_step(*b)
        rng.advance()
        assert_close(got, want, rtol=RTOL, atol=1e-5, what=f"step {i} losses")
    out = dict(outlier_frac=1e-3, outlier_atol=0.02 * 0.005 * 4)
    assert_state_close(t.DQN_1.state_dict(), ref.DQN_1.state_dict(), rtol=RTOL, atol=ATOL_P, **out)
    assert_state_close(t.DQN_2.state_dict(), ref.DQN_2.state_dict(), rtol=RTOL, atol=ATOL_P, **out)


def test_wide_supervised_bidir_gru4rec(pkg):
    """BidirGRU4Rec supervised steps at H = 128 (D = 256), no dropout."""
    kw = dict(hidden_dim=128, embedding_dim=128, gru_layers=1, dropout=0.0, train_pad_embed=True, use_packed_seq=True,
              learning_rate=0.005, item_num=1500, state_size=8, action_dim=1500)
    ref = oracle.GRUTrainer(family="bidir", **kw)
    t = pkg.BidirGRU4Rec_trainer(device=DEV, **kw)
    assert_state_close(t.gru_model.state_dict(), ref.gru_model.state_dict(), rtol=0, atol=0)
    t.send_to_device(); t.set_train(); ref.gru_model.train()
    rows = _syn().make_replay_rows(3 * 64, 1500, 8, seed=4)
    for i in range(3):
        s, a, _, _, ln, _, _ = _syn().as_torch_batch(rows, i * 64, (i + 1) * 64)
        want = ref.train_step(s, a, ln)
        got = t.train_step(s, a, ln)
        assert_close([got], [want], rtol=RTOL, atol=1e-5, what=f"step {i} loss")
    assert_state_close(t.gru_model.state_dict(), ref.gru_model.state_dict(), rtol=RTOL, atol=ATOL_P,
                       outlier_frac=1e-3, outlier_atol=0.02 * 0.005 * 3)


def test_cfg3_real_shape_against_live_oracle(pkg):
    """BASELINE configs[2] at its REAL shape: BidirGRU4Rec-SQN, V = N = 250 000, L = 50, E = H = 256 (D = 512), B = 256.
    Two steps (both twins) against the live CPU oracle: losses, final parameters of the trained twins."""
    V, L, H, B = 250_000, 50, 256, 256
    kw = dict(hidden_dim=H, embedding_dim=H, train_pad_embed=True, use_packed_seq=True, learning_rate=0.005,
              item_num=V, state_size=L, action_dim=V, gamma=0.5, gru_layers=1)
    ref = oracle.SQNTrainer(family="bidir_sqn", **kw)
    t = pkg.SQN_trainer(device=DEV, bidirectional=True, **kw)
    t.send_to_device()
    rows = _syn().make_replay_rows_fast(2 * B, V, L, seed=12)
    rng = synced_random()
    for i in range(2):
        b = _syn().as_torch_batch(rows, i * B, (i + 1) * B)
        rng.replay(); want = ref.train_step(*b)
        rng.replay(); got = t.train_step(*b)
        rng.advance()
        report(f"cfg3 real shape step {i}: (sup, q) losses native vs oracle", [list(map(float, got)), list(map(float, want))])
        assert_close(got, want, rtol=RTOL, atol=1e-5, what=f"step {i} losses")
    # every twin has taken ONE Adam step here: the update is lr * g / (|g| + eps), i.e. the RELATIVE error of a gradient
    # element shows up as an absolute error of lr x (relative error).  Elements whose gradient nearly cancels over the
    # 12 800 token positions carry percent-level relative noise in any fp32 evaluation order (observed worst case:
    # 2.3 % of lr on weight_hh_l0; the assertion reports it), hence 5 % of lr x steps for the <= 0.1 % outliers.
    out = dict(outlier_frac=1e-3, outlier_atol=0.05 * 0.005 * 2)
    for mine, theirs in ((t.DQN_1, ref.DQN_1), (t.DQN_2, ref.DQN_2)):
        sd_m, sd_r = mine.state_dict(), theirs.state_dict()
        small = {k: v for k, v in sd_m.items() if v.numel() < 5_000_000}
        assert_state_close(small, {k: sd_r[k] for k in small}, rtol=RTOL, atol=ATOL_P, **out)
        for k in sd_m:  # the big tables: a strided sample of rows (the full comparison costs minutes in float64)
            if k in small:
                continue
            a, b_ = sd_m[k].cpu(), sd_r[k]
            idx = torch.arange(0, a.shape[0], 97)
            assert_state_close({k: a[idx]}, {k: b_[idx]}, rtol=RTOL, atol=ATOL_P, **out)
