"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI by the
drop-in classes, against (a) the CPU oracle on the same seeded inputs and (b) the golden fixtures
produced from the real reference.  Tolerances: indices / top-k ids exact; loss, Q-values, parameters
after Adam within 1e-3 relative (north_star), atol 2e-5 for near-zero parameters."""
import os
import random

import numpy as np
import pytest
import torch

import oracle
from helpers import (load_golden, sd_from_golden, rows_from_golden, assert_close, assert_state_close,
                     synced_random, RTOL)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ATOL_P = 2e-5
# outliers (see helpers.assert_state_close): <= 0.1 % of a tensor, each within 2 % of lr * steps
OUT = dict(outlier_frac=1e-3, outlier_atol=0.02 * 0.01 * 4)


def _syn():
    from ikea_recommender_system_b200 import synthetic
    return synthetic


def _batches(rows, B, steps):
    return [_syn().as_torch_batch(rows, i * B, (i + 1) * B) for i in range(steps)]


def _meta(g):
    m = g["meta"]
    return dict(item_num=int(m[0]), action_dim=int(m[1]), embedding_dim=int(m[2]), hidden_dim=int(m[3]),
                state_size=int(m[4])), int(m[5]), int(m[6]), bool(m[7]), bool(m[8]), int(m[9])


def test_native_library_is_loaded(pkg):
    import ctypes
    assert pkg.LIB.rec_abi_version() == 1
    with open("/proc/self/maps") as f:
        assert "librecsys_b200.so" in f.read()


# ------------------------------------------------------------------------------------ forward
@pytest.mark.parametrize("family,packed,E,H", [("gru4rec", True, 64, 64), ("gru4rec", False, 12, 20),
                                                ("bidir", True, 16, 24), ("sqn", True, 64, 64),
                                                ("smorl", True, 32, 64), ("bidir_sqn", True, 64, 64)])
def test_forward_matches_oracle(pkg, family, packed, E, H):
    torch.manual_seed(5)
    N, L, B = 333, 9, 37
    kw = dict(hidden_dim=H, embedding_dim=E, item_num=N, state_size=L, action_dim=N, gru_layers=1,
              use_packed_seq=packed)
    onet = oracle.SessionNet(family=family, **kw)
    with torch.no_grad():
        onet.embedding.weight.mul_(20.0)  # make the recurrence non-trivial
    if family == "gru4rec":
        net = pkg.GRU4Rec(hidden_size=H, embedding_dim=E, item_num=N, state_size=L, action_dim=N, use_packed_seq=packed)
    elif family == "bidir":
        net = pkg.BidirGRU4Rec(hidden_size=H, embedding_dim=E, item_num=N, state_size=L, action_dim=N, use_packed_seq=packed)
    elif family in ("sqn", "bidir_sqn"):
        net = pkg.SQN_Network(hidden_dim=H, item_num=N, state_size=L, action_dim=N, gamma=0.5, gru_layers=1,
                              embedding_dim=E, use_packed_seq=packed, bidirectional=family == "bidir_sqn")
    else:
        net = pkg.SMORL_GRU_Net(hidden_dim=H, embedding_dim=E, item_num=N, state_size=L, action_dim=N,
                                q_weights=torch.ones(3), gamma=0.5, use_packed_seq=packed)
    assert list(net.state_dict().keys()) == list(onet.state_dict().keys())
    net.load_state_dict(onet.state_dict())
    net.to(DEV).eval()
    onet.eval()
    rows = _syn().make_replay_rows(B, N, L, seed=2)
    s, a, _, _, ln, _, _ = _syn().as_torch_batch(rows, 0, B)
    with torch.no_grad():
        want = onet(s, ln)
        h_want = onet.final_state(s, ln)
    got = net(s, ln)
    assert_close(net.final_state(s, ln), h_want, rtol=1e-4, atol=1e-6, what="final state")
    if isinstance(want, tuple):
        for g_, w_ in zip(got, want):
            assert g_.shape == w_.shape
            assert_close(g_, w_, rtol=1e-4, atol=1e-5, what="logits")
    else:
        assert_close(got, want, rtol=1e-4, atol=1e-5, what="logits")


def test_zero_length_raises_like_pack_padded_sequence(pkg):
    net = pkg.GRU4Rec(hidden_size=8, embedding_dim=8, item_num=10, state_size=4, action_dim=10).to(DEV)
    with pytest.raises(RuntimeError):
        net(torch.zeros(2, 4, dtype=torch.long), torch.tensor([1, 0]))


# ------------------------------------------------------------------------------------ supervised
SUP = [("gru4rec_small", "gru4rec"), ("gru4rec_unpacked_frozenpad", "gru4rec"), ("gru4rec_2layer", "gru4rec"),
       ("bidir_small", "bidir"), ("bidir_unpacked", "bidir")]


@pytest.mark.parametrize("name,family", SUP)
def test_supervised_train_steps_match_reference_fixture(pkg, name, family):
    g = load_golden(name)
    cfg, B, steps, packed, train_pad, layers = _meta(g)
    kw = dict(hidden_dim=cfg["hidden_dim"], embedding_dim=cfg["embedding_dim"], gru_layers=layers,
              train_pad_embed=train_pad, use_packed_seq=packed, learning_rate=0.01, item_num=cfg["item_num"],
              state_size=cfg["state_size"], action_dim=cfg["action_dim"], device=DEV)
    t = pkg.BidirGRU4Rec_trainer(dropout=0.0, **kw) if family == "bidir" else pkg.GRU4Rec_trainer(**kw)
    # the drop-in's seeded init IS the reference's init (same RNG consumption)
    assert_state_close(t.gru_model.state_dict(), sd_from_golden(g, "init"), rtol=0, atol=0)
    t.send_to_device()
    t.set_train()
    batches = _batches(rows_from_golden(g), B, steps)
    s, a, _, _, ln, _, _ = batches[0]
    assert_close(t.gru_model(s, ln), g["fwd_logits0"], rtol=1e-4, atol=1e-5, what="fwd logits")
    losses = [t.train_step(b[0], b[1], b[4]) for b in batches]
    assert_close(losses, g["losses"], rtol=RTOL, atol=1e-6, what="losses")
    assert_state_close(t.gru_model.state_dict(), sd_from_golden(g, "final"), rtol=RTOL, atol=ATOL_P, **OUT)


# ------------------------------------------------------------------------------------ SQN / SMORL
@pytest.mark.parametrize("name", ["sqn_small", "sqn_unpacked", "sqn_64"])
def test_sqn_train_steps_match_reference_fixture(pkg, name):
    g = load_golden(name)
    cfg, B, steps, packed, train_pad, layers = _meta(g)
    t = pkg.SQN_trainer(hidden_dim=cfg["hidden_dim"], embedding_dim=cfg["embedding_dim"], train_pad_embed=train_pad,
                        use_packed_seq=packed, learning_rate=0.01, item_num=cfg["item_num"],
                        state_size=cfg["state_size"], action_dim=cfg["action_dim"], gamma=0.5, gru_layers=layers,
                        device=DEV)
    assert_state_close(t.DQN_1.state_dict(), sd_from_golden(g, "init1"), rtol=0, atol=0)
    assert_state_close(t.DQN_2.state_dict(), sd_from_golden(g, "init2"), rtol=0, atol=0)
    t.send_to_device()
    losses, mains = [], []
    for b in _batches(rows_from_golden(g), B, steps):
        losses.append(t.train_step(*b))
        mains.append(t.last_main)
    assert mains == list(g["mains"])
    assert_close(losses, g["losses"], rtol=RTOL, atol=1e-5, what="(sup, q) losses")
    assert_state_close(t.DQN_1.state_dict(), sd_from_golden(g, "final1"), rtol=RTOL, atol=ATOL_P, **OUT)
    assert_state_close(t.DQN_2.state_dict(), sd_from_golden(g, "final2"), rtol=RTOL, atol=ATOL_P, **OUT)


def test_smorl_train_steps_match_oracle_fixture(pkg):
    g = load_golden("smorl_small")
    cfg, B, steps, *_ = _meta(g)
    unpop = set(int(i) for i in g["unpop"])
    e_div = torch.nn.Embedding.from_pretrained(torch.from_numpy(g["e_div"]), freeze=True)
    t = pkg.SMORL_trainer(hidden_dim=cfg["hidden_dim"], embedding_dim=cfg["embedding_dim"], padding_pos="end",
                          train_pad_embed=True, use_packed_seq=True, learning_rate=0.01, item_num=cfg["item_num"],
                          state_size=cfg["state_size"], action_dim=cfg["action_dim"], gamma=0.5, gru_layers=1,
                          q_weights=torch.tensor([1.0, 0.7, 0.4]), alpha=0.8, div_embedding=e_div,
                          unpopular_actions_set=unpop, topk_div=3, device=DEV, topk_nov=2, nov_rew_sig=1.0)
    assert_state_close(t.SMORL_1.state_dict(), sd_from_golden(g, "init1"), rtol=0, atol=0)
    t.send_to_device()
    batches = _batches(rows_from_golden(g), B, steps)
    s, a, _, _, ln, _, _ = batches[0]
    sup, q = t.SMORL_1(s, ln)
    assert q.shape == (B, 3, cfg["action_dim"])
    assert_close(sup, g["fwd_sup0"], rtol=1e-4, atol=1e-5)  # real reference net output
    assert_close(q, g["fwd_q0"], rtol=1e-4, atol=1e-5)
    losses, mains = [], []
    for b in batches:
        losses.append(t.train_step(*b))
        mains.append(t.last_main)
    assert mains == list(g["mains"])
    assert_close(losses, g["losses"], rtol=RTOL, atol=1e-5, what="(sup, q) losses")
    assert_state_close(t.SMORL_1.state_dict(), sd_from_golden(g, "final1"), rtol=RTOL, atol=ATOL_P, **OUT)
    assert_state_close(t.SMORL_2.state_dict(), sd_from_golden(g, "final2"), rtol=RTOL, atol=ATOL_P, **OUT)


def test_sqn_cfg2_shapes_against_live_oracle(pkg):
    """BASELINE cfg2 shapes (V=N=70852, B=256, L=10, E=H=64): 2 steps against the oracle run live."""
    kw = dict(hidden_dim=64, embedding_dim=64, train_pad_embed=True, use_packed_seq=True, learning_rate=0.01,
              item_num=70852, state_size=10, action_dim=70852, gamma=0.5, gru_layers=1)
    ref = oracle.SQNTrainer(**kw)
    t = pkg.SQN_trainer(device=DEV, **kw)
    assert_state_close(t.DQN_1.state_dict(), ref.DQN_1.state_dict(), rtol=0, atol=0)
    t.send_to_device()
    rows = _syn().make_replay_rows_fast(512, 70852, 10, seed=4)
    rng = synced_random()
    for i in range(2):
        b = _syn().as_torch_batch(rows, i * 256, (i + 1) * 256)
        rng.replay(); want = ref.train_step(*b)
        rng.replay(); got = t.train_step(*b)
        rng.advance()
        assert t.last_main == ref.last_main
        assert_close(got, want, rtol=RTOL, atol=1e-5, what=f"step {i} losses")
    assert_state_close(t.DQN_1.state_dict(), ref.DQN_1.state_dict(), rtol=RTOL, atol=ATOL_P, **OUT)
    assert_state_close(t.DQN_2.state_dict(), ref.DQN_2.state_dict(), rtol=RTOL, atol=ATOL_P, **OUT)


def test_bidir_sqn_cfg3_like(pkg):
    """cfg3 family (bidirectional trunk under the SQN heads) at a reduced catalogue, H=E=128, L=20."""
    kw = dict(hidden_dim=128, embedding_dim=128, train_pad_embed=True, use_packed_seq=True, learning_rate=0.005,
              item_num=3000, state_size=20, action_dim=3000, gamma=0.5, gru_layers=1)
    ref = oracle.SQNTrainer(family="bidir_sqn", **kw)
    t = pkg.SQN_trainer(device=DEV, bidirectional=True, **kw)
    assert_state_close(t.DQN_1.state_dict(), ref.DQN_1.state_dict(), rtol=0, atol=0)
    t.send_to_device()
    rows = _syn().make_replay_rows(3 * 48, 3000, 20, seed=6)
    rng = synced_random()
    for i in range(3):
        b = _syn().as_torch_batch(rows, i * 48, (i + 1) * 48)
        rng.replay(); want = ref.train_step(*b)
        rng.replay(); got = t.train_step(*b)
        rng.advance()
        assert_close(got, want, rtol=RTOL, atol=1e-5, what=f"step {i} losses")
    assert_state_close(t.DQN_1.state_dict(), ref.DQN_1.state_dict(), rtol=RTOL, atol=ATOL_P, **OUT)
    assert_state_close(t.DQN_2.state_dict(), ref.DQN_2.state_dict(), rtol=RTOL, atol=ATOL_P, **OUT)


# ------------------------------------------------------------------------------------ evaluation
def _eval_loader(rows, n, bs):
    out = []
    for lo in range(0, n, bs):
        s, a, _, _, ln, _, _ = _syn().as_torch_batch(rows, lo, min(lo + bs, n))
        out.append((s, a, ln))
    return out


def test_evaluate_matches_reference_fixture(pkg):
    g = load_golden("eval_sqn64")
    net = pkg.SQN_Network(hidden_dim=64, item_num=500, state_size=10, action_dim=500, gamma=0.5, gru_layers=1,
                          embedding_dim=64, use_packed_seq=True)
    net.load_state_dict(sd_from_golden(g, "net"))
    net.to(DEV)
    loader = _eval_loader(rows_from_golden(g), 90, 32)
    unpop = set(int(i) for i in g["unpop"])
    e_div = torch.nn.Embedding.from_pretrained(torch.from_numpy(g["e_div"]), freeze=True)
    kw = dict(head_idx=0, topk_hr_ndcg=[5, 10, 20], topk_to_consider_div=3, topk_to_consider_nov=2,
              topk_to_consider_cov=[1, 5, 10, 20], novelty_rew_signal=1)
    loss, hr, ndcg, cov, div, nov, reps = pkg.evaluate(loader, net, DEV, torch.nn.CrossEntropyLoss(), "end", e_div,
                                                       unpop, **kw)
    assert isinstance(loss, torch.Tensor) and isinstance(hr, np.ndarray) and isinstance(cov, dict)
    assert_close(loss, g["loss"], rtol=1e-4)
    assert np.array_equal(hr, g["hr"]), (hr, g["hr"])
    assert np.allclose(ndcg, g["ndcg"], rtol=1e-12)
    assert np.array_equal(reps, g["reps"])
    assert np.allclose([cov[k] for k in sorted(cov)], g["cov_vals"], rtol=0, atol=0)
    assert_close(div, g["div"], rtol=1e-4)
    assert np.isclose(nov, g["nov"], rtol=1e-12)
    # update_train_metrics on the first batch
    s, a, ln = loader[0]
    out = pkg.update_train_metrics(s, a, ln, net, DEV, "end", e_div, unpop, {k: set() for k in [1, 5, 10, 20]}, **kw)
    assert np.array_equal(out[0], g["utm_hr"]) and np.allclose(out[1], g["utm_ndcg"]) and np.array_equal(out[5], g["utm_reps"])
    assert_close(out[3], g["utm_div"], rtol=1e-4)
    assert np.isclose(out[4], g["utm_nov"])
    assert all(isinstance(v, set) for v in out[2].values())


@pytest.mark.parametrize("hidden,V,B", [(8, 300, 70), (64, 300, 70), (64, 40000, 1030), (128, 40000, 1030)])
def test_topk_ties_lowest_id_first(pkg, hidden, V, B):
    """Exact ties: zero head weights, bias with repeated values -> logits == bias for every session.
    hidden 8: CUDA-core head kernels; 64 / small: tcgen05 head kernels; B >= 1024 and V >= 32768: the chunk-maxima path
    (massive ties overflow its candidate list: the exact fallback must still return the lowest ids)."""
    N, K = 50, 20
    net = pkg.GRU4Rec(hidden_size=hidden, embedding_dim=hidden, item_num=N, state_size=5, action_dim=V)
    rng = np.random.default_rng(0)
    bias = torch.from_numpy(rng.integers(0, 6, size=V).astype(np.float32))
    with torch.no_grad():
        net.output.weight.zero_()
        net.output.bias.copy_(bias)
    net.to(DEV)
    eng = net._ready(B)
    rows = _syn().make_replay_rows(B, N, 5, seed=1)
    s, a, _, _, ln, _, _ = _syn().as_torch_batch(rows, 0, B)
    want = oracle.stable_topk(bias.repeat(B, 1), K)
    ids = torch.empty(B, K, dtype=torch.int32, device=DEV)
    sc = torch.empty(B, K, dtype=torch.float32, device=DEV)
    from ikea_recommender_system_b200 import _native as N_
    from ikea_recommender_system_b200.engine import EvalAccumulators
    o = N_.RecEvalOpts()
    o.head_idx, o.n_k, o.n_cov = 0, 1, 0
    o.ks[0] = K
    acc = EvalAccumulators(torch.device(DEV), V)
    ds, dl = net._dev_inputs(s, ln)
    da = (a % V).to(DEV)
    eng.eval_batch(0, eng._batch(B, ds, da, dl), o, acc.struct, topk_ids=ids, topk_scores=sc)
    assert torch.equal(ids.cpu().long(), want)
    assert torch.equal(sc.cpu(), bias[want])


def test_evaluate_large_catalog_against_live_oracle(pkg):
    """V = 70852, one batch of 300 sessions: ids exact, metrics equal."""
    torch.manual_seed(11)
    N = 70852
    onet = oracle.make_gru4rec(hidden_dim=64, embedding_dim=64, item_num=N, state_size=10, action_dim=N,
                               gru_layers=1, use_packed_seq=True)
    with torch.no_grad():
        onet.embedding.weight.mul_(30.0)
        onet.output.weight.mul_(20.0)
    net = pkg.GRU4Rec(hidden_size=64, embedding_dim=64, item_num=N, state_size=10, action_dim=N)
    net.load_state_dict(onet.state_dict())
    net.to(DEV)
    rows = _syn().make_replay_rows_fast(300, N, 10, seed=9)
    loader = _eval_loader(rows, 300, 300)
    unpop = _syn().unpopular_set_from_actions(rows["action"])
    torch.manual_seed(1)
    e_div = torch.nn.Embedding.from_pretrained(torch.randn(N + 1, 16), freeze=True)
    kw = dict(head_idx=0, topk_hr_ndcg=[5, 10, 20], topk_to_consider_div=2, topk_to_consider_nov=1,
              topk_to_consider_cov=[1, 5, 10, 20], novelty_rew_signal=1)
    want = oracle.evaluate(loader, onet, torch.nn.CrossEntropyLoss(), "end", e_div, unpop, **kw)
    got = pkg.evaluate(loader, net, DEV, torch.nn.CrossEntropyLoss(), "end", e_div, unpop, **kw)
    assert_close(got[0], want[0], rtol=1e-4)
    assert np.array_equal(got[1], want[1]) and np.allclose(got[2], want[2]) and np.array_equal(got[6], want[6])
    assert got[3] == want[3]
    assert_close(got[4], want[4], rtol=1e-4)
    assert np.isclose(got[5], want[5])


def test_eval_properties_at_1m_items(pkg):
    """cfg5 size (V = 1M): size-independent properties -- sorted scores, unique ids, idempotence,
    consistency of HR with the materialised logits of a few rows."""
    N = 1_000_000
    torch.manual_seed(3)
    net = pkg.GRU4Rec(hidden_size=64, embedding_dim=64, item_num=N, state_size=10, action_dim=N)
    with torch.no_grad():
        net.embedding.weight.mul_(30.0)
        net.output.weight.mul_(10.0)
    net.to(DEV)
    B, K = 128, 20
    rows = _syn().make_replay_rows_fast(B, N, 10, seed=5)
    s, a, _, _, ln, _, _ = _syn().as_torch_batch(rows, 0, B)
    from ikea_recommender_system_b200 import _native as N_
    from ikea_recommender_system_b200.engine import EvalAccumulators
    eng = net._ready(B)
    o = N_.RecEvalOpts()
    o.head_idx, o.n_k, o.n_cov = 0, 1, 1
    o.ks[0] = K
    o.cov_ks[0] = K
    ds, dl = net._dev_inputs(s, ln)
    da = a.to(DEV)
    res = []
    for _ in range(2):
        acc = EvalAccumulators(torch.device(DEV), N)
        ids = torch.empty(B, K, dtype=torch.int32, device=DEV)
        sc = torch.empty(B, K, dtype=torch.float32, device=DEV)
        eng.eval_batch(0, eng._batch(B, ds, da, dl), o, acc.struct, topk_ids=ids, topk_scores=sc)
        res.append((ids.cpu(), sc.cpu(), acc.read()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])  # idempotent
    ids, sc, r = res[0]
    assert bool((sc[:, :-1] >= sc[:, 1:]).all())
    assert all(len(set(row.tolist())) == K for row in ids)
    logits = net(s[:8], ln[:8]).cpu()
    tv, ti = torch.topk(logits, K, dim=1)
    assert torch.equal(ti, ids[:8].long())  # random fp32 scores: no ties
    assert_close(sc[:8], tv, rtol=1e-5, atol=1e-5)
    covered = int(np.unpackbits(r["cov_bits"][0].view(np.uint8)).sum())
    assert covered == len(set(ids.flatten().tolist()))


# ------------------------------------------------------------------------------------ vocabulary sharding
def _virtual_rank_step(trainers, hp_fn, batch, main, has_q=True):
    """Drive rec_train_phase_a..d for G 'virtual ranks' living on ONE GPU: the collectives become
    torch.stack / sum.  (Kernels of different ranks never wait on one another.)"""
    G = len(trainers)
    s, a, r, sn, ln, nl, e = [t.to(DEV) for t in batch]
    Bg, L = s.shape
    engs = [t._ready(Bg) for t in trainers]
    rec = engs[0].record_floats()
    r32, e8 = r.float().contiguous(), e.to(torch.uint8).contiguous()  # must outlive all four phases
    b = engs[0]._batch(Bg, s, a, ln, r32, sn, nl, e8)
    records = [torch.empty(Bg, rec, device=DEV) for _ in range(G)]
    for g in range(G):
        engs[g].train_phase_a(b, hp_fn(trainers[g]), main, records[g])
    gathered = torch.stack(records).contiguous()
    qs = [torch.zeros(2, Bg, 3, device=DEV) for _ in range(G)]
    for g in range(G):
        engs[g].train_phase_b(gathered, G, qs[g])
    q = torch.stack(qs).sum(0).contiguous()
    D = trainers[0]._nets[0].hidden_dim
    dhs = [torch.empty(Bg, D, device=DEV) for _ in range(G)]
    losses = [torch.zeros(8, device=DEV) for _ in range(G)]
    for g in range(G):
        engs[g].train_phase_c(q, losses[g], dhs[g])
    dh = torch.stack(dhs).sum(0).contiguous()
    for g in range(G):
        engs[g].train_phase_d(dh)
    torch.cuda.synchronize()
    return [l[:2].tolist() for l in losses]


@pytest.mark.parametrize("H", [64, 128])  # 128: the wide-layer kernels (K-loop heads incl. top-k, tensor-core GRU) under sharding
def test_vocab_sharded_phases_equal_unsharded_and_oracle(pkg, H):
    from ikea_recommender_system_b200.sharded import shard_bounds
    G, V = 3, 1000
    kw = dict(hidden_dim=H, embedding_dim=H, padding_pos="end", train_pad_embed=True, use_packed_seq=True,
              learning_rate=0.01, item_num=V, state_size=10, action_dim=V, gamma=0.5, gru_layers=1,
              q_weights=torch.tensor([1.0, 0.6, 0.3]), alpha=0.9, topk_div=2, topk_nov=1, nov_rew_sig=1.0)
    rows = _syn().make_replay_rows(3 * 96, V, 10, seed=8)
    unpop = _syn().unpopular_set_from_actions(rows["action"])
    torch.manual_seed(2)
    e_div = torch.nn.Embedding.from_pretrained(torch.randn(V + 1, 16), freeze=True)
    ref = oracle.SMORLTrainer(div_embedding=e_div, unpopular_actions_set=unpop, **kw)
    shards = []
    for g in range(G):
        t = pkg.SMORL_trainer(div_embedding=e_div, unpopular_actions_set=unpop, device=DEV, **kw)
        lo, hi = shard_bounds(V, g, G)
        for n in t._nets:
            n.shard_vocabulary(lo, hi)
        t.send_to_device()
        shards.append(t)
    rng = synced_random()
    for i in range(3):
        batch = _syn().as_torch_batch(rows, i * 96, (i + 1) * 96)
        rng.replay(); want = ref.train_step(*batch)
        main = ref.last_main - 1
        rng.advance()
        got = _virtual_rank_step(shards, lambda t: t._hp(), batch, main)
        for g in range(G):
            assert_close(got[g], want, rtol=RTOL, atol=1e-5, what=f"step {i} rank {g} losses")
        assert got[0] == got[1] == got[2]  # replicated quantities are bit-identical across ranks
    for net_i, full in enumerate([ref.SMORL_1, ref.SMORL_2]):
        sd = full.state_dict()
        for g in range(G):
            lo, hi = shard_bounds(V, g, G)
            mine = shards[g]._nets[net_i].state_dict()
            want_sd = {k: (sd[k][lo:hi] if ("head" in k) else sd[k]) for k in sd}
            assert_state_close(mine, want_sd, rtol=RTOL, atol=ATOL_P, outlier_frac=1e-3, outlier_atol=0.02 * 0.01 * 3)
        # replicated tensors stay bit-identical across ranks
        for k in ("embedding.weight", "base_model.weight_hh_l0"):
            assert torch.equal(shards[0]._nets[net_i].state_dict()[k], shards[1]._nets[net_i].state_dict()[k])


def _run_dist_equivalence(nproc, port, env_extra):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(root, "tests", "dist_equivalence.py")]
    env = dict(os.environ, **env_extra)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    if out.returncode != 0 or "dist_equivalence ok" not in out.stdout:
        err = out.stderr
        first = err.find("Traceback")  # the failing rank's own traceback comes before torchrun's summary of it
        raise AssertionError(out.stdout[-1500:] + (err[first:first + 4000] if first >= 0 else "") + err[-1500:])


@pytest.mark.parametrize("trunk", ["replicated", "data_parallel"])
def test_sharded_step_under_torchrun_world1(pkg, trunk):
    """tests/dist_equivalence.py with ONE rank over NCCL: the whole sharded driver (packed all-gather, phases,
    collectives, torch CUDA-graph capture + replay, and with `data_parallel` the rec_dp_* trunk entry points)
    against the oracle, on a single GPU."""
    _run_dist_equivalence(1, 29614, {"REC_DP_TRUNK": "1"} if trunk == "data_parallel" else {"REC_NO_DP_TRUNK": "1"})


@pytest.mark.parametrize("trunk", ["replicated", "data_parallel"])
def test_multi_gpu_equals_oracle_when_two_gpus_present(pkg, trunk):
    """Runs tests/dist_equivalence.py under torchrun over NCCL when the box has >= 2 GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (covered by the virtual-rank and world-1 tests on one GPU)")
    _run_dist_equivalence(2, 29613, {"REC_DP_TRUNK": "1"} if trunk == "data_parallel" else {"REC_NO_DP_TRUNK": "1"})


def test_tensor_core_heads_agree_with_cuda_core_heads(pkg):
    """A/B on the same engine: tcgen05 (bf16x3) head statistics vs the fp32 CUDA-core kernels."""
    from ikea_recommender_system_b200 import _native as N_
    from ikea_recommender_system_b200.engine import EvalAccumulators
    torch.manual_seed(4)
    N, B, K = 20000, 333, 20
    net = pkg.GRU4Rec(hidden_size=64, embedding_dim=64, item_num=N, state_size=10, action_dim=N)
    with torch.no_grad():
        net.embedding.weight.mul_(30.0)
        net.output.weight.mul_(15.0)
    net.to(DEV)
    eng = net._ready(B)
    rows = _syn().make_replay_rows_fast(B, N, 10, seed=2)
    s, a, _, _, ln, _, _ = _syn().as_torch_batch(rows, 0, B)
    ds, dl = net._dev_inputs(s, ln)
    da = a.to(DEV)
    o = N_.RecEvalOpts()
    o.head_idx, o.n_k, o.n_cov = 0, 1, 0
    o.ks[0] = K
    out = {}
    for on in (True, False):
        eng.set_tensor_cores(on)
        acc = EvalAccumulators(torch.device(DEV), N)
        ids = torch.empty(B, K, dtype=torch.int32, device=DEV)
        sc = torch.empty(B, K, dtype=torch.float32, device=DEV)
        eng.eval_batch(0, eng._batch(B, ds, da, dl), o, acc.struct, topk_ids=ids, topk_scores=sc)
        out[on] = (ids.cpu(), sc.cpu(), acc.read())
    eng.set_tensor_cores(True)
    assert torch.equal(out[True][0], out[False][0])
    assert_close(out[True][1], out[False][1], rtol=1e-4, atol=1e-4)
    assert abs(out[True][2]["loss_sum"] - out[False][2]["loss_sum"]) <= 1e-4 * abs(out[False][2]["loss_sum"])
    assert np.array_equal(out[True][2]["hits"], out[False][2]["hits"])


def test_head_gradient_wrt_state_matches_autograd(pkg):
    """dL/dh of the fused head backward (tensor-core path at D = 64) against torch autograd on the oracle:
    the 'gradients within 1e-3' clause of the north star, checked directly (phase C exposes dh)."""
    V, B, L = 3000, 96, 10
    kw = dict(hidden_dim=64, embedding_dim=64, train_pad_embed=True, use_packed_seq=True, learning_rate=0.01,
              item_num=V, state_size=L, action_dim=V, gamma=0.5, gru_layers=1)
    ref = oracle.SQNTrainer(**kw)
    t = pkg.SQN_trainer(device=DEV, **kw)
    with torch.no_grad():
        for n in (ref.DQN_1, ref.DQN_2):
            n.embedding.weight.mul_(25.0)
            n.sup_head_output.weight.mul_(8.0)
            n.q_head_output.weight.mul_(8.0)
    t.DQN_1.load_state_dict(ref.DQN_1.state_dict())
    t.DQN_2.load_state_dict(ref.DQN_2.state_dict())
    t.send_to_device()
    rows = _syn().make_replay_rows(B, V, L, seed=12)
    s, a, r, sn, ln, nl, e = _syn().as_torch_batch(rows, 0, B)
    # oracle: loss as a function of the final state of main(s)
    main, boot = ref.DQN_1, ref.DQN_2
    h = main.final_state(s, ln).detach().requires_grad_(True)
    sup, q = main.sup_head_output(h), main.q_head_output(h)
    with torch.no_grad():
        a_star = torch.argmax(main(sn, nl)[1], dim=1, keepdim=True)
        bv = boot(sn, ln)[1].gather(1, a_star)
        bv[e] = 0.0
    loss = torch.mean((r.unsqueeze(1) + 0.5 * bv - q.gather(1, a.unsqueeze(1))) ** 2) + torch.nn.functional.cross_entropy(sup, a)
    want = torch.autograd.grad(loss, h)[0]
    # native: phases A-C on the unsharded engine (one "shard")
    eng = t._ready(B)
    d = lambda x, dt: x.to(DEV, dt).contiguous()
    ds, da, dr, dsn, dln, dnl, de = d(s, torch.int64), d(a, torch.int64), d(r, torch.float32), d(sn, torch.int64), \
        d(ln, torch.int64), d(nl, torch.int64), d(e, torch.uint8)
    batch = eng._batch(B, ds, da, dln, dr, dsn, dnl, de)
    rec = torch.empty(B, eng.record_floats(), device=DEV)
    eng.train_phase_a(batch, t._hp(), 0, rec)
    qb = torch.zeros(2, B, 3, device=DEV)
    eng.train_phase_b(rec.unsqueeze(0).contiguous(), 1, qb)
    losses = torch.zeros(8, device=DEV)
    dh = torch.empty(B, 64, device=DEV)
    eng.train_phase_c(qb, losses, dh)
    eng.train_phase_d(dh)
    torch.cuda.synchronize()
    got = dh.cpu()
    scale = want.abs().max(dim=1, keepdim=True).values
    assert float(((got - want).abs() / scale).max()) <= 1e-3
    ce = float(torch.nn.functional.cross_entropy(sup, a).detach())
    assert_close(losses[:2], [ce, float(loss.detach()) - ce], rtol=RTOL, atol=1e-5)


@pytest.mark.parametrize("B,V", [(300, 3000), (600, 3000), (300, 40000)])
def test_large_batch_chunked_tensor_core_backward(pkg, B, V):
    """B > 256 exercises the chunked (non TMEM-resident dh) path of the tcgen05 backward kernel: a ragged second chunk
    (300), an odd number of chunks (600), and several tiles per CTA (40 000 items = 313 tiles on 148 SMs: the dh slices
    are accumulated across tiles in HBM)."""
    L = 10
    kw = dict(hidden_dim=64, embedding_dim=64, gru_layers=1, train_pad_embed=True, use_packed_seq=True,
              learning_rate=0.01, item_num=V, state_size=L, action_dim=V)
    ref = oracle.GRUTrainer(**kw)
    t = pkg.GRU4Rec_trainer(device=DEV, **kw)
    assert_state_close(t.gru_model.state_dict(), ref.gru_model.state_dict(), rtol=0, atol=0)
    t.send_to_device()
    rows = _syn().make_replay_rows(2 * B, V, L, seed=15)
    for i in range(2):
        s, a, _, _, ln, _, _ = _syn().as_torch_batch(rows, i * B, (i + 1) * B)
        want = ref.train_step(s, a, ln)
        got = t.train_step(s, a, ln)
        assert_close(got, want, rtol=RTOL, atol=1e-5, what=f"step {i} loss")
    assert_state_close(t.gru_model.state_dict(), ref.gru_model.state_dict(), rtol=RTOL, atol=ATOL_P, outlier_frac=1e-3,
                       outlier_atol=0.02 * 0.01 * 2)


@pytest.mark.gpu
@pytest.mark.parametrize("graphs", [True, False])
def test_host_entry_and_device_entry_agree_bitwise(pkg, graphs):
    """rec_train_step_q_host (CPU tensors in, floats out) and rec_train_step_q (device tensors) run the same
    kernels in the same order: losses and every parameter must be identical, with and without graph replay."""
    g = load_golden("sqn_64")
    cfg, B, steps, packed, train_pad, layers = _meta(g)
    batches = list(_batches(rows_from_golden(g), B, steps))

    def run(device_inputs):
        random.seed(7)
        torch.manual_seed(7)
        t = pkg.SQN_trainer(hidden_dim=cfg["hidden_dim"], embedding_dim=cfg["embedding_dim"], train_pad_embed=train_pad,
                            use_packed_seq=packed, learning_rate=0.01, item_num=cfg["item_num"],
                            state_size=cfg["state_size"], action_dim=cfg["action_dim"], gamma=0.5, gru_layers=layers,
                            device=DEV)
        t.send_to_device()
        t._ready(B).set_cuda_graphs(graphs)
        losses = []
        for rep in range(3):                       # > 2 sightings of the same shape: eager, capture, replay
            for b in batches:
                bb = tuple(x.to(DEV) for x in b) if device_inputs else b
                losses.append(tuple(t.train_step(*bb)))
        return losses, {k: v.detach().cpu().clone() for k, v in t.DQN_1.state_dict().items()}

    l_host, p_host = run(False)
    l_dev, p_dev = run(True)
    assert l_host == l_dev
    for k in p_host:
        assert torch.equal(p_host[k], p_dev[k]), k


class _FixedMaskDropout(torch.nn.Module):
    """nn.Dropout with the keep mask supplied from outside (what the oracle needs to share a mask with the GPU)."""

    def __init__(self, p):
        super().__init__()
        self.p, self.mask = p, None

    def forward(self, x):
        return x * self.mask.to(x.dtype) / (1.0 - self.p) if self.training else x


@pytest.mark.gpu
def test_bidir_dropout_train_step_matches_oracle_with_shared_mask(pkg):
    """BidirGRU4Rec/model.py:60,93: dropout on concat(h_fwd, h_bwd) in train mode.  torch's Philox stream cannot be
    reproduced on the device, so both sides get the same keep mask (A18 in SURVEY.md section 8a)."""
    V, L, B, H, p, steps = 3000, 10, 96, 64, 0.3, 4
    kw = dict(hidden_dim=H, embedding_dim=64, gru_layers=1, dropout=p, train_pad_embed=True, use_packed_seq=True,
              learning_rate=0.01, item_num=V, state_size=L, action_dim=V)
    t = pkg.BidirGRU4Rec_trainer(device=DEV, **kw)
    ref = oracle.GRUTrainer(family="bidir", **kw)
    ref.gru_model.dropout = _FixedMaskDropout(p)
    assert_state_close(t.gru_model.state_dict(), ref.gru_model.state_dict(), rtol=0, atol=0)
    t.send_to_device(); t.set_train(); ref.gru_model.train()
    rows = _syn().make_replay_rows(steps * B, V, L, seed=11)
    g = torch.Generator().manual_seed(5)
    for i in range(steps):
        s, a, _, _, ln, _, _ = _syn().as_torch_batch(rows, i * B, (i + 1) * B)
        mask = (torch.rand(B, 2 * H, generator=g) >= p).to(torch.uint8)
        ref.gru_model.dropout.mask = mask
        t.dropout_mask_override = mask
        want = ref.train_step(s, a, ln)
        got = t.train_step(s, a, ln)
        assert_close([got], [want], rtol=RTOL, atol=1e-5, what=f"step {i} loss with dropout")
    assert_state_close(t.gru_model.state_dict(), ref.gru_model.state_dict(), rtol=RTOL, atol=ATOL_P, **OUT)


@pytest.mark.gpu
def test_bidir_dropout_device_rng_is_deterministic_and_active(pkg):
    """Without an injected mask the keep mask comes from (seed, Adam step, element): same seed -> same run,
    dropout on != dropout off, eval mode ignores it."""
    V, L, B, H = 2000, 10, 64, 64
    kw = dict(hidden_dim=H, embedding_dim=64, gru_layers=1, train_pad_embed=True, use_packed_seq=True,
              learning_rate=0.01, item_num=V, state_size=L, action_dim=V)
    rows = _syn().make_replay_rows(4 * B, V, L, seed=12)

    def run(p, seed):
        t = pkg.BidirGRU4Rec_trainer(device=DEV, dropout=p, torch_rand_seed=seed, **kw)
        t.send_to_device(); t.set_train()
        out = []
        for i in range(4):
            s, a, _, _, ln, _, _ = _syn().as_torch_batch(rows, i * B, (i + 1) * B)
            out.append(t.train_step(s, a, ln))
        return out

    a1, a2, b0 = run(0.5, 118), run(0.5, 118), run(0.0, 118)
    assert a1 == a2
    assert all(abs(x - y) > 1e-6 for x, y in zip(a1[1:], b0[1:]))   # step 0 losses may coincide only by accident


# ------------------------------------------------------------------------ device-resident replay buffer
def test_device_replay_buffer_batches_equal_dataloader_and_train_identically(pkg):
    """rec_gather_batch batches are bit-identical to DataLoader(shuffle=True) batches of the same
    seeded epoch (ragged last batch included), and training from them reproduces the host-entry losses bit for bit."""
    from torch.utils.data import DataLoader
    V, L, B = 500, 10, 48
    rows = _syn().make_replay_rows(5 * B + 7, V, L, seed=12)
    arrays = dict(states=rows["state"], actions=rows["action"], reward=rows["r_act"], next_states=rows["next_state"],
                  true_state_len=rows["true_state_len"], true_next_state_len=rows["true_next_state_len"],
                  is_end=rows["is_end"])
    buf = pkg.DeviceReplayBuffer.from_arrays(**arrays).to_device(DEV)
    assert buf.bytes_on_device() > 0
    kw = dict(hidden_dim=64, embedding_dim=64, train_pad_embed=True, use_packed_seq=True, learning_rate=0.01,
              item_num=V, state_size=L, action_dim=V, gamma=0.5, gru_layers=1)
    t_dev = pkg.SQN_trainer(device=DEV, **kw)
    t_host = pkg.SQN_trainer(device=DEV, **kw)
    t_host.DQN_1.load_state_dict(t_dev.DQN_1.state_dict()); t_host.DQN_2.load_state_dict(t_dev.DQN_2.state_dict())
    t_dev.send_to_device(); t_host.send_to_device()
    eng = t_dev._ready(B)
    loader = DataLoader(buf, batch_size=B, shuffle=True, generator=torch.Generator().manual_seed(5))
    dev_iter = buf.batches(eng, B, shuffle=True, generator=torch.Generator().manual_seed(5))
    rng = synced_random()
    n_batches = 0
    for host_b, dev_b in zip(loader, dev_iter):
        s, a, r, sn, ln, nl, e = host_b
        ds, da, dr, dsn, dln, dnl, de = dev_b
        assert torch.equal(ds.cpu(), s) and torch.equal(dsn.cpu(), sn) and torch.equal(da.cpu(), a)
        assert torch.equal(dln.cpu(), ln) and torch.equal(dnl.cpu(), nl)
        assert torch.equal(dr.cpu(), r.to(torch.float32)) and torch.equal(de.cpu().bool(), e.bool())
        rng.replay(); want = t_host.train_step(s, a, r, sn, ln, nl, e)
        rng.replay(); got = t_dev.train_step_async(ds, da, dr, dsn, dln, dnl, de).cpu().tolist()
        rng.advance()
        assert list(want) == got
        n_batches += 1
    assert n_batches == 6  # 5 full batches + the ragged one


@pytest.mark.parametrize("pad_pos", ["end", "beg"])
def test_replay_rows_built_on_device_equal_the_reference_preprocessing(pkg, pad_pos):
    """rec_build_replay_rows vs (a) the golden output of the real preprocess_train_data_incl_act_rew and
    (b) the oracle on a larger seeded log (sessions of length 1 .. 3 L); integer work: bit-exact."""
    import os
    from oracle.preprocess import build_replay_rows
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess_rr.npz"))
    L, pad = int(g["state_len"]), int(g["pad_id"])
    kw = dict(hidden_dim=64, embedding_dim=64, train_pad_embed=True, use_packed_seq=True, learning_rate=0.01,
              item_num=pad, state_size=L, action_dim=pad, gamma=0.5, gru_layers=1)
    t = pkg.SQN_trainer(device=DEV, **kw); t.send_to_device()
    eng = t._ready(32)
    buf = pkg.DeviceReplayBuffer.from_event_log(eng, g["session_id"], g["item_id"], pad, pad_pos)
    for col, key in (("states", "state"), ("next_states", "next_state"), ("actions", "action"),
                     ("true_state_len", "true_state_len"), ("true_next_state_len", "true_next_state_len"), ("is_end", "is_end")):
        assert np.array_equal(getattr(buf, col), g[f"{pad_pos}_{key}"]), col
    rng = np.random.default_rng(9)
    lens = rng.integers(1, 3 * L + 1, size=500)
    sid = np.repeat(np.arange(500), lens)
    items = rng.integers(0, pad, size=int(lens.sum()))
    rew = rng.random(len(items)).astype(np.float32)
    want = build_replay_rows(sid, items, L, pad, pad_pos, rewards=rew)
    buf = pkg.DeviceReplayBuffer.from_event_log(eng, sid, items, pad, pad_pos, rewards=rew)
    assert len(buf) == len(items)
    for col, key in (("states", "state"), ("next_states", "next_state"), ("actions", "action"), ("reward", "r_act"),
                     ("true_state_len", "true_state_len"), ("true_next_state_len", "true_next_state_len"), ("is_end", "is_end")):
        assert np.array_equal(getattr(buf, col), want[key]), col
    # and it feeds the trainer directly
    b = next(buf.batches(eng, 32, shuffle=True, generator=torch.Generator().manual_seed(1)))
    losses = t.train_step_async(*b).cpu()
    assert torch.isfinite(losses).all()


def test_native_training_loop_matches_reference_style_loop(pkg, tmp_path):
    """train_native (device buffer, device-side metric accumulation, evaluation points, best-model
    checkpoint) against the reference-style loop of trainSQN.py:168-428 written with the drop-in pieces
    (DataLoader -> train_step -> update_train_metrics -> evaluate): same losses, metrics and coverage."""
    from torch.utils.data import DataLoader
    from ikea_recommender_system_b200.recommenders.ikea.training.native_loop import eval_points
    V, L, B, VB = 300, 8, 32, 64
    rows = _syn().make_replay_rows(5 * B + 9, V, L, seed=21)
    vrows = _syn().make_replay_rows(150, V, L, seed=22)
    unpop = _syn().unpopular_set_from_actions(rows["action"])
    torch.manual_seed(4)
    e_div = torch.nn.Embedding.from_pretrained(torch.randn(V + 1, 16), freeze=True)
    arrays = dict(states=rows["state"], actions=rows["action"], reward=rows["r_act"], next_states=rows["next_state"],
                  true_state_len=rows["true_state_len"], true_next_state_len=rows["true_next_state_len"],
                  is_end=rows["is_end"])
    kw = dict(hidden_dim=64, embedding_dim=64, train_pad_embed=True, use_packed_seq=True, learning_rate=0.01,
              item_num=V, state_size=L, action_dim=V, gamma=0.5, gru_layers=1)
    mk = dict(padding_pos="end", diversity_embedding=e_div, unpopular_actions_set=unpop)
    ks = dict(head_idx=0, topk_hr_ndcg=[5, 10, 20], topk_to_consider_div=1, topk_to_consider_nov=1,
              topk_to_consider_cov=[1, 5, 10], novelty_rew_signal=1)
    t_nat = pkg.SQN_trainer(device=DEV, **kw)
    t_ref = pkg.SQN_trainer(device=DEV, **kw)
    t_ref.DQN_1.load_state_dict(t_nat.DQN_1.state_dict()); t_ref.DQN_2.load_state_dict(t_nat.DQN_2.state_dict())
    buf = pkg.DeviceReplayBuffer.from_arrays(**arrays).to_device(DEV)
    val = pkg.DeviceEvaluationDataset(arrays=dict(states=vrows["state"], actions=vrows["action"],
                                                  true_state_len=vrows["true_state_len"])).to_device(DEV)
    eval_at = (0.5, 1.0)
    rng = synced_random()
    rng.replay()
    hist = pkg.train_native(t_nat, buf, val, epochs=1, batch_size=B, val_batch_size=VB, eval_at=eval_at,
                            generator=torch.Generator().manual_seed(3), out_dir=str(tmp_path), **mk)
    # reference-style loop with the drop-in pieces
    rng.replay()
    t_ref.send_to_device()
    loader = DataLoader(buf, batch_size=B, shuffle=True, generator=torch.Generator().manual_seed(3))
    points = eval_points(len(loader), eval_at)
    val_loader = [(torch.from_numpy(vrows["state"][lo:lo + VB]), torch.from_numpy(vrows["action"][lo:lo + VB]),
                   torch.from_numpy(vrows["true_state_len"][lo:lo + VB])) for lo in range(0, 150, VB)]
    tot = np.zeros(2); hr = np.zeros(3); nd = np.zeros(3); reps = np.zeros(3); n = 0; div = 0.0; nov = 0.0; nb = 0
    cov = {k: set() for k in (1, 5, 10)}
    want = []
    for i, (s, a, r, sn, ln, nl, e) in enumerate(loader):
        t_ref.set_train()
        sup, q = t_ref.train_step(s, a, r, sn, ln, nl, e)
        tot += (sup, q)
        h, d, cov, dv, nv, rp = pkg.update_train_metrics(s=s, a=a, s_len=ln, model=t_ref.DQN_1, device=DEV,
                                                         actions_covered_topk_dict=cov, **mk, **ks)
        hr += h; nd += d; reps += rp; n += len(a); div += float(dv); nov += float(nv)
        nb += 1
        if i + 1 in points:
            v1 = pkg.evaluate(val_loader, t_ref.DQN_1, DEV, t_ref.cross_entropy_loss, **mk, **ks)
            v2 = pkg.evaluate(val_loader, t_ref.DQN_2, DEV, t_ref.cross_entropy_loss, **mk, **ks)
            want.append(dict(sup=tot[0] / nb, q=tot[1] / nb, hr=hr / n, ndcg=nd / n, reps=reps / n, div=div / n,
                             nov=nov / n, cov={k: len(c) / V for k, c in cov.items()}, v1=v1, v2=v2))
            # the reference restarts every train-side sum after an evaluation point (trainSQN.py:417-428)
            tot = np.zeros(2); hr = np.zeros(3); nd = np.zeros(3); reps = np.zeros(3); n = 0; div = 0.0; nov = 0.0; nb = 0
            cov = {k: set() for k in (1, 5, 10)}
    assert len(hist) == len(want) == 2
    for got, w in zip(hist, want):
        assert_close([got["train_sup_loss"], got["train_q_loss"]], [w["sup"], w["q"]], rtol=1e-6, atol=1e-7, what="train losses")
        assert np.allclose(got["train_hr"], w["hr"]) and np.allclose(got["train_ndcg"], w["ndcg"])
        assert np.allclose(got["train_reps"], w["reps"])
        assert abs(got["train_div_rew"] - w["div"]) < 1e-5 and abs(got["train_nov_rew"] - w["nov"]) < 1e-6
        for k in (1, 5, 10):
            assert abs(got["train_cov"][k][1] - w["cov"][k]) < 1e-12
        for sfx, v in (("", w["v1"]), ("_2", w["v2"])):
            assert abs(got[f"val_loss{sfx}"] - float(v[0])) < 1e-6
            assert np.array_equal(got[f"val_hr{sfx}"], v[1]) and np.array_equal(got[f"val_ndcg{sfx}"], v[2])
            assert got[f"val_cov{sfx}"] == v[3]
        logs = got["logs"]
        assert logs["Val_HR@20"] == float(w["v1"][1][2]) and logs["Sec_Val_HR@20"] == float(w["v2"][1][2])
        assert logs["Supervised Train Loss"] == got["train_sup_loss"] and "Q-Modification-Signal" in logs
        assert got["best_model_idx"] == (2 if logs["Val_HR@20"] < logs["Sec_Val_HR@20"] else 1)
    assert [h["log_counter"] for h in hist] == [0, 1]
    best_seen = max(max(h["logs"]["Val_HR@20"], h["logs"]["Sec_Val_HR@20"]) for h in hist)
    path = os.path.join(str(tmp_path), "best_model.pt")
    if best_seen > 0:  # SaveBestModel only saves improvements over its initial 0
        ck = torch.load(path)
        assert set(ck) == {"epoch", "model_idx", "hidden_dim", "item_num", "action_dim", "state_size", "embedding_dim",
                           "model_state_dict"}
        assert "embedding.weight" in ck["model_state_dict"] and ck["epoch"] in (0, 1) and ck["model_idx"] in (1, 2)
    else:
        assert not os.path.exists(path)
