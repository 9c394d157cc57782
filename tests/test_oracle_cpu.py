"""CPU suite: the oracle against the golden fixtures generated from the REAL reference
(oracle/make_golden.py) and against the reference's own unit-test vectors."""
import random

import numpy as np
import pytest
import torch

import oracle
from helpers import load_golden, sd_from_golden, rows_from_golden, assert_close, assert_state_close, synced_random

from b200pkg import load as _load_pkg  # noqa: F401  (package import is exercised in test_host_cpu)


def _batches(rows, B, steps):
    from ikea_recommender_system_b200 import synthetic
    return [synthetic.as_torch_batch(rows, i * B, (i + 1) * B) for i in range(steps)]


def _meta(g):
    m = g["meta"]
    return dict(item_num=int(m[0]), action_dim=int(m[1]), embedding_dim=int(m[2]), hidden_dim=int(m[3]),
                state_size=int(m[4])), int(m[5]), int(m[6]), bool(m[7]), bool(m[8]), int(m[9])


SUP = [("gru4rec_small", "gru4rec"), ("gru4rec_unpacked_frozenpad", "gru4rec"), ("gru4rec_2layer", "gru4rec"),
       ("bidir_small", "bidir"), ("bidir_unpacked", "bidir")]


@pytest.mark.parametrize("name,family", SUP)
def test_oracle_supervised_matches_reference_fixture(name, family, pkg):
    g = load_golden(name)
    cfg, B, steps, packed, train_pad, layers = _meta(g)
    t = oracle.GRUTrainer(family=family, gru_layers=layers, train_pad_embed=train_pad, use_packed_seq=packed,
                          learning_rate=0.01, **cfg)
    # seeded init is bit-identical to the reference's
    assert_state_close(t.gru_model.state_dict(), sd_from_golden(g, "init"), rtol=0, atol=0)
    batches = _batches(rows_from_golden(g), B, steps)
    s, a, _, _, ln, _, _ = batches[0]
    with torch.no_grad():
        assert_close(t.gru_model(s, ln), g["fwd_logits0"], rtol=1e-5, atol=1e-6, what="fwd logits")
    losses = [t.train_step(b[0], b[1], b[4]) for b in batches]
    assert_close(losses, g["losses"], rtol=1e-5, atol=1e-6, what="losses")
    assert_state_close(t.gru_model.state_dict(), sd_from_golden(g, "final"), rtol=1e-3, atol=2e-5)


@pytest.mark.parametrize("name", ["sqn_small", "sqn_unpacked", "sqn_64"])
def test_oracle_sqn_matches_reference_fixture(name, pkg):
    g = load_golden(name)
    cfg, B, steps, packed, train_pad, layers = _meta(g)
    t = oracle.SQNTrainer(train_pad_embed=train_pad, use_packed_seq=packed, learning_rate=0.01, gamma=0.5,
                          gru_layers=layers, **cfg)
    assert_state_close(t.DQN_1.state_dict(), sd_from_golden(g, "init1"), rtol=0, atol=0)
    assert_state_close(t.DQN_2.state_dict(), sd_from_golden(g, "init2"), rtol=0, atol=0)
    losses, mains = [], []
    for b in _batches(rows_from_golden(g), B, steps):
        losses.append(t.train_step(*b))
        mains.append(t.last_main)
    assert mains == list(g["mains"])  # python RNG stream consumed like the reference
    assert_close(losses, g["losses"], rtol=1e-5, atol=1e-6, what="losses")
    assert_state_close(t.DQN_1.state_dict(), sd_from_golden(g, "final1"), rtol=1e-3, atol=2e-5)
    assert_state_close(t.DQN_2.state_dict(), sd_from_golden(g, "final2"), rtol=1e-3, atol=2e-5)


@pytest.mark.parametrize("name", ["sarm_small", "sarm_unpacked", "sarm_64"])
def test_oracle_sarm_matches_reference_fixture(name, pkg):
    """oracle.SARMTrainer against the fixtures written from the REAL SARM_trainer (oracle/make_golden_sarm.py)."""
    g = load_golden(name)
    cfg, B, steps, packed, train_pad, layers = _meta(g)
    t = oracle.SARMTrainer(train_pad_embed=train_pad, use_packed_seq=packed, learning_rate=0.01, gru_layers=layers, **cfg)
    assert_state_close(t.network.state_dict(), sd_from_golden(g, "init"), rtol=0, atol=0)
    losses = [t.train_step(*b) for b in _batches(rows_from_golden(g), B, steps)]
    assert_close(losses, g["losses"], rtol=1e-5, atol=1e-6, what="losses")
    assert_state_close(t.network.state_dict(), sd_from_golden(g, "final"), rtol=1e-3, atol=2e-5)


def test_oracle_smorl_fixture(pkg):
    g = load_golden("smorl_small")
    cfg, B, steps, *_ = _meta(g)
    unpop = set(int(i) for i in g["unpop"])
    t = oracle.SMORLTrainer(padding_pos="end", train_pad_embed=True, use_packed_seq=True, learning_rate=0.01,
                            gamma=0.5, gru_layers=1, q_weights=[1.0, 0.7, 0.4], alpha=0.8,
                            div_embedding=torch.from_numpy(g["e_div"]), unpopular_actions_set=unpop, topk_div=3,
                            topk_nov=2, nov_rew_sig=1.0, **cfg)
    assert_state_close(t.SMORL_1.state_dict(), sd_from_golden(g, "init1"), rtol=0, atol=0)
    batches = _batches(rows_from_golden(g), B, steps)
    s, a, _, _, ln, nln, _ = batches[0]
    with torch.no_grad():  # these four come from the REAL reference net / helpers
        sup, q = t.SMORL_1(s, ln)
        assert_close(sup, g["fwd_sup0"], rtol=1e-5, atol=1e-6)
        assert_close(q, g["fwd_q0"], rtol=1e-5, atol=1e-6)
        assert_close(oracle.diversity_rewards(s, sup, nln, "end", 3, torch.from_numpy(g["e_div"])), g["div_rew0"],
                     rtol=1e-5, atol=1e-6)
        assert np.array_equal(oracle.novelty_rewards(sup, unpop, 1, 2), g["nov_rew0"])
    losses = [t.train_step(*b) for b in batches]
    assert_close(losses, g["losses"], rtol=1e-5, atol=1e-6, what="losses")
    assert_state_close(t.SMORL_2.state_dict(), sd_from_golden(g, "final2"), rtol=1e-3, atol=2e-5)


def test_oracle_eval_matches_reference_fixture(pkg):
    from ikea_recommender_system_b200 import synthetic
    g = load_golden("eval_sqn64")
    net = oracle.make_sqn(hidden_dim=64, embedding_dim=64, item_num=500, state_size=10, action_dim=500,
                          gru_layers=1, use_packed_seq=True)
    net.load_state_dict(sd_from_golden(g, "net"))
    rows = rows_from_golden(g)
    loader = []
    for lo in range(0, 90, 32):
        s, a, _, _, ln, _, _ = synthetic.as_torch_batch(rows, lo, min(lo + 32, 90))
        loader.append((s, a, ln))
    unpop = set(int(i) for i in g["unpop"])
    r = oracle.evaluate(loader, net, torch.nn.CrossEntropyLoss(), "end", torch.from_numpy(g["e_div"]), unpop,
                        head_idx=0, topk_hr_ndcg=[5, 10, 20], topk_to_consider_div=3, topk_to_consider_nov=2,
                        topk_to_consider_cov=[1, 5, 10, 20], novelty_rew_signal=1)
    assert_close(r[0], g["loss"], rtol=1e-5)
    assert np.array_equal(r[1], g["hr"]) and np.allclose(r[2], g["ndcg"]) and np.array_equal(r[6], g["reps"])
    assert np.allclose([r[3][k] for k in sorted(r[3])], g["cov_vals"])
    assert_close(r[4], g["div"], rtol=1e-5)
    assert np.isclose(r[5], g["nov"])


# ---- the reference's own unit-test vectors (test/*.py), replayed through the oracle ---------------
def test_known_answer_hr_ndcg():
    # test/test_evaluation.py:155-267
    a = torch.tensor([9, 0, 2, 1, 1, 1, 9, 0, 1])
    p = torch.tensor([[1, 2, 3, 4, 5, 6, 7, 8, 9, 10], [10, 9, 8, 7, 6, 5, 4, 3, 2, 1], [1, 2, 100, 2, 2, 2, 2, 2, 2, 2],
                      [1, 2, 3, 4, 5, 6, 7, 8, 9, 10], [10, 9, 8, 7, 6, 5, 4, 3, 2, 1], [1, 1.5, 100, 2, 2, 2, 2, 2, 2, 2],
                      [1, 2, 3, 4, 5, 6, 7, 8, 9, 10], [10, 9, 8, 7, 6, 5, 4, 3, 2, 1], [1, 3, 100, 2, 2, 2, 2, 2, 2, 2]],
                     dtype=torch.float32)
    clicks, buys = slice(0, 6), slice(6, 9)
    h, n = oracle.hits_and_ndcg(p[clicks], a[clicks], [1, 2, 10])
    assert np.allclose(h / 6, [3 / 6, 4 / 6, 1])
    assert np.allclose(n / 6, [3 / 6, (3 + 1 / np.log2(3)) / 6, (3 + 1 / np.log2(3) + 2 / np.log2(10)) / 6])
    h, n = oracle.hits_and_ndcg(p[buys], a[buys], [1, 2, 10])
    assert np.allclose(h / 3, [2 / 3, 1, 1])
    assert np.allclose(n / 3, [2 / 3, (2 + 1 / np.log2(3)) / 3, (2 + 1 / np.log2(3)) / 3])


def test_known_answer_coverage():
    # test/test_coverage.py:8-39
    res = oracle.coverage_result({20: {1, 2, 10, 20, 30, 40}}, {1, 2, 3, 4, 5}, 10, [20])
    assert res[20] == (2 / 5, 6 / 10)
    preds = torch.tensor([[10, 9, 8, 7, 6], [0, 10, 9, 8, 7], [1, 9, 8, 7, 11]], dtype=torch.float32)
    d = oracle.coverage_update({1: {0, 101, 102, 103}, 2: {0, 1, 4, 101, 102, 103}}, preds, [1, 2])
    assert d[1] == {0, 1, 4, 101, 102, 103} and d[2] == {0, 1, 4, 2, 101, 102, 103}


def test_known_answer_novelty():
    # test/test_novelty.py:6-23
    preds = torch.tensor([[100, 50, 0, 0, 0], [100, -10, 10, 0, 0]], dtype=torch.float32)
    unpop = {0, 1, 10, 11, 12, 13}
    assert np.array_equal(oracle.novelty_rewards(preds, unpop, reward=2), [2, 2])
    assert np.array_equal(oracle.novelty_rewards(preds, unpop, reward=2, topk=2), [2, 1])


def test_known_answer_repetitions():
    # test/test_repetions.py:6-18 (ties among equal scores resolved lowest-id first)
    s = torch.tensor([[1, 1, 2, 2, 3, 4], [1, 2, 3, 4, 5, 6], [1, 1, 2, 2, 3, 4]])
    preds = torch.tensor([[0, 11, 10, 5, 5], [0, 11, 10, 9, 8], [9, 8, 7, 10, -10]], dtype=torch.float32)
    r = oracle.repetitions(s, preds, [1, 2, 5])
    assert np.allclose(r / 3, [4 / 3, 7 / 3, 16 / 3])


def test_known_answer_last_action():
    # test/test_diversity.py:5-19
    from oracle.evalproto import last_action
    front = torch.tensor([[0, 0, 1, 2, 3, 3], [0, 0, 0, 12, 13, 2]])
    back = torch.tensor([[1, 2, 3, 3, 0, 0], [12, 13, 2, 0, 0, 0]])
    ln = torch.tensor([4, 3])
    assert torch.equal(last_action(back, "end", ln), torch.tensor([3, 2]))
    assert torch.equal(last_action(front, "beg"), torch.tensor([3, 2]))


def test_known_answer_q_helpers():
    # test/test_tensor_operations.py:10-78 restated on the oracle's inline formulas
    q = torch.stack([torch.tensor([[1., 2, 3, 4, 5], [-1, -2, -3, -4, -5]]),
                     torch.tensor([[10., 20, 30, 40, 50], [0.1, 0.2, 0.3, 0.4, 0.5]]),
                     torch.tensor([[100., 200, 300, 400, 500], [11, 21, 31, 41, 51]])], dim=1)
    a = torch.tensor([3, 1])
    got = torch.gather(q, 2, a.view(-1, 1, 1).expand(-1, 3, 1)).squeeze(2)
    assert torch.equal(got, torch.tensor([[4, 40, 400], [-2, 0.2, 21]]))
    w = torch.tensor([0.1, 0.5, 0.4])
    assert torch.allclose(torch.matmul(got, w), torch.tensor([180.4, 8.3]))
    scal = torch.sum(q[:, :, :3] * w.view(1, -1, 1), dim=1)
    assert torch.equal(torch.argmax(scal, dim=1), torch.tensor([2, 2]))


def test_stable_topk_lowest_id_first():
    x = torch.tensor([[3., 3, 1, 3, 2], [0., 0, 0, 0, 0]])
    assert oracle.stable_topk(x, 3).tolist() == [[0, 1, 3], [0, 1, 2]]
    big = torch.zeros(2, 5000)
    big[0, 4000] = 1.0
    assert oracle.stable_topk(big, 4).tolist() == [[4000, 0, 1, 2], [0, 1, 2, 3]]


def test_replay_row_construction_oracle_matches_reference_golden():
    """oracle.preprocess.build_replay_rows vs the output of the real `preprocess_train_data_incl_act_rew`
    (recommenders/data_utils/preprocessing.py:199-320) stored by oracle/make_golden_preprocess.py; plus the edge cases
    of get_state / get_next_state (:5-29, :143-170): singleton sessions, sessions longer than state_len, both paddings."""
    import os
    from oracle.preprocess import build_replay_rows, session_offsets
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess_rr.npz"))
    L, pad = int(g["state_len"]), int(g["pad_id"])
    for pos in ("end", "beg"):
        mine = build_replay_rows(g["session_id"], g["item_id"], L, pad, pos)
        for k in ("state", "next_state", "action", "true_state_len", "true_next_state_len", "is_end"):
            assert np.array_equal(mine[k], g[f"{pos}_{k}"]), (pos, k)
    off = session_offsets([7, 7, 7, 3, 9, 9])
    assert off.tolist() == [0, 3, 4, 6]
    r = build_replay_rows([1, 1, 1, 1, 2], [10, 11, 12, 13, 14], 2, 99, "end")
    assert r["state"].tolist() == [[99, 99], [10, 99], [10, 11], [11, 12], [99, 99]]
    assert r["next_state"].tolist() == [[10, 99], [10, 11], [11, 12], [12, 13], [14, 99]]
    assert r["true_state_len"].tolist() == [1, 1, 2, 2, 1] and r["true_next_state_len"].tolist() == [1, 2, 2, 2, 1]
    assert r["is_end"].tolist() == [False, False, False, True, True]
    rb = build_replay_rows([1, 1, 1], [10, 11, 12], 3, 99, "beg")
    assert rb["state"].tolist() == [[99, 99, 99], [99, 99, 10], [99, 10, 11]]
    assert rb["next_state"].tolist() == [[99, 99, 10], [99, 10, 11], [10, 11, 12]]
