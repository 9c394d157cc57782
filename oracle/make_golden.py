"""Pin the oracle to the REAL reference and write the golden fixtures.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference exists):

    python -m oracle.make_golden            # validates + rewrites tests/golden/*.npz

It (1) imports the unmodified reference classes from /root/reference, (2) asserts that every
oracle net / trainer / helper is BIT-IDENTICAL to them on shared seeds and inputs, and
(3) stores inputs + the reference's outputs as fixtures so the GPU box (which has no
/root/reference) can check both the oracle and the CUDA path against the real thing.
Nothing here is imported by the product.
"""

from __future__ import annotations

import io
import os
import random
import sys
from contextlib import redirect_stdout

import numpy as np
import torch

REF_ROOT = os.environ.get("REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLD = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)

import oracle  # noqa: E402
from ikea_recommender_system_b200 import synthetic  # noqa: E402


def _ref():
    sys.path.insert(0, REF_ROOT)
    from recommenders.models.GRU4Rec.model import GRU4Rec_trainer
    from recommenders.models.BidirGRU4Rec.model import BidirGRU4Rec_trainer
    from recommenders.models.SQN.sqn_gru import SQN_trainer
    from recommenders.models.SMORL.smorl_gru import SMORL_GRU_Net
    from recommenders.evaluate import eval_protocol, coverage, diversity, novelty, repetitiveness
    from recommenders.utils import tensor_operations
    return dict(GRU4Rec_trainer=GRU4Rec_trainer, BidirGRU4Rec_trainer=BidirGRU4Rec_trainer,
                SQN_trainer=SQN_trainer, SMORL_GRU_Net=SMORL_GRU_Net, eval_protocol=eval_protocol,
                coverage=coverage, diversity=diversity, novelty=novelty, repetitiveness=repetitiveness,
                tensor_operations=tensor_operations)


def _sd_equal(a, b):
    ka, kb = list(a.keys()), list(b.keys())
    assert ka == kb, (ka, kb)
    for k in ka:
        assert torch.equal(a[k], b[k]), f"state_dict mismatch at {k}"


def _pack_sd(prefix, sd, out):
    for k, v in sd.items():
        out[f"{prefix}/{k}"] = v.detach().cpu().numpy().copy()


CFG_SMALL = dict(item_num=150, action_dim=150, embedding_dim=12, hidden_dim=20, state_size=7)
CFG_64 = dict(item_num=500, action_dim=500, embedding_dim=64, hidden_dim=64, state_size=10)
B_SMALL, STEPS = 24, 4


def _batches(cfg, B, steps, seed):
    rows = synthetic.make_replay_rows(B * steps, cfg["item_num"], cfg["state_size"], seed=seed)
    return rows, [synthetic.as_torch_batch(rows, i * B, (i + 1) * B) for i in range(steps)]


def _store_rows(out, rows):
    for k, v in rows.items():
        out[f"rows/{k}"] = v


def golden_supervised(ref, report, name, family, cfg, packed, train_pad, layers=1):
    kw = dict(hidden_dim=cfg["hidden_dim"], embedding_dim=cfg["embedding_dim"], gru_layers=layers,
              train_pad_embed=train_pad, use_packed_seq=packed, learning_rate=0.01,
              item_num=cfg["item_num"], state_size=cfg["state_size"], action_dim=cfg["action_dim"],
              device="cpu", torch_rand_seed=118, python_rand_seed=999)
    if family == "bidir":
        r = ref["BidirGRU4Rec_trainer"](dropout=0.0, **kw)
    else:
        r = ref["GRU4Rec_trainer"](**kw)
    o = oracle.GRUTrainer(family=family, **kw)
    _sd_equal(r.gru_model.state_dict(), o.gru_model.state_dict())
    out = {}
    _pack_sd("init", r.gru_model.state_dict(), out)
    rows, batches = _batches(cfg, B_SMALL, STEPS, seed=11)
    _store_rows(out, rows)
    with torch.no_grad():
        s, a, _, _, ln, _, _ = batches[0]
        lr_ = r.gru_model(s, ln)
        lo_ = o.gru_model(s, ln)
        assert torch.equal(lr_, lo_)
        out["fwd_logits0"] = lr_.numpy().copy()
    losses = []
    for (s, a, _, _, ln, _, _) in batches:
        l_ref = r.train_step(s, a, ln)
        l_or = o.train_step(s, a, ln)
        assert l_ref == l_or, (name, l_ref, l_or)
        losses.append(l_ref)
    _sd_equal(r.gru_model.state_dict(), o.gru_model.state_dict())
    _pack_sd("final", r.gru_model.state_dict(), out)
    out["losses"] = np.asarray(losses, dtype=np.float64)
    out["meta"] = np.asarray([cfg["item_num"], cfg["action_dim"], cfg["embedding_dim"], cfg["hidden_dim"],
                              cfg["state_size"], B_SMALL, STEPS, int(packed), int(train_pad), layers])
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), **out)
    report.append(f"{name}: reference == oracle bit-exact over {STEPS} steps; losses {losses}")


def golden_sqn(ref, report, name, cfg, packed, train_pad, layers=1):
    kw = dict(hidden_dim=cfg["hidden_dim"], embedding_dim=cfg["embedding_dim"], train_pad_embed=train_pad,
              use_packed_seq=packed, learning_rate=0.01, item_num=cfg["item_num"],
              state_size=cfg["state_size"], action_dim=cfg["action_dim"], gamma=0.5, gru_layers=layers,
              device="cpu", torch_rand_seed=118, python_rand_seed=999)
    r = ref["SQN_trainer"](**kw)
    rs = random.getstate()
    o = oracle.SQNTrainer(**kw)
    _sd_equal(r.DQN_1.state_dict(), o.DQN_1.state_dict())
    _sd_equal(r.DQN_2.state_dict(), o.DQN_2.state_dict())
    out = {}
    _pack_sd("init1", r.DQN_1.state_dict(), out)
    _pack_sd("init2", r.DQN_2.state_dict(), out)
    rows, batches = _batches(cfg, B_SMALL, STEPS, seed=12)
    _store_rows(out, rows)
    losses, mains = [], []
    for bt in batches:
        random.setstate(rs)
        l_ref = r.train_step(*bt)
        random.setstate(rs)
        l_or = o.train_step(*bt)
        rs = random.getstate()
        assert l_ref == l_or, (name, l_ref, l_or)
        losses.append(l_ref)
        mains.append(o.last_main)
    _sd_equal(r.DQN_1.state_dict(), o.DQN_1.state_dict())
    _sd_equal(r.DQN_2.state_dict(), o.DQN_2.state_dict())
    _pack_sd("final1", r.DQN_1.state_dict(), out)
    _pack_sd("final2", r.DQN_2.state_dict(), out)
    out["losses"] = np.asarray(losses, dtype=np.float64)  # [steps, 2] = (sup, q)
    out["mains"] = np.asarray(mains)
    out["meta"] = np.asarray([cfg["item_num"], cfg["action_dim"], cfg["embedding_dim"], cfg["hidden_dim"],
                              cfg["state_size"], B_SMALL, STEPS, int(packed), int(train_pad), layers])
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), **out)
    report.append(f"{name}: reference == oracle bit-exact over {STEPS} steps; mains {mains}; losses {losses}")


def golden_smorl(ref, report):
    """The reference's SMORL train_step raises at HEAD (2 reward columns, 3 heads); pin what CAN
    be pinned: the net (init + forward) and every helper on the path; the joined step is the
    oracle's restatement and the fixture says so."""
    cfg = CFG_SMALL
    torch.manual_seed(118)
    rnet = ref["SMORL_GRU_Net"](hidden_dim=cfg["hidden_dim"], embedding_dim=cfg["embedding_dim"],
                                item_num=cfg["item_num"], state_size=cfg["state_size"],
                                action_dim=cfg["action_dim"], q_weights=torch.tensor([1., 1., 1.]),
                                gamma=0.5, gru_layers=1, use_packed_seq=True)
    torch.manual_seed(118)
    onet = oracle.make_smorl(hidden_dim=cfg["hidden_dim"], embedding_dim=cfg["embedding_dim"],
                             item_num=cfg["item_num"], state_size=cfg["state_size"],
                             action_dim=cfg["action_dim"], gru_layers=1, use_packed_seq=True)
    _sd_equal(rnet.state_dict(), onet.state_dict())
    rows, batches = _batches(cfg, B_SMALL, STEPS, seed=13)
    s, a, r_acc, s_next, ln, nln, is_end = batches[0]
    with torch.no_grad():
        rs_, rq_ = rnet(s, ln)
        os_, oq_ = onet(s, ln)
    assert torch.equal(rs_, os_) and torch.equal(rq_, oq_)
    # helpers: reference vs oracle on identical tensors
    to = ref["tensor_operations"]
    w = torch.tensor([0.2, 0.5, 0.3])
    assert torch.equal(to.get_max_action(to.get_weighted_q_target(rq_, w)),
                       torch.argmax(torch.sum(oq_ * w.view(1, -1, 1), dim=1), dim=1))
    g_ref = to.gather_from_3d(rq_, a, num_heads=3)
    g_or = torch.gather(oq_, 2, a.view(-1, 1, 1).expand(-1, 3, 1)).squeeze(2)
    assert torch.equal(g_ref, g_or)
    torch.manual_seed(5)
    e_div = torch.nn.Embedding.from_pretrained(torch.randn(cfg["item_num"] + 1, 9), freeze=True)
    d_ref = ref["diversity"].get_batch_diversity_rewards(s, rs_, nln, "end", 3, e_div, device="cpu")
    d_or = oracle.diversity_rewards(s, os_, nln, "end", 3, e_div)
    assert torch.equal(d_ref, d_or)
    unpop = synthetic.unpopular_set_from_actions(rows["action"])
    n_ref = ref["novelty"].get_batch_novelty_rewards(rs_, unpop, reward=1, topk_to_consider=2)
    n_or = oracle.novelty_rewards(os_, unpop, 1, 2)
    assert np.array_equal(n_ref, n_or)

    # restated joined step -> fixture (marked restated)
    kw = dict(hidden_dim=cfg["hidden_dim"], embedding_dim=cfg["embedding_dim"], padding_pos="end",
              train_pad_embed=True, use_packed_seq=True, learning_rate=0.01, item_num=cfg["item_num"],
              state_size=cfg["state_size"], action_dim=cfg["action_dim"], gamma=0.5, gru_layers=1,
              q_weights=[1.0, 0.7, 0.4], alpha=0.8, div_embedding=e_div, unpopular_actions_set=unpop,
              topk_div=3, topk_nov=2, nov_rew_sig=1.0, torch_rand_seed=118, python_rand_seed=999)
    o = oracle.SMORLTrainer(**kw)
    out = {}
    _pack_sd("init1", o.SMORL_1.state_dict(), out)
    _pack_sd("init2", o.SMORL_2.state_dict(), out)
    _store_rows(out, rows)
    out["e_div"] = e_div.weight.numpy().copy()
    out["unpop"] = np.asarray(sorted(unpop), dtype=np.int64)
    losses, mains = [], []
    for bt in batches:
        losses.append(o.train_step(*bt))
        mains.append(o.last_main)
    _pack_sd("final1", o.SMORL_1.state_dict(), out)
    _pack_sd("final2", o.SMORL_2.state_dict(), out)
    out["losses"] = np.asarray(losses, dtype=np.float64)
    out["mains"] = np.asarray(mains)
    out["fwd_sup0"] = rs_.numpy().copy()  # from the REAL reference net
    out["fwd_q0"] = rq_.numpy().copy()
    out["div_rew0"] = d_ref.numpy().copy()
    out["nov_rew0"] = np.asarray(n_ref, dtype=np.float64)
    out["meta"] = np.asarray([cfg["item_num"], cfg["action_dim"], cfg["embedding_dim"], cfg["hidden_dim"],
                              cfg["state_size"], B_SMALL, STEPS, 1, 1, 1])
    np.savez_compressed(os.path.join(GOLD, "smorl_small.npz"), **out)
    report.append("smorl_small: net init/forward + gather_from_3d/get_weighted_q_target/get_max_action/"
                  "diversity/novelty == reference bit-exact; joined train_step RESTATED (reference raises); "
                  f"mains {mains}; losses {losses}")


class _Loader(list):
    pass


def golden_eval(ref, report):
    cfg = CFG_64
    torch.manual_seed(7)
    net = oracle.make_sqn(hidden_dim=cfg["hidden_dim"], embedding_dim=cfg["embedding_dim"],
                          item_num=cfg["item_num"], state_size=cfg["state_size"],
                          action_dim=cfg["action_dim"], gru_layers=1, use_packed_seq=True)
    # spread the logits so that top-k is far from ties
    with torch.no_grad():
        net.sup_head_output.weight.mul_(40.0)
        net.embedding.weight.mul_(30.0)
    rows = synthetic.make_replay_rows(90, cfg["item_num"], cfg["state_size"], seed=21)
    loader = _Loader()
    for lo in range(0, 90, 32):
        s, a, _, _, ln, _, _ = synthetic.as_torch_batch(rows, lo, min(lo + 32, 90))
        loader.append((s, a, ln))
    torch.manual_seed(3)
    e_div = torch.nn.Embedding.from_pretrained(torch.randn(cfg["item_num"] + 1, 16), freeze=True)
    unpop = synthetic.unpopular_set_from_actions(rows["action"])
    ce = torch.nn.CrossEntropyLoss()
    kw = dict(head_idx=0, topk_hr_ndcg=[5, 10, 20], topk_to_consider_div=3, topk_to_consider_nov=2,
              topk_to_consider_cov=[1, 5, 10, 20], novelty_rew_signal=1)
    with redirect_stdout(io.StringIO()):
        r = ref["eval_protocol"].evaluate(loader, net, "cpu", ce, "end", e_div, unpop, **kw)
    o = oracle.evaluate(loader, net, ce, "end", e_div, unpop, **kw)
    assert torch.equal(r[0], o[0])
    assert np.array_equal(r[1], o[1]) and np.array_equal(r[2], o[2])
    assert r[3] == o[3]
    assert torch.equal(r[4], o[4]) and r[5] == o[5] and np.array_equal(r[6], o[6])
    out = {}
    _pack_sd("net", net.state_dict(), out)
    _store_rows(out, rows)
    out["e_div"] = e_div.weight.numpy().copy()
    out["unpop"] = np.asarray(sorted(unpop), dtype=np.int64)
    out["loss"] = np.asarray(float(r[0]))
    out["hr"], out["ndcg"], out["reps"] = r[1], r[2], r[6]
    out["cov_keys"] = np.asarray(sorted(r[3].keys()))
    out["cov_vals"] = np.asarray([r[3][k] for k in sorted(r[3].keys())], dtype=np.float64)
    out["div"] = np.asarray(float(r[4]))
    out["nov"] = np.asarray(float(r[5]))
    # update_train_metrics on the first batch
    s, a, ln = loader[0]
    cov0 = {k: set() for k in [1, 5, 10, 20]}
    ru = ref["eval_protocol"].update_train_metrics(s, a, ln, net, "cpu", "end", e_div, unpop, cov0, **kw)
    ou = oracle.update_train_metrics(s, a, ln, net, "end", e_div, unpop, {k: set() for k in [1, 5, 10, 20]}, **kw)
    assert np.array_equal(ru[0], ou[0]) and np.array_equal(ru[1], ou[1]) and ru[2] == ou[2]
    assert torch.equal(ru[3], ou[3]) and ru[4] == ou[4] and np.array_equal(ru[5], ou[5])
    out["utm_hr"], out["utm_ndcg"], out["utm_reps"] = ru[0], ru[1], ru[5]
    out["utm_div"] = np.asarray(float(ru[3]))
    out["utm_nov"] = np.asarray(float(ru[4]))
    np.savez_compressed(os.path.join(GOLD, "eval_sqn64.npz"), **out)
    report.append(f"eval_sqn64: reference.evaluate == oracle.evaluate exactly (loss {float(r[0]):.6f}, "
                  f"hr {r[1]}, ndcg {r[2]}, div {float(r[4]):.6f}, nov {r[5]:.6f}, reps {r[6]}); "
                  "update_train_metrics likewise")


def known_answers(ref, report):
    """The reference's own unit-test vectors, replayed through the oracle (and the reference)."""
    # test/test_evaluation.py:155-267
    a = torch.tensor([9, 0, 2, 1, 1, 1, 9, 0, 1])
    preds = torch.tensor([[1, 2, 3, 4, 5, 6, 7, 8, 9, 10], [10, 9, 8, 7, 6, 5, 4, 3, 2, 1],
                          [1, 2, 100, 2, 2, 2, 2, 2, 2, 2], [1, 2, 3, 4, 5, 6, 7, 8, 9, 10],
                          [10, 9, 8, 7, 6, 5, 4, 3, 2, 1], [1, 1.5, 100, 2, 2, 2, 2, 2, 2, 2],
                          [1, 2, 3, 4, 5, 6, 7, 8, 9, 10], [10, 9, 8, 7, 6, 5, 4, 3, 2, 1],
                          [1, 3, 100, 2, 2, 2, 2, 2, 2, 2]], dtype=torch.float32)
    h_r, n_r = ref["eval_protocol"].get_hits_for_batch(preds, a, top_k=[1, 2, 10])
    h_o, n_o = oracle.hits_and_ndcg(preds, a, [1, 2, 10])
    assert np.array_equal(h_o, [5, 7, 9]) and np.array_equal(h_r, h_o) and np.allclose(n_r, n_o)
    report.append(f"known answers: HR/NDCG vector of test_evaluation.py reproduced: hits {h_o}, ndcg {n_o}")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(1)  # fixed reduction order for the fixtures
    ref = _ref()
    report = [f"torch {torch.__version__}; reference at {REF_ROOT}"]
    golden_supervised(ref, report, "gru4rec_small", "gru4rec", CFG_SMALL, packed=True, train_pad=True)
    golden_supervised(ref, report, "gru4rec_unpacked_frozenpad", "gru4rec", CFG_SMALL, packed=False, train_pad=False)
    golden_supervised(ref, report, "gru4rec_2layer", "gru4rec", CFG_SMALL, packed=True, train_pad=True, layers=2)
    golden_supervised(ref, report, "bidir_small", "bidir", CFG_SMALL, packed=True, train_pad=True)
    golden_supervised(ref, report, "bidir_unpacked", "bidir", CFG_SMALL, packed=False, train_pad=True)
    golden_sqn(ref, report, "sqn_small", CFG_SMALL, packed=True, train_pad=True)
    golden_sqn(ref, report, "sqn_unpacked", CFG_SMALL, packed=False, train_pad=True)
    golden_sqn(ref, report, "sqn_64", CFG_64, packed=True, train_pad=True)
    golden_smorl(ref, report)
    golden_eval(ref, report)
    known_answers(ref, report)
    with open(os.path.join(GOLD, "VALIDATION.txt"), "w") as f:
        f.write("\n".join(report) + "\n")
    print("\n".join(report))


if __name__ == "__main__":
    main()
