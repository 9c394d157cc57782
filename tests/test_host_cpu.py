"""CPU suite for the host side: the C-ABI library loads and exports every symbol the header declares,
fails loudly without a GPU, and the drop-in classes mirror the reference interface."""
import ctypes
import inspect
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "recsys_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rec_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(pkg):
    from ikea_recommender_system_b200 import _native
    declared = _declared_functions()
    assert declared, "no functions parsed from the header"
    assert sorted(_native.SYMBOLS) == declared  # binding table == header
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (rec_[a-z_0-9]+)", out))
    assert set(declared) <= exported, set(declared) - exported
    for name in declared:
        assert hasattr(pkg.LIB, name)


def test_struct_sizes_match_header(pkg):
    """sizeof() of the ctypes mirrors against a C translation unit compiled from the header."""
    from ikea_recommender_system_b200 import _native as N
    import tempfile
    code = '#include <stdio.h>\n#include "recsys_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",' \
           'sizeof(rec_config),sizeof(rec_net_params),sizeof(rec_batch),sizeof(rec_train_hparams),' \
           'sizeof(rec_eval_opts),sizeof(rec_eval_accum));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.c")
        open(src, "w").write(code)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    mine = [ctypes.sizeof(c) for c in (N.RecConfig, N.RecNetParams, N.RecBatch, N.RecTrainHparams, N.RecEvalOpts,
                                       N.RecEvalAccum)]
    assert mine == sizes


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(pkg):
    from ikea_recommender_system_b200 import _native as N
    cfg = N.RecConfig(item_num=10, action_dim=10, embedding_dim=8, hidden_dim=8, state_size=4, bidirectional=0,
                      n_heads=1, n_nets=1, use_packed_seq=1, frozen_pad_row=-1, max_batch=4, vocab_lo=0, vocab_hi=10,
                      max_topk=8)
    h = ctypes.c_void_p()
    rc = pkg.LIB.rec_create(ctypes.byref(cfg), None, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in pkg.LIB.rec_last_error(None) or b"sm_" in pkg.LIB.rec_last_error(None)
    net = pkg.GRU4Rec(hidden_size=8, embedding_dim=8, item_num=10, state_size=4, action_dim=10)
    with pytest.raises(RuntimeError, match="No CPU fallback|no CPU fallback"):
        net(torch.zeros(2, 4, dtype=torch.long), torch.ones(2, dtype=torch.long))
    t = pkg.GRU4Rec_trainer(hidden_dim=8, embedding_dim=8, gru_layers=1, train_pad_embed=True, use_packed_seq=True,
                            learning_rate=0.01, item_num=10, state_size=4, action_dim=10, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        t.train_step(torch.zeros(2, 4, dtype=torch.long), torch.zeros(2, dtype=torch.long), torch.ones(2, dtype=torch.long))


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "ikea-recommender-system_b200")
    for dp, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dp, f)


FAMILIES = [
    ("gru4rec", lambda p: p.GRU4Rec(hidden_size=8, embedding_dim=12, item_num=30, state_size=4, action_dim=30, gru_layers=2)),
    ("bidir", lambda p: p.BidirGRU4Rec(hidden_size=8, embedding_dim=12, item_num=30, state_size=4, action_dim=30)),
    ("sqn", lambda p: p.SQN_Network(hidden_dim=8, item_num=30, state_size=4, action_dim=30, gamma=0.5, gru_layers=1, embedding_dim=12)),
    ("smorl", lambda p: p.SMORL_GRU_Net(hidden_dim=8, embedding_dim=12, item_num=30, state_size=4, action_dim=30, q_weights=[1, 1, 1], gamma=0.5)),
]


@pytest.mark.parametrize("family,ctor", FAMILIES)
def test_state_dict_keys_and_seeded_init_match_oracle(pkg, family, ctor):
    torch.manual_seed(42)
    mine = ctor(pkg)
    torch.manual_seed(42)
    layers = 2 if family == "gru4rec" else 1
    ref = oracle.SessionNet(family=family, hidden_dim=8, embedding_dim=12, item_num=30, state_size=4, action_dim=30,
                            gru_layers=layers, use_packed_seq=mine.use_packed_seq)
    a, b = mine.state_dict(), ref.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k
    for attr in ("hidden_dim", "item_num", "action_dim", "state_size", "embedding_dim"):  # read by SaveBestModel
        assert getattr(mine, attr) == getattr(ref, attr)


def test_trainer_signatures_match_reference(pkg):
    """Constructor / train_step parameter names of the reference classes."""
    want = {
        pkg.GRU4Rec_trainer: ["hidden_dim", "embedding_dim", "gru_layers", "train_pad_embed", "use_packed_seq",
                              "learning_rate", "item_num", "state_size", "action_dim", "device", "padding_idx",
                              "torch_rand_seed", "python_rand_seed"],
        pkg.BidirGRU4Rec_trainer: ["hidden_dim", "embedding_dim", "gru_layers", "dropout", "train_pad_embed",
                                   "use_packed_seq", "learning_rate", "item_num", "state_size", "action_dim", "device",
                                   "padding_idx", "torch_rand_seed", "python_rand_seed"],
        pkg.SQN_trainer: ["hidden_dim", "embedding_dim", "train_pad_embed", "use_packed_seq", "learning_rate",
                          "item_num", "state_size", "action_dim", "gamma", "gru_layers", "device", "padding_idx",
                          "torch_rand_seed", "python_rand_seed", "name_1", "name_2"],
        pkg.SMORL_trainer: ["hidden_dim", "embedding_dim", "padding_pos", "train_pad_embed", "use_packed_seq",
                            "learning_rate", "item_num", "state_size", "action_dim", "gamma", "gru_layers", "q_weights",
                            "alpha", "div_embedding", "unpopular_actions_set", "topk_div", "device", "input_tokenizer",
                            "output_tokenizer", "padding_idx", "torch_rand_seed", "python_rand_seed", "name_1", "name_2"],
    }
    for cls, names in want.items():
        got = [p for p in inspect.signature(cls.__init__).parameters if p != "self"]
        assert got[:len(names)] == names, cls
    assert list(inspect.signature(pkg.SQN_trainer.train_step).parameters)[1:] == \
        ["s", "a", "r", "s_next", "true_len", "true_next_len", "is_end"]
    assert list(inspect.signature(pkg.GRU4Rec_trainer.train_step).parameters)[1:] == ["s", "a", "true_len"]
    ev = list(inspect.signature(pkg.evaluate).parameters)
    assert ev == ["evaluation_data_loader", "model", "device", "loss_function", "padding_pos", "diversity_embedding",
                  "unpopular_actions_set", "head_idx", "topk_hr_ndcg", "topk_to_consider_div", "topk_to_consider_nov",
                  "topk_to_consider_cov", "novelty_rew_signal", "input_tokenizer", "output_tokenizer"]
    for t in (pkg.GRU4Rec_trainer, pkg.SQN_trainer, pkg.SMORL_trainer):
        for m in ("set_train", "set_eval", "send_to_device"):
            assert hasattr(t, m)


def test_batch_stager_layout(pkg):
    from ikea_recommender_system_b200.recommenders._base import BatchStager
    st = BatchStager("cpu", L=5)
    B = 7
    buf = torch.zeros(B * (2 * 5 + 3) * 8 + 5 * B + 16, dtype=torch.uint8)
    s, sn, a, ln, nl, r, e = st._views(buf, B)
    assert s.shape == (B, 5) and sn.shape == (B, 5) and a.shape == (B,) and r.dtype == torch.float32
    s.fill_(1); sn.fill_(2); a.fill_(3); ln.fill_(4); nl.fill_(5); r.fill_(0.5); e.fill_(1)
    s2, sn2, a2, ln2, nl2, r2, e2 = st._np_views(buf.numpy(), B)  # numpy views alias the same bytes
    assert int(s2.sum()) == 35 and int(sn2.sum()) == 70 and int(a2.sum()) == 21 and int(ln2.sum()) == 28
    assert int(nl2.sum()) == 35 and float(r2.sum()) == 3.5 and int(e2.sum()) == 7


def test_synthetic_rows_follow_the_replay_buffer_contract(pkg):
    from ikea_recommender_system_b200 import synthetic
    for maker in (synthetic.make_replay_rows, synthetic.make_replay_rows_fast):
        rows = maker(500, 100, 6, seed=1)
        s, ns, a = rows["state"], rows["next_state"], rows["action"]
        ln, nl = rows["true_state_len"], rows["true_next_state_len"]
        assert s.shape == (500, 6) and ns.shape == (500, 6) and s.dtype == np.int64
        assert ln.min() >= 1 and ln.max() <= 6 and nl.min() >= 1 and nl.max() <= 6
        assert ((s >= 0) & (s <= 100)).all() and (a < 100).all()
        n_real = (s != 100).sum(1)
        assert (np.maximum(n_real, 1) == ln).all()          # "end" padding, len forced to >= 1
        assert ((ns != 100).sum(1) == nl).all()
        last = ns[np.arange(500), nl - 1]
        assert (last == a).all()                             # next_state ends with the action
        assert set(np.unique(rows["r_act"]).tolist()) <= {np.float32(0.2), np.float32(1.0)}
    a1 = synthetic.make_replay_rows(50, 100, 6, seed=1)
    a2 = synthetic.make_replay_rows(50, 100, 6, seed=1)
    assert all(np.array_equal(a1[k], a2[k]) for k in a1)


def _replay_arrays(n=37, L=6, N=50, seed=4):
    from ikea_recommender_system_b200 import synthetic
    rows = synthetic.make_replay_rows(n, N, L, seed=seed)
    return dict(states=rows["state"], actions=rows["action"], reward=rows["r_act"], next_states=rows["next_state"],
                true_state_len=rows["true_state_len"], true_next_state_len=rows["true_next_state_len"],
                is_end=rows["is_end"]), rows


def test_device_replay_buffer_follows_the_reference_dataset_protocol(pkg, tmp_path):
    """Host side of SURVEY 8f N2: same tuple order as ReplayBuffer.__getitem__ (ikea/data_utils/replay_buffer.py:65-74),
    same JSON-lines reader, and an epoch order identical to DataLoader(shuffle=True) with the same generator."""
    import pandas as pd
    from torch.utils.data import DataLoader
    arrays, rows = _replay_arrays()
    buf = pkg.DeviceReplayBuffer.from_arrays(**arrays)
    n = len(buf)
    assert n == len(rows["action"])
    item = buf[5]
    assert len(item) == 7
    for got, key in zip(item, ("state", "action", "r_act", "next_state", "true_state_len", "true_next_state_len", "is_end")):
        assert np.array_equal(np.asarray(got), np.asarray(rows[key][5]))
    # JSON-lines file in the reference's format
    df = pd.DataFrame({k: (list(map(list, v)) if np.asarray(v).ndim == 2 else list(np.asarray(v).tolist())) for k, v in rows.items()})
    path = tmp_path / "replay.jsonl"
    df.to_json(path, orient="records", lines=True)
    buf2 = pkg.DeviceReplayBuffer(str(path))
    for c in pkg.DeviceReplayBuffer.COLUMNS:
        assert np.array_equal(np.asarray(getattr(buf2, c)).astype(np.float32), np.asarray(getattr(buf, c)).astype(np.float32)), c
    # epoch order == RandomSampler's; batch slices == BatchSampler's
    g1, g2 = torch.Generator().manual_seed(11), torch.Generator().manual_seed(11)
    perm = pkg.DeviceReplayBuffer.epoch_permutation(n, True, g1)
    loader = DataLoader(buf, batch_size=8, shuffle=True, generator=g2)
    seen = torch.cat([b[1] for b in loader])  # actions in visiting order
    assert torch.equal(seen, torch.as_tensor(arrays["actions"])[perm])
    torch.manual_seed(77)  # the reference's scripts pass no generator: global RNG
    perm_g = pkg.DeviceReplayBuffer.epoch_permutation(n, True)
    torch.manual_seed(77)
    seen_g = torch.cat([b[1] for b in DataLoader(buf, batch_size=8, shuffle=True)])
    assert torch.equal(seen_g, torch.as_tensor(arrays["actions"])[perm_g])
    assert pkg.DeviceReplayBuffer.batch_bounds(n, 8) == [(lo, min(lo + 8, n)) for lo in range(0, n, 8)]
    assert pkg.DeviceReplayBuffer.batch_bounds(n, 8, drop_last=True)[-1] == (24, 32)
    assert torch.equal(pkg.DeviceReplayBuffer.epoch_permutation(5, False), torch.arange(5))
    # multi-GPU slicing: the ranks' local batches concatenate to the single-process global batches (ragged tail dropped)
    G, Bl = 3, 4
    per_rank = [pkg.DeviceReplayBuffer.batch_bounds(n, Bl, rank=r, world=G) for r in range(G)]
    assert len({len(b) for b in per_rank}) == 1 and len(per_rank[0]) == n // (G * Bl)
    glob = pkg.DeviceReplayBuffer.batch_bounds(n, G * Bl, drop_last=True)
    for step, (lo, hi) in enumerate(glob):
        cat = [i for r in range(G) for i in range(*per_rank[r][step])]
        assert cat == list(range(lo, hi))
    with pytest.raises(ValueError):
        pkg.DeviceReplayBuffer.batch_bounds(n, Bl, rank=3, world=3)
    with pytest.raises(RuntimeError):
        buf.to_device("cpu")
    with pytest.raises(RuntimeError):
        next(buf.batches(None, 8))
    ev = pkg.DeviceEvaluationDataset(arrays=dict(states=arrays["states"], actions=arrays["actions"],
                                                 true_state_len=arrays["true_state_len"]))
    assert len(ev) == n and len(ev[3]) == 3


def test_native_loop_host_pieces(pkg, tmp_path):
    """Host logic of train_native: evaluation points follow trainSQN.py:160-161 and the checkpoint has
    SaveBestModel's keys (utils/save_best_model.py:29-41)."""
    from ikea_recommender_system_b200.recommenders.ikea.training import native_loop as NL
    assert NL.eval_points(100, (0.25, 0.5, 0.75, 1.0)) == [25, 50, 75, 100]
    assert NL.eval_points(7, (0.5, 1.0)) == [3, 7]
    net = pkg.SQN_Network(hidden_dim=8, item_num=20, state_size=4, action_dim=20, gamma=0.5, gru_layers=1, embedding_dim=8,
                          train_pad_embed=True, use_packed_seq=True)
    path = tmp_path / "best_model.pt"
    NL.save_best_checkpoint(str(path), epoch=3, model=net, model_idx=2)
    ck = torch.load(str(path))
    assert set(ck) == {"epoch", "model_idx", "hidden_dim", "item_num", "action_dim", "state_size", "embedding_dim",
                       "model_state_dict"}
    assert ck["epoch"] == 3 and ck["model_idx"] == 2 and ck["item_num"] == 20
    assert set(ck["model_state_dict"]) == set(net.state_dict())
    with pytest.raises(TypeError):
        NL._twins(object())
    # logging keys == the reference's get_logging_dict_train (golden written by oracle/make_golden_logging.py)
    import json
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "logging_train.json")))
    kw = dict(g["inputs"])
    for c in ("train_coverage_res", "val_coverage_res"):
        kw[c] = {int(k): v for k, v in kw[c].items()}
    assert NL.logging_dict_train(**kw, q_included=True, prefix="") == g["first"]
    assert NL.logging_dict_train(**kw, q_included=True, prefix="Sec_") == g["second"]
    assert "Val_HR@20" in g["first"] and "Sec_Val_HR@20" in g["second"] and "Train_HR@20" not in g["second"]


def test_bench_algorithmic_bytes_follow_survey_8d():
    """bench.py's roofline numerators: bytes/step = 24 P + 4 (K_h + 1) D V with P = (N+1)E + 3HE + 3H^2 + 6H + K_h (DV + V)
    (SURVEY 8d); per-kernel figures are the Adam traffic (24 B/param) resp. one weight stream (4 B/param)."""
    import bench
    wl = bench.WORKLOADS["cfg2"]
    V, E, H, Kh = wl["item_num"], wl["E"], wl["H"], 4
    P = (V + 1) * E + (3 * H * E + 3 * H * H + 6 * H) + Kh * (H * V + V)
    ab = bench.algorithmic_bytes(wl)
    assert ab["step"] == 24 * P + 4 * (Kh + 1) * H * V
    assert round(ab["step"] / 1e6) == 642
    assert ab["sup_head"] == 24 * (H + 1) * V
    assert ab["q_heads"] == 3 * 24 * H * V  # the timed adam_stream_kernel launch sweeps the weights only (biases: adam_bias_kernel)
    assert ab["emb_adam"] == 24 * (V + 1) * E
    assert ab["sup_stats"] == 4 * (H + 1) * V and ab["greedy_stats"] == 3 * ab["sup_stats"]
    big = bench.algorithmic_bytes(bench.WORKLOADS["cfg4"])
    assert 9.0e9 < big["step"] < 9.1e9  # 9.06 GB at 1 M items
    # FLOPs/step = 2 B D V (3 + (K_h - 1)) + 5 * 2 B L 3H (E + H) dirs: SURVEY 8d quotes 197 GF (cfg4) and 363 GF (cfg3)
    assert round(bench.algorithmic_flops(bench.WORKLOADS["cfg4"]) / 1e9) == 197
    assert round(bench.algorithmic_flops(bench.WORKLOADS["cfg3"]) / 1e9) == 363
    w4 = dict(bench.WORKLOADS["cfg4"], batch=4096)  # `--batch 4096`: 16x the head FLOPs, the same bytes
    assert bench.algorithmic_flops(w4) == 16 * bench.algorithmic_flops(bench.WORKLOADS["cfg4"])
    assert bench.algorithmic_bytes(w4)["step"] == big["step"]


def test_coverage_from_packed_bitmaps_equals_the_reference_counts(pkg):
    """`_coverage` counts bits of the device bitmaps with population counts on packed words; the result must equal the
    reference's set arithmetic (evaluate/coverage.py:24-53) -- here restated on unpacked bits -- including catalogue
    sizes that are not a multiple of 32 and stray bits beyond the last action."""
    from ikea_recommender_system_b200.recommenders.evaluate import eval_protocol as EP
    for n in (37, 1024, 70852, 1_000_001):
        words = (n + 31) // 32
        rng = np.random.default_rng(n)
        cov = rng.integers(0, 2 ** 32, size=(4, words), dtype=np.uint32)
        unpop_set = set(int(i) for i in np.nonzero(rng.random(n) < 0.3)[0])
        packed = EP._unpopular_bitmap(unpop_set, n, "cpu", packed=True)
        table = EP._unpopular_bitmap(unpop_set, n, "cpu").numpy()
        got = EP._coverage(cov, [1, 5, 10, 20], packed, n, len(unpop_set))
        for i, k in enumerate([1, 5, 10, 20]):
            covered = set(int(j) for j in np.nonzero(np.unpackbits(cov[i].view(np.uint8), bitorder="little")[:n])[0])
            assert got[k] == (len(covered & unpop_set) / len(unpop_set), len(covered) / n), (n, k)
        assert table.sum() == len(unpop_set)


def test_built_library_contains_the_tcgen05_and_tma_kernels():
    """SASS of the in-tree librecsys_b200.so (cuobjdump; no GPU needed): built for sm_100a only, and the kernels DESIGN.md
    section 4 calls tensor-core kernels really issue tcgen05 MMAs (UTCHMMA), read their accumulators back from TMEM (LDTM),
    commit to mbarriers (UTCBAR) and stage operands with TMA bulk copies (UBLKCP) -- a library recompiled around
    mma.sync / plain loads would pass every numerics test and fail here."""
    import re
    import shutil
    import subprocess
    from ikea_recommender_system_b200 import _native as N
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", N.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    archs = set(re.findall(r"arch = (sm_\w+)", sass))
    assert archs == {"sm_100a"}, archs
    per_fn, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            per_fn[name] = {"UTCHMMA": 0, "LDTM": 0, "UTCBAR": 0, "UBLKCP": 0}
        elif name is not None:
            for op in per_fn[name]:
                if op in line:
                    per_fn[name][op] += 1

    def kernels(fragment):
        ks = {k: v for k, v in per_fn.items() if fragment in k}
        assert ks, f"no kernel named *{fragment}* in the library"
        return ks

    # D = 64 head kernels (heads_tc.cu): statistics / greedy action (every instantiation) and both backward instantiations
    for frag in ("head_stats_tc_kernel", "head_bwd_adam_tc2_kernel"):
        for k, ops in kernels(frag).items():
            assert ops["UTCHMMA"] > 0 and ops["LDTM"] > 0 and ops["UTCBAR"] > 0, (k, ops)
            # template <NB, NT, ARG, RING>: the RING instantiations stream the weight tiles through the TMA slot ring
            if frag == "head_bwd_adam_tc2_kernel" or re.search(r"Lb[01]ELb1EE", k):
                assert ops["UBLKCP"] > 0, (k, ops)
    assert any(re.search(r"Lb[01]ELb1EE", k) for k in kernels("head_stats_tc_kernel"))
    assert len(kernels("head_bwd_adam_tc2_kernel")) == 2  # resident (<= 256 sessions) and chunked
    # the K-loop skeleton (tck.cuh): wide heads, evaluation chunk maxima, tensor-core GRU trunk
    # (HeadCmaxPairT<n != 0> are the experiment variants of DESIGN 4.7 -- "no MMAs", "no epilogue", ... -- not product kernels)
    product = ["GruBptt", "GruStep", "4GemmE", "HeadDwAdam", "HeadDhE", "HeadFwdILi0E", "HeadFwdILi1E", "HeadFwdILi2E",
               "HeadTopkILb0E", "HeadTopkILb1E", "HeadCmaxFlat", "HeadCmaxPairTILi0E"]
    tck = kernels("tck_kernel")
    for frag in product:
        ks = {k: v for k, v in tck.items() if frag in k}
        assert len(ks) == 1, (frag, list(ks))
        for k, ops in ks.items():
            assert ops["UTCHMMA"] > 0 and ops["LDTM"] > 0 and ops["UBLKCP"] > 0, (k, ops)
