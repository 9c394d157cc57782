"""torchrun script (NOT collected by pytest): G-GPU vocabulary-sharded SMORL training must equal the
oracle on the concatenated global batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tests/dist_equivalence.py

DIST_BACKEND=gloo DIST_ONE_GPU=1: the same ranks as separate processes that SHARE cuda:0 and talk over gloo (NCCL refuses
two ranks on one device).  Everything above the transport is the production path: `ShardedStep`, the phase entry points of
the C ABI, sharded `evaluate()`; the step runs eagerly (a gloo collective cannot be captured into a CUDA graph).
"""
import os
import random
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _stage_collectives_through_host():
    """gloo transport for CUDA tensors: copy to the host, run the collective there, copy back (test harness only)."""
    real = dict(ag=dist.all_gather_into_tensor, ar=dist.all_reduce, bc=dist.broadcast, agl=dist.all_gather)

    def all_gather_into_tensor(out, inp, group=None, **kw):
        o = out.cpu().view(-1); real["ag"](o, inp.cpu().contiguous().view(-1), group=group); out.copy_(o.view_as(out))

    def all_reduce(t, op=dist.ReduceOp.SUM, group=None, **kw):
        c = t.cpu(); real["ar"](c, op=op, group=group); t.copy_(c)

    def broadcast(t, src=0, group=None, **kw):
        c = t.cpu(); real["bc"](c, src=src, group=group); t.copy_(c)

    def all_gather(lst, t, group=None, **kw):
        cl = [x.cpu() for x in lst]; real["agl"](cl, t.cpu(), group=group)
        for x, c in zip(lst, cl):
            x.copy_(c)

    dist.all_gather_into_tensor, dist.all_reduce, dist.broadcast, dist.all_gather = all_gather_into_tensor, all_reduce, broadcast, all_gather


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("DIST_HANG_DUMP_S", "150")), exit=True)  # a hang prints where
    import b200pkg
    pkg = b200pkg.load()
    import oracle
    from helpers import assert_close
    from ikea_recommender_system_b200 import synthetic
    from ikea_recommender_system_b200.sharded import shard_bounds
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = 0 if os.environ.get("DIST_ONE_GPU") else int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("DIST_BACKEND", "nccl") == "gloo":
        os.environ["REC_NO_GRAPH"] = "1"
        dist.init_process_group("gloo")
        _stage_collectives_through_host()
    else:
        dist.init_process_group("nccl", device_id=dev)
    V, L, B, steps = 5000, 10, 64, int(os.environ.get("DIST_STEPS", "12"))  # > 2 x warm: both twins get captured + replayed
    kw = dict(hidden_dim=64, embedding_dim=64, padding_pos="end", train_pad_embed=True, use_packed_seq=True,
              learning_rate=0.01, item_num=V, state_size=L, action_dim=V, gamma=0.5, gru_layers=1,
              q_weights=torch.tensor([1.0, 0.5, 0.25]), alpha=1.0, topk_div=2, topk_nov=1, nov_rew_sig=1.0)
    rows = synthetic.make_replay_rows(steps * B * world, V, L, seed=3)
    unpop = synthetic.unpopular_set_from_actions(rows["action"])
    torch.manual_seed(9)
    e_div = torch.nn.Embedding.from_pretrained(torch.randn(V + 1, 16), freeze=True)
    t = pkg.SMORL_trainer(div_embedding=e_div, unpopular_actions_set=unpop, device=dev, **kw)
    t.shard_vocabulary(rank, world)
    t.send_to_device()
    ref = oracle.SMORLTrainer(div_embedding=e_div, unpopular_actions_set=unpop, **kw)
    state = random.getstate()
    for i in range(steps):
        g0 = i * B * world
        local_b = synthetic.as_torch_batch(rows, g0 + rank * B, g0 + (rank + 1) * B)
        global_b = synthetic.as_torch_batch(rows, g0, g0 + B * world)
        random.setstate(state); want = ref.train_step(*global_b)
        random.setstate(state); got = t.train_step(*local_b)
        state = random.getstate()
        assert t.last_main == ref.last_main
        assert_close(got, want, rtol=1e-3, atol=1e-5, what=f"rank {rank} step {i} losses")
        if os.environ.get("DIST_REGROW") and i == 8:
            # a batch beyond the engine's max_batch re-creates the native handle: the step graphs captured so far
            # point at freed workspace and must be dropped; training continues bit-for-bit afterwards
            nbig = t._engine.max_batch + 64
            big = synthetic.make_replay_rows(nbig, V, L, seed=99)
            sb, _, _, _, lb, _, _ = synthetic.as_torch_batch(big, 0, nbig)
            gen0 = t._engine.generation
            _ = t.SMORL_1.final_state(sb, lb)
            assert t._engine.generation > gen0
    if os.environ.get("REC_SHARD_EMBEDDING") == "1":  # the row-sharded sweep really was in effect
        assert t._sharded_step.shard_embedding and t._engine._emb_shard == shard_bounds(V + 1, rank, world)
        assert t.SMORL_1._emb_stale or t.SMORL_2._emb_stale
    lo, hi = shard_bounds(V, rank, world)
    for mine, full in ((t.SMORL_1, ref.SMORL_1), (t.SMORL_2, ref.SMORL_2)):
        sd, fsd = mine.state_dict(), full.state_dict()
        for k in fsd:
            want_k = fsd[k][lo:hi] if "head" in k else fsd[k]
            err = (sd[k].cpu().double() - want_k.double()).abs()
            bad = err > 2e-5 + 1e-3 * want_k.double().abs()
            assert int(bad.sum()) <= max(1, int(1e-3 * bad.numel())) and float(err.max()) <= 0.02 * 0.01 * steps, \
                f"rank {rank} {k}: {int(bad.sum())} off, max {float(err.max()):.2e}"
    # replicated parameters are bit-identical across ranks
    emb = t.SMORL_1.state_dict()["embedding.weight"].clone()
    ref_emb = emb.clone()
    dist.broadcast(ref_emb, src=0)
    assert torch.equal(emb, ref_emb)
    # sharded evaluation through the public evaluate(): candidates per shard -> all-gather -> merge (replicated result)
    vrows = synthetic.make_replay_rows(300, V, L, seed=17)
    loader = []
    for lo_ in (0, 150):
        s_, a_, _, _, ln_, _, _ = synthetic.as_torch_batch(vrows, lo_, lo_ + 150)
        loader.append((s_, a_, ln_))
    ekw = dict(head_idx=0, topk_hr_ndcg=[5, 10, 20], topk_to_consider_div=2, topk_to_consider_nov=1,
               topk_to_consider_cov=[1, 5, 10, 20], novelty_rew_signal=1)
    got_by_mode = {}
    for mode in ("sessions", "vocab"):  # sharded by sessions (default; rank r scores batch r) / per-batch candidate exchange
        os.environ["REC_EVAL_SHARD"] = mode
        got_by_mode[mode] = pkg.evaluate(loader, t.SMORL_1, dev, torch.nn.CrossEntropyLoss(), "end", e_div, unpop, **ekw)
    os.environ.pop("REC_EVAL_SHARD")
    if world > 1:  # the session-sharded sweep really ran on an all-gathered copy of the scored head
        assert getattr(t.SMORL_1, "_eval_replica", None) is not None and t.SMORL_1._eval_replica.world == world
    # the oracle net with THIS run's trained parameters (gathered from all shards)
    import copy
    onet = copy.deepcopy(ref.SMORL_1)
    assert V % world == 0
    sd = {}
    for k, v in t.SMORL_1.state_dict().items():
        if "head" in k:
            parts = [torch.empty_like(v) for _ in range(world)]
            dist.all_gather(parts, v.contiguous())
            sd[k] = torch.cat(parts).cpu()
        else:
            sd[k] = v.cpu()
    onet.load_state_dict(sd)
    want_e = oracle.evaluate(loader, onet, torch.nn.CrossEntropyLoss(), "end", e_div, unpop, **ekw)
    import numpy as np
    from helpers import topk_margin, double_copy, MARGIN_MIN
    with torch.no_grad():
        o64 = double_copy(onet)
        assert min(topk_margin(o64(s_, ln_)[0], 20) for s_, a_, ln_ in loader) > MARGIN_MIN  # ids are well-posed
    for mode, got_e in got_by_mode.items():
        assert abs(float(got_e[0]) - float(want_e[0])) <= 1e-4 * abs(float(want_e[0])), (mode, got_e[0], want_e[0])
        assert np.array_equal(got_e[1], want_e[1]) and np.allclose(got_e[2], want_e[2], rtol=1e-12), (mode, got_e[1], want_e[1])
        assert got_e[3] == want_e[3] and np.array_equal(got_e[6], want_e[6]), mode
        assert abs(float(got_e[4]) - float(want_e[4])) <= 1e-4 and np.isclose(got_e[5], want_e[5]), mode
    # every rank holds the same reduced result (sum of accumulators, OR of coverage bitmaps)
    mine_hr = torch.tensor(np.concatenate([got_by_mode["sessions"][1], got_by_mode["sessions"][2]]), device=dev)
    lo_hr, hi_hr = mine_hr.clone(), mine_hr.clone()
    dist.all_reduce(lo_hr, op=dist.ReduceOp.MIN); dist.all_reduce(hi_hr, op=dist.ReduceOp.MAX)
    assert torch.equal(lo_hr, hi_hr)
    dist.barrier()
    if rank == 0:
        print(f"dist_equivalence ok: world={world}")
    t.release_graphs()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
