"""world_size-2 gloo tests (CPU) of the host-side sharding protocol: shard bounds, the packed-row
all-gather, and the record merge (log-sum-exp combine, (score desc, id asc) top-k merge, argmax merge)
checked against the full-vocabulary oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- host-side mirror of the packed-batch all-gather (the product packs on the device: rec_pack_batch) ----------
def pack_rows(s, a, true_len, r=None, s_next=None, true_next_len=None, is_end=None):
    """[B, 2L+4] int64: s | s_next | a | len | next_len | (r bits | is_end << 32)."""
    B, L = s.shape
    out = torch.zeros(B, 2 * L + 4, dtype=torch.int64, device=s.device)
    out[:, :L] = s
    out[:, 2 * L] = a
    out[:, 2 * L + 1] = true_len
    if r is not None:
        out[:, L:2 * L] = s_next
        out[:, 2 * L + 2] = true_next_len
        rbits = r.to(torch.float32).contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        out[:, 2 * L + 3] = rbits | (is_end.to(torch.int64) << 32)
    return out


def unpack_rows(rows, L, with_q=True):
    s = rows[:, :L].contiguous()
    a = rows[:, 2 * L].contiguous()
    ln = rows[:, 2 * L + 1].contiguous()
    if not with_q:
        return s, a, ln, None, None, None, None
    sn = rows[:, L:2 * L].contiguous()
    nl = rows[:, 2 * L + 2].contiguous()
    last = rows[:, 2 * L + 3]
    r = (last & 0xFFFFFFFF).to(torch.int32).view(torch.float32).contiguous()
    e = ((last >> 32) & 1).to(torch.uint8).contiguous()
    return s, a, ln, r, sn, nl, e


def all_gather_rows(local_rows, group=None):
    world = dist.get_world_size(group)
    out = torch.empty(world * local_rows.shape[0], local_rows.shape[1], dtype=local_rows.dtype,
                      device=local_rows.device)
    dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
    return out




def _merge_records(stats, topv, topi, argv, argi, K):
    """Python mirror of head_merge_kernel over the shard axis (axis 0)."""
    m = stats[:, :, 0].max(0).values
    ssum = (stats[:, :, 1] * torch.exp(stats[:, :, 0] - m)).sum(0)
    lse = m + torch.log(ssum)
    tgt = stats[:, :, 2].max(0).values
    G, B, _ = topv.shape
    v = topv.permute(1, 0, 2).reshape(B, -1)
    i = topi.permute(1, 0, 2).reshape(B, -1)
    ids = torch.empty(B, K, dtype=torch.int64)
    for b in range(B):
        order = np.lexsort((i[b].numpy(), -v[b].numpy()))
        ids[b] = i[b][torch.from_numpy(order[:K])]
    av = argv.permute(1, 0)
    ai = argi.permute(1, 0)
    best = torch.empty(B, dtype=torch.int64)
    for b in range(B):
        order = np.lexsort((ai[b].numpy(), -av[b].numpy()))
        best[b] = ai[b][order[0]]
    return lse, tgt, ids, best


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import b200pkg
        b200pkg.load()
        import oracle
        from ikea_recommender_system_b200 import synthetic
        from ikea_recommender_system_b200.sharded import shard_bounds
        V, L, B, K = 301, 6, 10, 7
        lo, hi = shard_bounds(V, rank, world)
        bounds = [shard_bounds(V, g, world) for g in range(world)]
        assert bounds[0][0] == 0 and bounds[-1][1] == V and all(bounds[g][1] == bounds[g + 1][0] for g in range(world - 1))
        # 1) packed-row all-gather round trip: every rank ends with the same global batch
        rows = synthetic.make_replay_rows(B * world, V, L, seed=5)
        local = synthetic.as_torch_batch(rows, rank * B, (rank + 1) * B)
        s, a, r, sn, ln, nl, e = local
        g_rows = all_gather_rows(pack_rows(s, a, ln, r, sn, nl, e))
        gs, ga, gln, gr, gsn, gnl, ge = unpack_rows(g_rows, L)
        full = synthetic.as_torch_batch(rows, 0, B * world)
        assert torch.equal(gs, full[0]) and torch.equal(ga, full[1]) and torch.equal(gr, full[2])
        assert torch.equal(gsn, full[3]) and torch.equal(gln, full[4]) and torch.equal(gnl, full[5])
        assert torch.equal(ge.bool(), full[6])
        # 2) record protocol against the full-vocabulary oracle
        torch.manual_seed(0)
        net = oracle.make_sqn(hidden_dim=8, embedding_dim=8, item_num=V, state_size=L, action_dim=V, gru_layers=1,
                              use_packed_seq=True)
        with torch.no_grad():
            net.embedding.weight.mul_(30)
            net.sup_head_output.weight.mul_(10)
            net.q_head_output.bias.copy_(torch.randint(0, 3, (V,)).float())  # exact ties across shards
            net.q_head_output.weight.zero_()
            sup, q = net(gs, gln)
        loc_sup, loc_q = sup[:, lo:hi], q[:, lo:hi]
        m = loc_sup.max(1).values
        stats = torch.stack([m, torch.exp(loc_sup - m[:, None]).sum(1),
                             torch.where((ga >= lo) & (ga < hi), sup.gather(1, ga[:, None])[:, 0],
                                         torch.full_like(m, -3.4e38))], 1)
        tk = oracle.stable_topk(loc_sup, K)
        topv, topi = loc_sup.gather(1, tk), tk + lo
        am = oracle.stable_topk(loc_q, 1)[:, 0]
        argv, argi = loc_q.gather(1, am[:, None])[:, 0], am + lo

        def gather(t):
            buf = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(buf, t.contiguous())
            return torch.stack(buf)

        lse, tgt, ids, best = _merge_records(gather(stats), gather(topv), gather(topi), gather(argv), gather(argi), K)
        assert torch.allclose(lse, torch.logsumexp(sup, 1), rtol=1e-6, atol=1e-6)
        assert torch.equal(tgt, sup.gather(1, ga[:, None])[:, 0])
        assert torch.equal(ids, oracle.stable_topk(sup, K))
        assert torch.equal(best, oracle.stable_topk(q, 1)[:, 0])  # lowest id among exact ties across shards
        out[rank] = "ok"
    except Exception as ex:  # pragma: no cover
        import traceback
        out[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_sharding_protocol_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out.get(r) == "ok" for r in range(world)), dict(out)


def _replica_worker(rank, world, port, out):
    """Host logic of the session-sharded evaluation (eval_protocol._FullHeadReplica) on CPU tensors over gloo: the
    all-gather of UNEVEN head shards into the full [V, D] / [V] copy, the batch assignment, and the final reduction
    (sum of the float64 accumulators, OR of the coverage bitmaps)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import b200pkg
        pkg = b200pkg.load()
        from ikea_recommender_system_b200.sharded import shard_bounds
        from ikea_recommender_system_b200.engine import EvalAccumulators
        from ikea_recommender_system_b200.recommenders.evaluate.eval_protocol import _FullHeadReplica
        V, D = 1001, 64  # 1001 rows over 3 ranks: 333 / 334 / 334
        t = pkg.SQN_trainer(hidden_dim=D, embedding_dim=D, train_pad_embed=True, use_packed_seq=True, learning_rate=0.01,
                            item_num=V, state_size=5, action_dim=V, gamma=0.5, gru_layers=1, device="cpu")
        full_w = t.DQN_1.sup_head_output.weight.data.clone()
        full_b = t.DQN_1.sup_head_output.bias.data.clone()
        assert _FullHeadReplica.plan(t.DQN_1, 0) is None  # unsharded model: evaluate() runs on the model itself
        t.shard_vocabulary(rank, world)
        lo, hi = shard_bounds(V, rank, world)
        assert t.DQN_1.sup_head_output.weight.shape[0] == hi - lo
        rep = _FullHeadReplica.plan(t.DQN_1, 0)
        assert rep is not None and rep.world == world and rep.rank == rank
        assert torch.equal(rep.W, full_w) and torch.equal(rep.b, full_b)  # seeded init is identical on every rank
        t.DQN_1.sup_head_output.weight.data.add_(1.0 + rank)             # parameters changed since the last sweep
        rep2 = _FullHeadReplica.plan(t.DQN_1, 0)
        assert rep2 is rep                                              # buffers are reused, contents re-gathered
        for g in range(world):
            glo, ghi = shard_bounds(V, g, world)
            assert torch.equal(rep.W[glo:ghi], full_w[glo:ghi] + (1.0 + g))
        assert [rep.mine(i) for i in range(2 * world)] == [i % world == rank for i in range(2 * world)]
        os.environ["REC_EVAL_SHARD"] = "vocab"
        assert _FullHeadReplica.plan(t.DQN_1, 0) is None
        os.environ.pop("REC_EVAL_SHARD")
        acc = EvalAccumulators("cpu", V)
        acc.f64 += float(rank + 1)
        acc.cov[rank] = 1 << rank
        acc.cov[5] = 7 if rank == 0 else 8
        rep.reduce(acc)
        assert torch.all(acc.f64 == float(sum(range(1, world + 1))))
        assert [int(acc.cov[g]) for g in range(world)] == [1 << g for g in range(world)] and int(acc.cov[5]) == 15
        out[rank] = "ok"
    except Exception:  # pragma: no cover
        import traceback
        out[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_session_sharded_evaluation_host_logic_world3_gloo():
    world = 3
    port = 31500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_replica_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out.get(r) == "ok" for r in range(world)), dict(out)
