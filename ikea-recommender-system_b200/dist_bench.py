"""N > 1 arm of bench.py: vocabulary-sharded SMORL train step, one process per GPU (torchrun).

Weak scaling: every rank contributes B local sessions per step (global batch G*B); heads are sharded
1/G per rank, embedding + GRU replicated.  Timing: CUDA events on each rank's stream between barriers,
MAX over ranks; value = G*B*K / t.
"""

from __future__ import annotations

import json
import os
import time

import torch
import torch.distributed as dist


def _replicated_params_identical(trainer, dev):
    """Embedding + GRU are replicated: after any number of steps every rank must hold the SAME bits."""
    ok = True
    for net in trainer._nets:
        for name, t in net.state_dict().items():
            if "head" in name:
                continue
            lo, hi = t.detach().clone(), t.detach().clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            ok = ok and bool(torch.equal(lo, hi))
    return ok


def _eval_secondary(args, pkg, trainer, wl, eval_kw, eval_metric, synthetic, dev, rank, world):
    """BASELINE configs[4]: full-catalogue evaluation of the vocabulary-sharded trainer through the public evaluate().
    Default: sharded by SESSIONS (every rank scores every world-th batch against an all-gathered copy of the scored head,
    one reduction of the accumulators at the end); REC_EVAL_SHARD=vocab: per batch one all-gather of per-shard candidate
    records, merge + metrics replicated on every rank."""
    N, B, L = wl["item_num"], wl["batch"], wl["L"]
    n_batches = max(8, world)  # the single-GPU evaluation object sweeps 8 batches too
    by_sessions = os.environ.get("REC_EVAL_SHARD", "sessions") != "vocab"
    rows = synthetic.make_replay_rows_fast(n_batches * B, N, L, seed=7)
    unpop = synthetic.unpopular_set_from_actions(rows["action"])
    e_div = trainer.div_embedding
    loader = []
    for i in range(n_batches):
        s_, a_, _, _, ln_, _, _ = synthetic.as_torch_batch(rows, i * B, (i + 1) * B)
        loader.append((s_, a_, ln_))
    ce = torch.nn.CrossEntropyLoss()
    net = trainer._nets[0]
    # grows the engine workspace to B on every rank (sharded by sessions, rank r takes batch r of the warm-up sweep)
    pkg.evaluate(loader[:world], net, dev, ce, "end", e_div, unpop, **eval_kw)
    # three sweeps, each = one evaluate() call timed between barriers (CUDA events and wall clock, MAX over ranks); the
    # reported sweep is the MEDIAN (a single sweep is 5-20 ms: one host hiccup on one of the ranks doubles it)
    sweeps = []
    for _ in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        out = pkg.evaluate(loader, net, dev, ce, "end", e_div, unpop, **eval_kw)
        ev1.record()
        torch.cuda.synchronize()
        w = torch.tensor([time.perf_counter() - t0, ev0.elapsed_time(ev1) / 1e3], device=dev)
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
        sweeps.append([float(w[0]), float(w[1])])
    wall = sorted(sweeps, key=lambda x: x[1])[1]
    hr = torch.tensor([float(x) for x in out[1]], device=dev)
    hr_lo, hr_hi = hr.clone(), hr.clone()
    dist.all_reduce(hr_lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hr_hi, op=dist.ReduceOp.MAX)
    sessions = n_batches * B
    return {"metric": eval_metric, "value": sessions / float(wall[1]), "unit": "sessions/s", "n_gpus": world,
            "steps": n_batches, "ms_per_step": 1e3 * float(wall[1]) / n_batches, "higher_is_better": True,
            "scaling": "strong (one validation set of %d batches for the whole job)" % n_batches,
            "config": {"workload": wl["name"], "parallelism": (
                f"sharded by sessions x{world}: the scored head all-gathered once per sweep (part of the timed region), every "
                "rank scores every world-th batch against the full catalogue, accumulators summed / coverage bitmaps OR-ed once"
                if by_sessions else
                f"vocab-sharded heads x{world}: per batch one all-gather of per-shard records (max, sum-exp, target logit, "
                "fp32-exact top-k candidates), merge + metrics replicated")},
            "e2e": {"value": sessions / float(wall[0]), "unit": "sessions/s", "h2d_bytes_per_step": B * (L + 2) * 8,
                    "d2h_bytes_per_step": 8 * 27 + 4 * 8 * ((N + 31) // 32) // n_batches},
            "sweeps_ms": [round(1e3 * x[1], 3) for x in sweeps], "sweep_reported": "median of 3",
            "ranks_agree_on_metrics": bool(torch.equal(hr_lo, hr_hi)),
            "metrics_sample": {"hr": [float(x) for x in out[1]]}}


def run(args, wl, metric, make_data, trainer_kwargs, algorithmic_bytes, peaks, ClockSampler, eval_wl=None, eval_kw=None,
        eval_metric=None, synthetic=None):
    import b200pkg
    pkg = b200pkg.load()
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    K, W, B = args.steps, args.warmup, wl["batch"]
    n_b = min(K + W, 128 if wl["item_num"] < 500_000 else 32)
    # same catalogue statistics on every rank, different sessions per rank
    batches_all, unpop, e_div = make_data(wl, n_b * world, seed=0)
    batches = batches_all[rank::world][:n_b]
    trainer = pkg.SMORL_trainer(device=dev, **trainer_kwargs(wl, e_div, unpop))
    trainer.shard_vocabulary(rank, world)
    trainer.send_to_device()
    trainer.set_train()
    dev_batches = [tuple(t.to(dev) for t in b) for b in batches]
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()  # before the warm-up: nvidia-smi needs ~100 ms to emit its first sample
    for i in range(max(W, 30)):
        trainer.train_step_async(*dev_batches[i % n_b])
    torch.cuda.synchronize()
    eng = trainer._engine
    m0 = clocks.mark()
    l0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for i in range(K):
        trainer.train_step_async(*dev_batches[(W + i) % n_b])
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = eng.launch_count() - l0
    clk = clocks.stop(m0, max(clocks.mark(), m0 + 1)) if rank == 0 else None
    value = B * world * K / (ms / 1e3)

    # dominant kernel on this rank (its shard of every head)
    eng.enable_kernel_timing(True)
    kms = []
    for i in range(min(K, 20)):
        trainer.train_step_async(*dev_batches[(W + i) % n_b])
        kms.append(eng.last_kernel_ms(0))
    eng.enable_kernel_timing(False)
    head_ms = sum(kms) / len(kms)

    # e2e: host tensors in, python floats out, every step
    for i in range(3):
        trainer.train_step(*batches[i % n_b])
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for i in range(K):
        trainer.train_step(*batches[(W + i) % n_b])
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())

    replicas_ok = _replicated_params_identical(trainer, dev)
    assert replicas_ok, "replicated embedding / GRU parameters diverged across ranks"
    secondary = None
    if eval_wl is not None:
        trainer.release_graphs()  # the evaluation batch re-creates the engine workspace: captured step graphs are stale
        secondary = {"eval": _eval_secondary(args, pkg, trainer, eval_wl, eval_kw, eval_metric, synthetic, dev, rank, world)}
    if rank == 0:
        ab = algorithmic_bytes(wl)
        peak, peak_src = peaks()
        shard_bytes = ab["sup_head"] / world
        achieved = shard_bytes / (head_ms / 1e3) / 1e9
        L = wl["L"]
        line = {"metric": metric, "value": value, "unit": "sessions/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "global_batch": B * world,
                           "parallelism": f"vocab-sharded heads x{world}, "
                                          + ("embedding table row-sharded (token rows all-reduced from their owners every step), GRU replicated, trunk "
                                             if trainer._sharded_step.shard_embedding else "embedding+GRU replicated, trunk ")
                                          + ("data-parallel (6 collectives/step)" if trainer._sharded_step.dp_trunk
                                             else "on the global batch (4 collectives/step)")
                                          + " over NCCL, whole step replayed as one CUDA graph",
                           "l2_policy": "distinct batch every step; twin nets alternate"},
                "clocks": clk,
                "e2e": {"value": B * world * K / e2e_s, "unit": "sessions/s",
                        "h2d_bytes_per_step": int(trainer._stager.h2d_bytes), "d2h_bytes_per_step": 8,
                        "ms_per_step": 1e3 * e2e_s / K},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": "head_bwd_adam_tc2_kernel (supervised head, this rank's vocabulary shard, global batch in chunks of 256)",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": shard_bytes,
                             "kernel_ms": head_ms, "kernel_share_of_step": head_ms / (ms / K)},
                "replicated_params_bit_identical_across_ranks": replicas_ok,
                "cpu_baseline": None}
        emb_bytes = 24 * (wl["item_num"] + 1) * wl["E"]
        emb_sharded = trainer._sharded_step.shard_embedding
        line["roofline"]["step"] = {"algorithmic_bytes_per_rank": (ab["step"] - emb_bytes) / world
                                    + (emb_bytes / world if emb_sharded else emb_bytes),
                                    "note": "heads 1/G per rank, embedding table "
                                            + ("row-sharded 1/G per rank" if emb_sharded else "replicated")}
        line["roofline"]["step"]["achieved"] = line["roofline"]["step"]["algorithmic_bytes_per_rank"] / (ms / K / 1e3) / 1e9
        line["roofline"]["step"]["frac"] = line["roofline"]["step"]["achieved"] / peak
        if secondary is not None:
            line["secondary"] = secondary
        print(json.dumps(line), flush=True)
    dist.barrier()
    trainer.release_graphs()
    dist.destroy_process_group()
