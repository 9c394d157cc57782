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


def run(args, wl, metric, make_data, trainer_kwargs, algorithmic_bytes, peaks, ClockSampler):
    import b200pkg
    pkg = b200pkg.load()
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    K, W, B = args.steps, args.warmup, wl["batch"]
    n_b = min(K + W, 128)
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
                           "parallelism": f"vocab-sharded heads x{world}, embedding+GRU replicated, trunk "
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
                "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    dist.barrier()
    trainer.release_graphs()
    dist.destroy_process_group()
