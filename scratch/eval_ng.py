"""torchrun script: evaluate() of a vocabulary-sharded trainer at 1 M items, sharded by sessions, with per-phase wall clock
(REC_EVAL_TRACE=1) printed by every rank."""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200pkg; pkg = b200pkg.load()
import bench
from ikea_recommender_system_b200 import synthetic
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
wl = bench.EVAL_WORKLOADS["eval"]
N, B, L = wl["item_num"], wl["batch"], wl["L"]
twl = dict(bench.WORKLOADS["cfg4"])
batches, unpop, e_div = bench._make_data(twl, 2)
t = pkg.SMORL_trainer(device=dev, **bench._trainer_kwargs(twl, e_div, unpop))
t.shard_vocabulary(rank, world); t.send_to_device()
nb = max(8, world)
rows = synthetic.make_replay_rows_fast(nb * B, N, L, seed=7)
loader = []
for i in range(nb):
    s_, a_, _, _, ln_, _, _ = synthetic.as_torch_batch(rows, i * B, (i + 1) * B)
    loader.append((s_, a_, ln_))
ce = torch.nn.CrossEntropyLoss()
os.environ["REC_EVAL_TRACE"] = "1"
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    pkg.evaluate(loader[:world] if rep == 0 else loader, t._nets[0], dev, ce, "end", e_div, unpop, **bench.EVAL_KW)
    torch.cuda.synchronize(); print(f"rank {rank} rep {rep} total {1e3 * (time.perf_counter() - t0):.2f} ms", flush=True)
dist.barrier(); dist.destroy_process_group()
