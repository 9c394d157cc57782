"""Is the evaluation sweep's time data-dependent?  Train the cfg4 SMORL trainer on ONE GPU for a few hundred steps at batch
B (argv[1]), then time evaluate() over 1 M items per batch (REC_EVAL_TRACE) and the kernels of one batch (timeline)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200pkg; pkg = b200pkg.load()
import bench
from ikea_recommender_system_b200 import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 243
dev = torch.device("cuda", 0)
wl = dict(bench.WORKLOADS["cfg4"]); wl["batch"] = B
batches, unpop, e_div = bench._make_data(wl, 8)
t = pkg.SMORL_trainer(device=dev, **bench._trainer_kwargs(wl, e_div, unpop))
t.send_to_device()
ewl = bench.EVAL_WORKLOADS["eval"]
N, EB, L = ewl["item_num"], ewl["batch"], ewl["L"]
rows = synthetic.make_replay_rows_fast(4 * EB, N, L, seed=7)
loader = []
for i in range(4):
    s_, a_, _, _, ln_, _, _ = synthetic.as_torch_batch(rows, i * EB, (i + 1) * EB)
    loader.append((s_, a_, ln_))
ce = torch.nn.CrossEntropyLoss()
os.environ["REC_EVAL_TRACE"] = "1"
def ev(tag):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = pkg.evaluate(loader, t._nets[0], dev, ce, "end", e_div, unpop, **bench.EVAL_KW)
        torch.cuda.synchronize(); print(tag, f"rep {rep} total {1e3 * (time.perf_counter() - t0):.2f} ms  hr {out[1]} loss {float(out[0]):.4f}", flush=True)
ev("untrained")
done = 0
for chunk in (60, steps - 60):
    for i in range(chunk):
        t.train_step(*batches[(done + i) % len(batches)])
    done += chunk
    ev(f"after {done} steps at B={B}")
eng = t._nets[0]._ready(EB)
eng.enable_kernel_timing(True)
