"""Where does the host time of one synchronous train_step go?  (run on the GPU box)"""
import sys, time, cProfile, pstats, io
sys.path.insert(0, ".")
import torch
import b200pkg
pkg = b200pkg.load()
import bench as BN

wl = BN.WORKLOADS["cfg2"]
batches, unpop, e_div = BN._make_data(wl, 64, seed=0)
dev = torch.device("cuda", 0)
tr = pkg.SMORL_trainer(device=dev, **BN._trainer_kwargs(wl, e_div, unpop))
tr.send_to_device(); tr.set_train()
for i in range(50):
    tr.train_step(*batches[i % 64])
torch.cuda.synchronize()
N = 2000
t0 = time.perf_counter()
for i in range(N):
    tr.train_step(*batches[i % 64])
t1 = time.perf_counter()
print("train_step e2e us/step", (t1 - t0) / N * 1e6)
# C call alone
from ikea_recommender_system_b200.recommenders.models import _native_models as M
eng = tr._ready(256)
b = batches[0]
hp = tr._hp()
hs, ha, hl = M._host(b[0], torch.int64), M._host(b[1], torch.int64), M._host(b[4], torch.int64)
hsn, hnl = M._host(b[3], torch.int64), M._host(b[5], torch.int64)
hr = M._host(b[2], torch.float32).reshape(-1); he = M._host(b[6], torch.uint8)
rb = eng._batch(256, hs, ha, hl, hr, hsn, hnl, he)
t0 = time.perf_counter()
for i in range(N):
    eng.train_step_q_host(rb, hp, i & 1)
t1 = time.perf_counter()
print("C call alone us/step", (t1 - t0) / N * 1e6)
pr = cProfile.Profile(); pr.enable()
for i in range(N):
    tr.train_step(*batches[i % 64])
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14); print(s.getvalue()[:3500])
