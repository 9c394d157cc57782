"""Kernels of ONE rank of a G-way sharded step with the data-parallel trunk, on a single GPU (collectives
replaced by local replication: values are meaningless, launch shapes and timings are representative)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import b200pkg; pkg = b200pkg.load()
import bench
from ikea_recommender_system_b200.sharded import shard_bounds
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = 256
wl = dict(bench.WORKLOADS['cfg2']); wl['batch'] = B
batches, unpop, e_div = bench._make_data(wl, 6)
dev = torch.device('cuda:0')
t = pkg.SMORL_trainer(device=dev, **bench._trainer_kwargs(wl, e_div, unpop))
lo, hi = shard_bounds(wl['item_num'], 0, G)
for n in t._nets: n.shard_vocabulary(lo, hi)
t.send_to_device()
Bg, L = B * G, wl['L']
eng = t._ready(Bg)
rec = eng.record_floats()
records = torch.empty(Bg, rec, device=dev); gathered = torch.empty(G, Bg, rec, device=dev)
q = torch.zeros(2, Bg, 3, device=dev); dh = torch.empty(Bg, 64, device=dev); losses = torch.zeros(8, device=dev)
nb = eng.dp_packed_bytes(B)
packed = torch.zeros(nb, dtype=torch.uint8, device=dev); gin = torch.zeros(G * nb, dtype=torch.uint8, device=dev)
i64 = dict(dtype=torch.int64, device=dev)
g_s, g_sn = torch.zeros(Bg, L, **i64), torch.zeros(Bg, L, **i64)
g_a, g_ln, g_nl = torch.zeros(Bg, **i64), torch.zeros(Bg, **i64), torch.zeros(Bg, **i64)
g_r = torch.zeros(Bg, device=dev); g_e = torch.zeros(Bg, dtype=torch.uint8, device=dev)
gb = eng._batch(Bg, g_s, g_a, g_ln, g_r, g_sn, g_nl, g_e)
grads = torch.zeros(eng.dp_grad_floats(), device=dev)
dx_send = torch.zeros(B * L * 64, device=dev); dx_all = torch.zeros(G * B * L * 64, device=dev)
def step(b):
    s, a, r, sn, ln, nl, e = [x.to(dev) for x in b]
    keep = (s, a, ln, r.float().contiguous(), sn, nl, e.to(torch.uint8).contiguous())
    lb = eng._batch(B, *keep)
    eng.dp_forward(lb, 0, packed)
    gin.copy_(packed.repeat(G))
    eng.dp_unpack(gin, G, B, gb)
    eng.train_phase_a_heads(gb, t._hp(), 0, records)
    gathered.copy_(records.unsqueeze(0).expand(G, -1, -1))
    eng.train_phase_b(gathered, G, q)
    eng.train_phase_c(q, losses, dh)
    eng.dp_backward(dh, 0, grads, dx_send)
    dx_all.copy_(dx_send.repeat(G))
    eng.dp_apply(grads, dx_all)
    return keep
for i in range(3): k = step(batches[i]); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3): k = step(batches[3 + i])
e1.record(); torch.cuda.synchronize()
print("G", G, "ms/step (no collectives)", e0.elapsed_time(e1) / 3)
