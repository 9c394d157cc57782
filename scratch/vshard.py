import sys, torch, ctypes, time
sys.path.insert(0, '/root/repo')
import b200pkg; pkg = b200pkg.load()
import bench
from ikea_recommender_system_b200.sharded import shard_bounds
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4
wl = dict(bench.WORKLOADS['cfg2']); wl['batch'] = 256 * G
batches, unpop, e_div = bench._make_data(wl, 6)
dev = torch.device('cuda:0')
kw = bench._trainer_kwargs(wl, e_div, unpop)
t = pkg.SMORL_trainer(device=dev, **kw)
lo, hi = shard_bounds(wl['item_num'], 0, G)
for n in t._nets: n.shard_vocabulary(lo, hi)
t.send_to_device()
Bg = wl['batch']
eng = t._ready(Bg)
rec = eng.record_floats()
records = torch.empty(Bg, rec, device=dev); gathered = torch.empty(G, Bg, rec, device=dev)
q = torch.zeros(2, Bg, 3, device=dev); dh = torch.empty(Bg, 64, device=dev); losses = torch.zeros(8, device=dev)
def step(b):
    s, a, r, sn, ln, nl, e = [x.to(dev) for x in b]
    keep = (s, a, ln, r.float().contiguous(), sn, nl, e.to(torch.uint8).contiguous())
    batch = eng._batch(Bg, keep[0], keep[1], keep[2], keep[3], keep[4], keep[5], keep[6])
    eng.train_phase_a(batch, t._hp(), 0, records)
    gathered.copy_(records.unsqueeze(0).expand(G, -1, -1))
    eng.train_phase_b(gathered, G, q)
    eng.train_phase_c(q, losses, dh)
    eng.train_phase_d(dh)
    return keep
for i in range(3): k = step(batches[i]); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3): k = step(batches[3 + i])
e1.record(); torch.cuda.synchronize()
print("G", G, "ms/step (no collectives)", e0.elapsed_time(e1) / 3)
