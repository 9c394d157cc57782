"""Phase trace (clock64) of head_bwd_adam_tc_kernel, CTA 0: python scratch/trace_bwd.py [cfg2|cfg4]"""
import sys, torch, ctypes, collections
sys.path.insert(0, '/root/repo')
import b200pkg; pkg = b200pkg.load()
import bench
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
wl = bench.WORKLOADS[name]
batches, unpop, e_div = bench._make_data(wl, 8)
dev = torch.device('cuda:0')
t = pkg.SMORL_trainer(device=dev, **bench._trainer_kwargs(wl, e_div, unpop)); t.send_to_device()
db = [tuple(x.to(dev) for x in b) for b in batches]
for i in range(4): t.train_step_async(*db[i])
torch.cuda.synchronize()
buf = torch.zeros(240, dtype=torch.int64, device=dev)
eng = t._engine
eng.lib.rec_debug_set_trace(eng.handle, ctypes.c_void_p(buf.data_ptr()))
t.train_step_async(*db[5]); torch.cuda.synchronize()
eng.lib.rec_debug_set_trace(eng.handle, None)
v = buf.cpu().tolist()
prev = None
tot = collections.defaultdict(list)
line = []
for i in range(0, 240, 2):
    tag, clk = v[i], v[i+1]
    if tag == 0: break
    d = clk - (prev if prev else clk)
    tot[tag].append(d)
    line.append(f"{tag}:{d}")
    if tag == 12:
        print(" ".join(line)); line = []
    prev = clk
print(" ".join(line))
print(name, "mean cycles spent reaching each tag (all tiles):")
for k in sorted(tot): print("  tag", k, "n", len(tot[k]), "mean", sum(tot[k]) / len(tot[k]))
