"""clock64 stamps of the forward GRU step kernel (CTA 0,0,0) at cfg3: REC_TRACE_SEL=3 python scratch/trace_gru.py"""
import os, sys, torch, ctypes
os.environ["REC_TRACE_SEL"] = "3"
sys.path.insert(0, '/root/repo')
import b200pkg; pkg = b200pkg.load()
from ikea_recommender_system_b200 import synthetic
V, L, H, B = 250_000, 50, 256, 256
kw = dict(hidden_dim=H, embedding_dim=H, train_pad_embed=True, use_packed_seq=True, learning_rate=0.005,
          item_num=V, state_size=L, action_dim=V, gamma=0.5, gru_layers=1)
dev = torch.device('cuda:0')
t = pkg.SQN_trainer(device=dev, bidirectional=True, **kw); t.send_to_device()
rows = synthetic.make_replay_rows_fast(4 * B, V, L, seed=0)
bs = [tuple(x.to(dev) for x in synthetic.as_torch_batch(rows, i * B, (i + 1) * B)) for i in range(4)]
for i in range(3): t.train_step_async(*bs[i])
torch.cuda.synchronize()
buf = torch.zeros(8 * L, dtype=torch.int64, device=dev)
eng = t._engine
eng.lib.rec_debug_set_trace(eng.handle, ctypes.c_void_p(buf.data_ptr()))
t.train_step_async(*bs[3]); torch.cuda.synchronize()
eng.lib.rec_debug_set_trace(eng.handle, None)
v = buf.cpu().view(L, 8).tolist()
names = ["ld_in", "ld_out", "mma0", "mmaL", "epi_in", "epi_out"]
base = v[0][0]
for u in range(0, L):
    r = v[u]
    print(u, " ".join(f"{n}={r[k]-base}" for k, n in enumerate(names)), " step_total=", (v[u][0] - v[u-1][0]) if u else 0)
