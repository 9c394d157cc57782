"""cfg3 (BidirGRU4Rec SQN, V=N=250000, L=50, E=H=256, B=256): device-timed train step (generic CUDA-core kernels)."""
import os, sys, torch, time
sys.path.insert(0, '/root/repo')
import b200pkg; pkg = b200pkg.load()
from ikea_recommender_system_b200 import synthetic
V, L, H, B = 250_000, 50, 256, 256
kw = dict(hidden_dim=H, embedding_dim=H, train_pad_embed=True, use_packed_seq=True, learning_rate=0.005,
          item_num=V, state_size=L, action_dim=V, gamma=0.5, gru_layers=1)
dev = torch.device('cuda:0')
t = pkg.SQN_trainer(device=dev, bidirectional=True, **kw); t.send_to_device()
rows = synthetic.make_replay_rows_fast(8 * B, V, L, seed=0)
bs = [tuple(x.to(dev) for x in synthetic.as_torch_batch(rows, i * B, (i + 1) * B)) for i in range(8)]
for i in range(4): t.train_step_async(*bs[i])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = int(os.environ.get('N_STEPS', '10'))
for i in range(n): t.train_step_async(*bs[i % 8])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
P = (V + 1) * H + 2 * (3 * H * H * 2 + 6 * H) + 2 * (2 * H * V + V)
byts = 24 * P + 4 * 3 * 2 * H * V
print(f"cfg3 ms/step {ms:.3f}  sessions/s {B / ms * 1e3:.0f}  algorithmic bytes {byts / 1e9:.2f} GB  HBM-roofline frac {byts / (ms * 1e-3) / 6550.7e9:.3f}")
