import sys, torch, ctypes, time
sys.path.insert(0, '/root/repo')
import b200pkg; pkg = b200pkg.load()
from ikea_recommender_system_b200 import synthetic, _native as N_
from ikea_recommender_system_b200.engine import EvalAccumulators
dev = torch.device('cuda:0')
N, B, L = 1_000_000, 5000, 10
rows = synthetic.make_replay_rows_fast(B, N, L, seed=7)
net = pkg.SQN_Network(hidden_dim=64, item_num=N, state_size=L, action_dim=N, gamma=0.5, gru_layers=1, embedding_dim=64, use_packed_seq=True).to(dev)
s_, a_, _, _, ln_, _, _ = synthetic.as_torch_batch(rows, 0, B)
eng = net._ready(B)
ds, dl = net._dev_inputs(s_, ln_); da = a_.to(dev)
for K in (1, 4, 8, 9, 20):
    o = N_.RecEvalOpts(); o.head_idx = 0; o.n_k = 1; o.ks[0] = K; o.n_cov = 0
    acc = EvalAccumulators(dev, N)
    eng.eval_batch(0, eng._batch(B, ds, da, dl), o, acc.struct); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): eng.eval_batch(0, eng._batch(B, ds, da, dl), o, acc.struct)
    e1.record(); torch.cuda.synchronize()
    print("K", K, "ms/batch", e0.elapsed_time(e1) / 3)
