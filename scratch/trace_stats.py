import sys, torch, ctypes
sys.path.insert(0, '/root/repo')
import b200pkg; pkg = b200pkg.load()
import bench
from ikea_recommender_system_b200 import synthetic
from ikea_recommender_system_b200.recommenders.evaluate.eval_protocol import _opts
from ikea_recommender_system_b200.engine import EvalAccumulators
dev = torch.device('cuda:0')
N, B, L = 1_000_000, 5000, 10
rows = synthetic.make_replay_rows_fast(B, N, L, seed=7)
unpop = synthetic.unpopular_set_from_actions(rows["action"])
e_div = torch.nn.Embedding.from_pretrained(torch.randn(N + 1, 64), freeze=True)
net = pkg.SQN_Network(hidden_dim=64, item_num=N, state_size=L, action_dim=N, gamma=0.5, gru_layers=1, embedding_dim=64, use_packed_seq=True).to(dev)
s_, a_, _, _, ln_, _, _ = synthetic.as_torch_batch(rows, 0, B)
eng = net._ready(B)
o, kmax, keep = _opts(net, dev, 0, [5, 10, 20], 1, 1, [1, 5, 10, 20], 1, "end", e_div, unpop, None, None)
ds, dl = net._dev_inputs(s_, ln_); da = a_.to(dev)
acc = EvalAccumulators(dev, N)
eng.eval_batch(0, eng._batch(B, ds, da, dl), o, acc.struct); torch.cuda.synchronize()
buf = torch.zeros(240, dtype=torch.int64, device=dev)
eng.lib.rec_debug_set_trace(eng.handle, ctypes.c_void_p(buf.data_ptr()))
eng.eval_batch(0, eng._batch(B, ds, da, dl), o, acc.struct); torch.cuda.synchronize()
eng.lib.rec_debug_set_trace(eng.handle, None)
v = buf.cpu().tolist(); prev = None
for i in range(0, 240, 2):
    tag, clk = v[i], v[i+1]
    if tag == 0: break
    if tag == 1: print()
    print(tag, clk - (prev if prev else clk), end=" | ")
    prev = clk
print()
