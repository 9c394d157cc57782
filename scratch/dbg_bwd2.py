import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import b200pkg; pkg = b200pkg.load()
from ikea_recommender_system_b200 import synthetic
DEV='cuda:0'
V,B=500,24
kw = dict(hidden_dim=64, embedding_dim=64, gru_layers=1, train_pad_embed=True, use_packed_seq=True, learning_rate=0.01, item_num=V, state_size=10, action_dim=V)
rows = synthetic.make_replay_rows(B, V, 10, seed=15)
s, a, _, _, ln, _, _ = synthetic.as_torch_batch(rows, 0, B)
res = {}
for on in (False, True):
    t = pkg.GRU4Rec_trainer(device=DEV, **kw); t.send_to_device()
    eng = t._ready(B); eng.set_tensor_cores(on)
    loss = t.train_step(s, a, ln)
    torch.cuda.synchronize()
    res[on] = {k: v.cpu().clone() for k, v in t.gru_model.state_dict().items()}
    print("tc", on, "loss", loss)
w0, w1 = res[False]['output.weight'], res[True]['output.weight']
d = (w1 - w0).abs().nan_to_num(9.0)
print("rows with err>1e-4:", (d.max(1).values > 1e-4).nonzero().flatten().tolist())
r = (d.max(1).values > 1e-4).nonzero().flatten()
if len(r):
    i = int(r[0]); print("row", i, "cols bad", (d[i] > 1e-4).nonzero().flatten().tolist()); print(w0[i,:8], w1[i,:8])
    i = int(r[-1]); print("row", i, "cols bad", (d[i] > 1e-4).nonzero().flatten().tolist()); print(w0[i,:8], w1[i,:8])
b0, b1 = res[False]['output.bias'], res[True]['output.bias']
print("bias bad", ((b1-b0).abs().nan_to_num(9.0) > 1e-4).nonzero().flatten().tolist()[:40])
e0, e1 = res[False]['embedding.weight'], res[True]['embedding.weight']
print("emb maxerr", float((e1-e0).abs().nan_to_num(9.0).max()))
