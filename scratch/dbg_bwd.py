import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import b200pkg; pkg = b200pkg.load()
import oracle
from ikea_recommender_system_b200 import synthetic
DEV='cuda:0'
for (V,B) in [(500,24),(500,130),(3000,96),(3000,300)]:
    kw = dict(hidden_dim=64, embedding_dim=64, gru_layers=1, train_pad_embed=True, use_packed_seq=True, learning_rate=0.01, item_num=V, state_size=10, action_dim=V)
    ref = oracle.GRUTrainer(**kw)
    t = pkg.GRU4Rec_trainer(device=DEV, **kw); t.send_to_device()
    rows = synthetic.make_replay_rows(B, V, 10, seed=15)
    s, a, _, _, ln, _, _ = synthetic.as_torch_batch(rows, 0, B)
    want = ref.train_step(s, a, ln); got = t.train_step(s, a, ln)
    print(V, B, "loss", want, got)
    sd, rsd = t.gru_model.state_dict(), ref.gru_model.state_dict()
    for k in rsd:
        d = (sd[k].cpu() - rsd[k])
        print("   ", k, "nan", int(torch.isnan(sd[k]).sum()), "maxerr", float(d.abs().nan_to_num(1e9).max()))
        if k == 'output.weight':
            bad = torch.isnan(sd[k].cpu()).any(1).nonzero().flatten()
            err_rows = (d.abs().max(1).values > 1e-4).nonzero().flatten()
            print("      nan rows", bad[:10].tolist(), len(bad), "err rows", err_rows[:20].tolist(), len(err_rows))
