import sys, os, random, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import b200pkg; pkg = b200pkg.load()
import oracle
from ikea_recommender_system_b200 import synthetic
from ikea_recommender_system_b200.sharded import shard_bounds
from test_gpu_parity import _virtual_rank_step
DEV='cuda:0'
G, V = 3, 1000
kw = dict(hidden_dim=64, embedding_dim=64, padding_pos="end", train_pad_embed=True, use_packed_seq=True,
          learning_rate=0.01, item_num=V, state_size=10, action_dim=V, gamma=0.5, gru_layers=1,
          q_weights=torch.tensor([1.0, 0.6, 0.3]), alpha=0.9, topk_div=2, topk_nov=1, nov_rew_sig=1.0)
rows = synthetic.make_replay_rows(3 * 96, V, 10, seed=8)
unpop = synthetic.unpopular_set_from_actions(rows["action"])
torch.manual_seed(2)
e_div = torch.nn.Embedding.from_pretrained(torch.randn(V + 1, 16), freeze=True)
ref = oracle.SMORLTrainer(div_embedding=e_div, unpopular_actions_set=unpop, **kw)
full = pkg.SMORL_trainer(div_embedding=e_div, unpopular_actions_set=unpop, device=DEV, **kw); full.send_to_device()
shards = []
for g in range(G):
    t = pkg.SMORL_trainer(div_embedding=e_div, unpopular_actions_set=unpop, device=DEV, **kw)
    lo, hi = shard_bounds(V, g, G)
    for n in t._nets: n.shard_vocabulary(lo, hi)
    t.send_to_device(); shards.append(t)
st = random.getstate()
for i in range(2):
    batch = synthetic.as_torch_batch(rows, i * 96, (i + 1) * 96)
    random.setstate(st); want = ref.train_step(*batch)
    random.setstate(st); got_full = full.train_step(*batch)
    st = random.getstate()
    main = ref.last_main - 1
    got = _virtual_rank_step(shards, lambda t: t._hp(), batch, main)
    print("oracle", want, "full", got_full, "sharded", got)
    import ctypes
    e = full._engine
    # compare rewards buffers: not exposed; skip
