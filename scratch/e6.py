"""pair kernel at training shapes: B = 256 sessions against 1 M items (chunk path forced by REC_CHUNK_MIN_B)"""
import torch, b200pkg, time
pkg = b200pkg.load()
from ikea_recommender_system_b200 import synthetic, _native as N_
from ikea_recommender_system_b200.engine import EvalAccumulators
dev = torch.device("cuda:0")
for V, B in ((1_000_000, 256), (70852, 256), (1_000_000, 512)):
    torch.manual_seed(0)
    net = pkg.SQN_Network(hidden_dim=64, item_num=V, state_size=10, action_dim=V, gamma=0.5, gru_layers=1, embedding_dim=64, use_packed_seq=True)
    net.to(dev)
    rows = synthetic.make_replay_rows_fast(B, V, 10, seed=7)
    s, a, _, _, ln, _, _ = synthetic.as_torch_batch(rows, 0, B)
    eng = net._ready(B)
    o = N_.RecEvalOpts(); o.head_idx, o.n_k, o.n_cov = 0, 1, 0; o.ks[0] = 20
    acc = EvalAccumulators(dev, V)
    ds, dl = net._dev_inputs(s, ln)
    da = a.to(dev)
    eng.enable_kernel_timing(True)
    eng.eval_hold_params(True)
    for _ in range(5):
        eng.eval_batch(net._net_id, eng._batch(B, ds, da, dl), o, acc.struct)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(20):
        eng.eval_batch(net._net_id, eng._batch(B, ds, da, dl), o, acc.struct)
    ev1.record(); torch.cuda.synchronize()
    print(f"V={V} B={B}: head kernel {eng.last_kernel_ms(1)*1e3:.1f} us, whole eval batch {ev0.elapsed_time(ev1)/20*1e3:.1f} us")
    eng.eval_hold_params(False)
    del net, eng
