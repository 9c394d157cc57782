"""Device-resident input pipeline at cfg2: rec_gather_batch + train_step_async per step, no host
copies; compared with the host-entry loop over DataLoader batches of the same buffer."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import b200pkg; pkg = b200pkg.load()
import bench
from ikea_recommender_system_b200 import synthetic
from torch.utils.data import DataLoader
wl = bench.WORKLOADS['cfg2']; B = wl['batch']
rows = synthetic.make_replay_rows_fast(400 * B, wl['item_num'], wl['L'], seed=0)
unpop = synthetic.unpopular_set_from_actions(rows['action'])
e_div = torch.randn(wl['item_num'] + 1, 64, generator=torch.Generator().manual_seed(1))
dev = torch.device('cuda:0')
t = pkg.SMORL_trainer(device=dev, **bench._trainer_kwargs(wl, e_div, unpop)); t.send_to_device()
buf = pkg.DeviceReplayBuffer.from_arrays(states=rows['state'], actions=rows['action'], reward=rows['r_act'],
                                         next_states=rows['next_state'], true_state_len=rows['true_state_len'],
                                         true_next_state_len=rows['true_next_state_len'], is_end=rows['is_end']).to_device(dev)
eng = t._ready(B)
for ep in range(2):  # epoch 0 warms up (graph capture)
    torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
    for b in buf.batches(eng, B, shuffle=True, generator=torch.Generator().manual_seed(ep), drop_last=True):
        t.train_step_async(*b); n += 1
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"device pipeline: {n} steps, {dt / n * 1e3:.4f} ms/step, {n * B / dt:.0f} sessions/s (wall clock, incl. shuffle + gather)")
loader = DataLoader(buf, batch_size=B, shuffle=True, drop_last=True)
torch.cuda.synchronize(); t0 = time.perf_counter(); n = 0
for b in loader:
    t.train_step(*b); n += 1
    if n == 100: break
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"DataLoader + host entry: {n} steps, {dt / n * 1e3:.4f} ms/step, {n * B / dt:.0f} sessions/s")
