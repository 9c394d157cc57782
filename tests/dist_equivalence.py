"""torchrun script (NOT collected by pytest): G-GPU vocabulary-sharded SMORL training must equal the
oracle on the concatenated global batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
        tests/dist_equivalence.py
"""
import os
import random
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("DIST_HANG_DUMP_S", "150")), exit=True)  # a hang prints where
    import b200pkg
    pkg = b200pkg.load()
    import oracle
    from helpers import assert_close
    from ikea_recommender_system_b200 import synthetic
    from ikea_recommender_system_b200.sharded import shard_bounds
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    V, L, B, steps = 5000, 10, 64, int(os.environ.get("DIST_STEPS", "12"))  # > 2 x warm: both twins get captured + replayed
    kw = dict(hidden_dim=64, embedding_dim=64, padding_pos="end", train_pad_embed=True, use_packed_seq=True,
              learning_rate=0.01, item_num=V, state_size=L, action_dim=V, gamma=0.5, gru_layers=1,
              q_weights=torch.tensor([1.0, 0.5, 0.25]), alpha=1.0, topk_div=2, topk_nov=1, nov_rew_sig=1.0)
    rows = synthetic.make_replay_rows(steps * B * world, V, L, seed=3)
    unpop = synthetic.unpopular_set_from_actions(rows["action"])
    torch.manual_seed(9)
    e_div = torch.nn.Embedding.from_pretrained(torch.randn(V + 1, 16), freeze=True)
    t = pkg.SMORL_trainer(div_embedding=e_div, unpopular_actions_set=unpop, device=dev, **kw)
    t.shard_vocabulary(rank, world)
    t.send_to_device()
    ref = oracle.SMORLTrainer(div_embedding=e_div, unpopular_actions_set=unpop, **kw)
    state = random.getstate()
    for i in range(steps):
        g0 = i * B * world
        local_b = synthetic.as_torch_batch(rows, g0 + rank * B, g0 + (rank + 1) * B)
        global_b = synthetic.as_torch_batch(rows, g0, g0 + B * world)
        random.setstate(state); want = ref.train_step(*global_b)
        random.setstate(state); got = t.train_step(*local_b)
        state = random.getstate()
        assert t.last_main == ref.last_main
        assert_close(got, want, rtol=1e-3, atol=1e-5, what=f"rank {rank} step {i} losses")
    lo, hi = shard_bounds(V, rank, world)
    for mine, full in ((t.SMORL_1, ref.SMORL_1), (t.SMORL_2, ref.SMORL_2)):
        sd, fsd = mine.state_dict(), full.state_dict()
        for k in fsd:
            want_k = fsd[k][lo:hi] if "head" in k else fsd[k]
            err = (sd[k].cpu().double() - want_k.double()).abs()
            bad = err > 2e-5 + 1e-3 * want_k.double().abs()
            assert int(bad.sum()) <= max(1, int(1e-3 * bad.numel())) and float(err.max()) <= 0.02 * 0.01 * steps, \
                f"rank {rank} {k}: {int(bad.sum())} off, max {float(err.max()):.2e}"
    # replicated parameters are bit-identical across ranks
    emb = t.SMORL_1.state_dict()["embedding.weight"].clone()
    ref_emb = emb.clone()
    dist.broadcast(ref_emb, src=0)
    assert torch.equal(emb, ref_emb)
    dist.barrier()
    if rank == 0:
        print(f"dist_equivalence ok: world={world}")
    t.release_graphs()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
