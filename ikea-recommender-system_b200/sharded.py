"""Vocabulary-sharded multi-GPU execution (one process per GPU, torch.distributed for the plumbing).

Partitioning: rank g of G owns rows [V*g/G, V*(g+1)/G) of EVERY head (weights,
bias and their Adam state) -- the dominant HBM term (24 B/param/step of dense Adam) scales 1/G -- and, with
`shard_embedding`, rows [(N+1)*g/G, (N+1)*(g+1)/G) of the embedding tables' Adam sweep (every rank keeps a full-size copy
whose foreign rows are refreshed from their owners when a step reads them).  The GRU is replicated and updated
identically on every rank.  Sessions are data-parallel at the input: each rank contributes its local batch, the (tiny)
batches are all-gathered and every rank multiplies the full batch by its vocabulary slice.

Collectives per train step (latency-bound; all <= 1 MB except the row refresh):
    1. all_gather  packed local batches (rec_pack_batch: one byte buffer per rank)
    1a. all_reduce token rows of the step from their owners, [3 * G*B*L, E] (row-sharded embedding sweep only)
    2. all_gather  per-shard head records (max, sum-exp, target logit, argmax, top-k candidates)
    3. all_reduce  Q(s,a) / Q_boot(s',a*) contributions         [2, G*B, 3]
    4. all_reduce  dL/dh partials                                [G*B, D]
(data-parallel trunk, world > 4: + all_reduce GRU gradients, all_gather dx rows, and -- row-sharded sweep -- one 20 KB
all_gather of the token ids ahead of the local GRU passes)
"""

from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _native as N


def shard_bounds(V: int, rank: int, world: int):
    """Balanced contiguous split of [0, V) (the same on every rank, no communication)."""
    return V * rank // world, V * (rank + 1) // world


class ShardedStep:
    """Drives rec_train_phase_a..d of one engine with the three collectives in between."""

    def __init__(self, engine, world, group, D, rank=0, dp_trunk=None, shard_embedding=False):
        self.eng, self.world, self.group, self.D = engine, world, group, D
        # dp_trunk: embedding + GRU run on each rank's own sessions (two more, small, collectives) instead of on
        # the all-gathered global batch on every rank.  Measured on B200 (cfg2, B = 256 per GPU): the replicated
        # trunk wins at 2 GPUs (0.39 vs 0.45 ms), both tie at 4 (0.49), the data-parallel trunk wins at 8
        # (0.61 vs 0.72 ms) -> default: world > 4.  REC_DP_TRUNK=1 / REC_NO_DP_TRUNK=1 force either.
        import os
        self.rank = rank
        if dp_trunk is None:
            dp_trunk = world > 4
        if os.environ.get("REC_DP_TRUNK"):
            dp_trunk = True
        if os.environ.get("REC_NO_DP_TRUNK"):
            dp_trunk = False
        self.dp_trunk = bool(dp_trunk)
        # row-sharded embedding sweep (the engine was told its row range by rec_set_embedding_shard): one more all-reduce
        # per step carries the token rows of the step from their owners (and, with the data-parallel trunk, one more
        # small all-gather the token ids)
        self.shard_embedding = bool(shard_embedding)
        self.rec = engine.record_floats()
        self.cap = 0
        self.B_local = -1
        self._graphs = {}
        self._cap_stream = None
        self.use_graphs = True

    def alloc_inputs(self, B_local, L, dev):
        """Persistent buffers of the input all-gather (packed local batch, gathered bytes, global field arrays)."""
        self.B_local = B_local
        Bg = B_local * self.world
        nbytes = int(self.eng.lib.rec_packed_batch_bytes(self.eng.handle, B_local))
        self.packed = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self.gathered_in = torch.zeros(self.world * nbytes, dtype=torch.uint8, device=dev)
        i64 = dict(dtype=torch.int64, device=dev)
        self.g_s, self.g_sn = torch.zeros(Bg, L, **i64), torch.zeros(Bg, L, **i64)
        self.g_a, self.g_ln, self.g_nl = torch.zeros(Bg, **i64), torch.zeros(Bg, **i64), torch.zeros(Bg, **i64)
        self.g_r = torch.zeros(Bg, dtype=torch.float32, device=dev)
        self.g_e = torch.zeros(Bg, dtype=torch.uint8, device=dev)
        self.global_batch = self.eng._batch(Bg, self.g_s, self.g_a, self.g_ln, self.g_r, self.g_sn, self.g_nl, self.g_e)
        if self.shard_embedding:
            E = self.eng.cfg["embedding_dim"]
            self.emb_rows = torch.zeros(3 * Bg * L, E, dtype=torch.float32, device=dev)  # main(s) | main(s') | boot(s')
            if self.dp_trunk:
                self.ids_send = torch.zeros(2 * B_local * L, **i64)
                self.ids_all = torch.zeros(self.world * 2 * B_local * L, **i64)
                self.ids_s, self.ids_sn = torch.zeros(Bg * L, **i64), torch.zeros(Bg * L, **i64)
        if self.dp_trunk:
            self.local_packed = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            self.l_s, self.l_sn = torch.zeros(B_local, L, **i64), torch.zeros(B_local, L, **i64)
            self.l_a, self.l_ln, self.l_nl = torch.zeros(B_local, **i64), torch.zeros(B_local, **i64), torch.zeros(B_local, **i64)
            self.l_r = torch.zeros(B_local, dtype=torch.float32, device=dev)
            self.l_e = torch.zeros(B_local, dtype=torch.uint8, device=dev)
            self.local_batch = self.eng._batch(B_local, self.l_s, self.l_a, self.l_ln, self.l_r, self.l_sn, self.l_nl, self.l_e)
            self.eng.ensure_batch(Bg)
            cfg = self.eng.cfg
            nb = self.eng.dp_packed_bytes(B_local)
            self.packed = torch.zeros(nb, dtype=torch.uint8, device=dev)
            self.gathered_in = torch.zeros(self.world * nb, dtype=torch.uint8, device=dev)
            f32 = dict(dtype=torch.float32, device=dev)
            self.gru_grads = torch.zeros(self.eng.dp_grad_floats(), **f32)
            n_dx = B_local * L * (2 if cfg["bidirectional"] else 1) * cfg["embedding_dim"]
            self.dx_send = torch.zeros(n_dx, **f32)
            self.dx_all = torch.zeros(self.world * n_dx, **f32)

    def _alloc(self, Bg, dev):
        self.cap = Bg
        self.records = torch.empty(Bg, self.rec, dtype=torch.float32, device=dev)
        self.gathered = torch.empty(self.world, Bg, self.rec, dtype=torch.float32, device=dev)
        self.q = torch.zeros(2, Bg, 3, dtype=torch.float32, device=dev)
        self.dh = torch.empty(Bg, self.D, dtype=torch.float32, device=dev)

    # ---- whole-step graph: input all-gather, unpack, phases A-D and their three collectives --------------
    # The sequence has ~45 launches and 4 NCCL calls for ~0.4 ms of GPU work; issued eagerly it is bound by the
    # host.  After `warm` eager executions of a (twin, kind, lr) key the sequence is captured ONCE into a torch
    # CUDA graph (NCCL collectives are capturable) on a private stream, with the engine switched to that stream
    # (rec_set_stream); every later step is: pack the local batch into the static buffer, replay.
    # REC_NO_GRAPH=1 disables.  All ranks capture at the same step (same keys on every rank by construction).
    def step(self, local_batch, hp, main_net, losses_out, has_q, warm=3):
        import os
        eng = self.eng
        if self.dp_trunk:
            # the local batch pointers vary from step to step: copy it into the static staging fields eagerly
            # (one tiny kernel); everything after that, incl. the local GRU passes, is replayable
            N.check(eng.lib, eng.handle,
                    eng.lib.rec_pack_batch(eng.handle, C.byref(local_batch), C.c_void_p(self.local_packed.data_ptr())),
                    "rec_pack_batch")
        else:
            N.check(eng.lib, eng.handle,
                    eng.lib.rec_pack_batch(eng.handle, C.byref(local_batch), C.c_void_p(self.packed.data_ptr())), "rec_pack_batch")
        if eng.generation != getattr(self, "_eng_generation", None):
            # the engine re-created its native handle (a larger batch arrived): graphs captured before that replay
            # kernels on freed workspace pointers
            self._graphs.clear()
            self._eng_generation = eng.generation
        key = (int(main_net), bool(has_q), float(hp.lr), id(losses_out), float(hp.dropout_p), int(hp.dropout_seed),
               int(hp.dropout_mask or 0))  # everything of `hp` that a captured sequence bakes in
        ent = self._graphs.setdefault(key, {"seen": 0, "graph": None, "launches": 0})
        if os.environ.get("REC_NO_GRAPH") or not self.use_graphs or eng.timing:
            return self._sequence(hp, main_net, losses_out, has_q)
        if ent["graph"] is None:
            if ent["seen"] < warm:
                ent["seen"] += 1
                return self._sequence(hp, main_net, losses_out, has_q)
            user_stream = torch.cuda.current_stream()
            if self._cap_stream is None:
                self._cap_stream = torch.cuda.Stream()
            g = torch.cuda.CUDAGraph()
            l0 = int(eng.lib.rec_launch_count(eng.handle))
            try:
                with torch.cuda.graph(g, stream=self._cap_stream):
                    eng.set_stream(torch.cuda.current_stream().cuda_stream)
                    self._sequence(hp, main_net, losses_out, has_q)
            finally:
                eng.set_stream(user_stream.cuda_stream)
            ent["graph"] = g
            ent["launches"] = int(eng.lib.rec_launch_count(eng.handle)) - l0
            eng._replayed_launches -= ent["launches"]  # nothing ran during the capture itself
        ent["graph"].replay()
        eng._replayed_launches += ent["launches"]

    def release_graphs(self):
        """Drop the captured graphs.  Must run before dist.destroy_process_group(): tearing down an NCCL
        communicator that live CUDA graphs still reference hangs."""
        import gc
        torch.cuda.synchronize()
        self._graphs.clear()
        gc.collect()

    def _refresh_rows(self, main_net, has_q, ids_s, ids_sn):
        """Row-sharded embedding table: the token rows this step reads -- main net on s (and s'), bootstrap net on s' --
        travel from their owners into this rank's copy: owned rows gathered (zeros elsewhere) -> ONE all-reduce(sum)
        (bit-exact: a single owner per row) -> rows this rank does not own written into its table."""
        eng, n = self.eng, ids_s.numel()
        nets_ids = [(main_net, ids_s)] + ([(main_net, ids_sn), (1 - main_net, ids_sn)] if has_q else [])
        for k, (net, ids) in enumerate(nets_ids):
            eng.emb_rows_gather(net, ids, self.emb_rows[k * n:(k + 1) * n])
        dist.all_reduce(self.emb_rows[:len(nets_ids) * n].view(-1), group=self.group)
        for k, (net, ids) in enumerate(nets_ids):
            eng.emb_rows_scatter(net, ids, self.emb_rows[k * n:(k + 1) * n])

    def _sequence(self, hp, main_net, losses_out, has_q):
        if self.dp_trunk:
            return self._sequence_dp(hp, main_net, losses_out, has_q)
        eng = self.eng
        dist.all_gather_into_tensor(self.gathered_in, self.packed, group=self.group)
        gb = self.global_batch
        N.check(eng.lib, eng.handle,
                eng.lib.rec_unpack_batch(eng.handle, C.c_void_p(self.gathered_in.data_ptr()), self.world, self.B_local,
                                         C.byref(gb)), "rec_unpack_batch")
        if self.shard_embedding:
            self._refresh_rows(main_net, has_q, self.g_s.view(-1), self.g_sn.view(-1))
        if not has_q:
            gb = eng._batch(self.B_local * self.world, self.g_s, self.g_a, self.g_ln)
        self.run(gb, hp, main_net, losses_out, has_q)

    def _sequence_dp(self, hp, main_net, losses_out, has_q):
        """Data-parallel trunk: 6 collectives, all small; per-rank work independent of the number of GPUs
        except for the (global-batch x local-vocabulary) head kernels."""
        eng, world, B = self.eng, self.world, self.B_local
        Bg = B * world
        # static local fields <- the packed staging record (world = 1 unpack)
        N.check(eng.lib, eng.handle,
                eng.lib.rec_unpack_batch(eng.handle, C.c_void_p(self.local_packed.data_ptr()), 1, B, C.byref(self.local_batch)),
                "rec_unpack_batch")
        lb = self.local_batch if has_q else eng._batch(B, self.l_s, self.l_a, self.l_ln)
        if self.shard_embedding:
            # the owners need every rank's token ids before the local GRU passes: one more (20 KB per rank) all-gather
            nl = self.l_s.numel()
            self.ids_send[:nl].copy_(self.l_s.view(-1))
            self.ids_send[nl:].copy_(self.l_sn.view(-1))
            dist.all_gather_into_tensor(self.ids_all, self.ids_send, group=self.group)
            v = self.ids_all.view(world, 2, nl)
            self.ids_s.view(world, nl).copy_(v[:, 0])
            self.ids_sn.view(world, nl).copy_(v[:, 1])
            self._refresh_rows(main_net, has_q, self.ids_s, self.ids_sn)
        eng.dp_forward(lb, main_net, self.packed)
        dist.all_gather_into_tensor(self.gathered_in, self.packed, group=self.group)
        eng.dp_unpack(self.gathered_in, world, B, self.global_batch)
        gb = self.global_batch if has_q else eng._batch(Bg, self.g_s, self.g_a, self.g_ln)
        dev = losses_out.device
        if Bg != self.cap:
            self._alloc(Bg, dev)
        eng.train_phase_a_heads(gb, hp, main_net, self.records)
        dist.all_gather_into_tensor(self.gathered, self.records, group=self.group)
        eng.train_phase_b(self.gathered, world, self.q)
        if has_q:
            dist.all_reduce(self.q, group=self.group)
        eng.train_phase_c(self.q, losses_out, self.dh)
        dist.all_reduce(self.dh, group=self.group)
        eng.dp_backward(self.dh, self.rank, self.gru_grads, self.dx_send)
        dist.all_reduce(self.gru_grads, group=self.group)
        dist.all_gather_into_tensor(self.dx_all, self.dx_send, group=self.group)
        eng.dp_apply(self.gru_grads, self.dx_all)

    def run(self, batch, hp, main_net, losses_out, has_q):
        Bg = batch.B
        dev = losses_out.device
        if Bg != self.cap:
            self._alloc(Bg, dev)
        eng = self.eng
        eng.train_phase_a(batch, hp, main_net, self.records)
        dist.all_gather_into_tensor(self.gathered, self.records, group=self.group)
        eng.train_phase_b(self.gathered, self.world, self.q)
        if has_q:
            dist.all_reduce(self.q, group=self.group)
        eng.train_phase_c(self.q, losses_out, self.dh)
        dist.all_reduce(self.dh, group=self.group)
        eng.train_phase_d(self.dh)
