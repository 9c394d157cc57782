"""Python handle around one native engine (one per GPU / process).

PyTorch is plumbing here: it owns the device memory of parameters / optimizer state / batches and
the CUDA stream; every computation is a call through the C ABI (`_native.py`).
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _native as N


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class NetTensors:
    """Flat view of one net's parameters (+ optional Adam state) in the engine's order."""

    def __init__(self, emb, w_ih, w_hh, b_ih, b_hh, head_w, head_b):
        self.emb = emb
        self.w_ih, self.w_hh, self.b_ih, self.b_hh = list(w_ih), list(w_hh), list(b_ih), list(b_hh)
        self.head_w, self.head_b = list(head_w), list(head_b)
        self.m: Optional["NetTensors"] = None
        self.v: Optional["NetTensors"] = None

    def all(self) -> List[torch.Tensor]:
        return [self.emb, *self.w_ih, *self.w_hh, *self.b_ih, *self.b_hh, *self.head_w, *self.head_b]

    def zeros_like(self) -> "NetTensors":
        z = torch.zeros_like
        return NetTensors(z(self.emb), [z(t) for t in self.w_ih], [z(t) for t in self.w_hh],
                          [z(t) for t in self.b_ih], [z(t) for t in self.b_hh],
                          [z(t) for t in self.head_w], [z(t) for t in self.head_b])

    def signature(self):
        sig = tuple(t.data_ptr() for t in self.all())
        if self.m is not None:
            sig += tuple(t.data_ptr() for t in self.m.all()) + tuple(t.data_ptr() for t in self.v.all())
        return sig


class Engine:
    """One native engine: static shapes + workspace.  Grows (re-creates) when a larger batch arrives."""

    def __init__(self, *, item_num, action_dim, embedding_dim, hidden_dim, state_size, bidirectional,
                 n_heads, n_nets, use_packed_seq, frozen_pad_row, device, max_batch=256, max_topk=N.REC_MAX_TOPK,
                 vocab_lo=0, vocab_hi=None):
        self.lib = N.load_library()
        self._host_losses = (C.c_float * 4)()
        self._replayed_launches = 0
        self.timing = False
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"the B200 engine runs on CUDA devices only (got device={device!r}); "
                               "there is no CPU fallback")
        self.cfg = dict(item_num=int(item_num), action_dim=int(action_dim), embedding_dim=int(embedding_dim),
                        hidden_dim=int(hidden_dim), state_size=int(state_size), bidirectional=int(bool(bidirectional)),
                        n_heads=int(n_heads), n_nets=int(n_nets), use_packed_seq=int(bool(use_packed_seq)),
                        frozen_pad_row=int(frozen_pad_row), max_batch=int(max_batch), vocab_lo=int(vocab_lo),
                        vocab_hi=int(action_dim if vocab_hi is None else vocab_hi), max_topk=int(max_topk))
        self.handle = None
        self._bound: Dict[int, tuple] = {}
        self._nets: Dict[int, NetTensors] = {}
        self._adam_steps: Dict[int, int] = {}
        self._keep = []  # keeps ctypes structs / tensors of the last call alive
        self.generation = 0  # bumped whenever the native handle (and with it every workspace pointer) is re-created
        self._tc_on, self._graphs_on = True, None
        self._emb_shard = None  # (row_lo, row_hi) owned by this rank (rec_set_embedding_shard); None: all rows
        self._create()

    # -- lifetime ----------------------------------------------------------------------------
    def _create(self):
        cfg = N.RecConfig(**self.cfg)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            rc = self.lib.rec_create(C.byref(cfg), C.c_void_p(stream), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"rec_create failed (rc={rc}): {N.last_error(self.lib, None)}")
        self.handle = h
        self._stream = stream
        self.generation += 1
        # settings that live in the native handle survive a re-create
        if not self._tc_on:
            self.lib.rec_set_tensor_cores(h, 0)
        if self._graphs_on is not None:
            self.lib.rec_set_cuda_graphs(h, int(self._graphs_on))
        if self.timing:
            self.lib.rec_enable_kernel_timing(h, 1)
        if self._emb_shard is not None:
            self.set_embedding_shard(*self._emb_shard)

    def set_embedding_shard(self, row_lo: int, row_hi: int):
        """This rank owns rows [row_lo, row_hi) of every net's embedding table: its Adam sweep touches no other row."""
        self._emb_shard = (int(row_lo), int(row_hi))
        N.check(self.lib, self.handle, self.lib.rec_set_embedding_shard(self.handle, int(row_lo), int(row_hi)),
                "rec_set_embedding_shard")

    def emb_rows_gather(self, net_id, ids, rows_out):
        """rows_out[i] = table[ids[i]] where this rank owns the row, zeros elsewhere (ids: contiguous int64 on the device)."""
        self.follow_stream()
        N.check(self.lib, self.handle,
                self.lib.rec_emb_rows_gather(self.handle, net_id, _ptr(ids), int(ids.numel()), _ptr(rows_out)), "rec_emb_rows_gather")

    def emb_rows_scatter(self, net_id, ids, rows):
        """table[ids[i]] = rows[i] for the rows this rank does NOT own (refresh from the owners after the all-reduce)."""
        self.follow_stream()
        N.check(self.lib, self.handle,
                self.lib.rec_emb_rows_scatter(self.handle, net_id, _ptr(ids), int(ids.numel()), _ptr(rows)), "rec_emb_rows_scatter")

    def follow_stream(self):
        """Launch on torch's CURRENT stream of the engine's device (called at the top of every compute call: work
        issued under `with torch.cuda.stream(s)` must not race with torch's own copies / reads on `s`)."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if stream != self._stream:
            self.lib.rec_set_stream(self.handle, C.c_void_p(stream))
            self._stream = stream

    def close(self):
        if self.handle is not None:
            self.lib.rec_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def max_batch(self):
        return self.cfg["max_batch"]

    def ensure_batch(self, B: int):
        """Grows the workspace by re-creating the native handle (`generation` changes: anything that cached engine
        pointers -- CUDA graphs captured by ShardedStep -- must be rebuilt; see ShardedStep.step)."""
        self.follow_stream()
        if B <= self.cfg["max_batch"]:
            return
        steps = {i: int(self.lib.rec_get_adam_step(self.handle, i)) for i in self._nets}
        self.close()
        self.cfg["max_batch"] = int(max(B, 2 * self.cfg["max_batch"]))
        self._create()
        self._bound.clear()
        for i, nt in self._nets.items():
            self.bind(i, nt, force=True)
            self.lib.rec_set_adam_step(self.handle, i, steps[i])

    # -- parameters --------------------------------------------------------------------------
    def bind(self, net_id: int, nt: NetTensors, force: bool = False):
        self._nets[net_id] = nt
        sig = nt.signature()
        if not force and self._bound.get(net_id) == sig:
            return
        for t in nt.all():
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError("engine parameters must be contiguous fp32 tensors on the engine's device")
        p = N.RecNetParams()
        p.emb = nt.emb.data_ptr()
        if nt.m is not None:
            p.emb_m, p.emb_v = nt.m.emb.data_ptr(), nt.v.emb.data_ptr()
        for name in ("w_ih", "w_hh", "b_ih", "b_hh", "head_w", "head_b"):
            for i, t in enumerate(getattr(nt, name)):
                getattr(p, name)[i] = t.data_ptr()
                if nt.m is not None:
                    getattr(p, name + "_m")[i] = getattr(nt.m, name)[i].data_ptr()
                    getattr(p, name + "_v")[i] = getattr(nt.v, name)[i].data_ptr()
        N.check(self.lib, self.handle, self.lib.rec_bind_params(self.handle, net_id, C.byref(p)), "rec_bind_params")
        self._bound[net_id] = sig

    def adam_step(self, net_id):
        return int(self.lib.rec_get_adam_step(self.handle, net_id))

    def set_adam_step(self, net_id, step):
        self.lib.rec_set_adam_step(self.handle, net_id, int(step))

    # -- calls -------------------------------------------------------------------------------
    def _batch(self, B, s, a, true_len, r=None, s_next=None, true_next_len=None, is_end=None):
        b = N.RecBatch()
        b.B = int(B)
        b.s, b.a, b.true_len = s.data_ptr(), a.data_ptr(), true_len.data_ptr()
        if r is not None:
            b.r, b.s_next = r.data_ptr(), s_next.data_ptr()
            b.true_next_len, b.is_end = true_next_len.data_ptr(), is_end.data_ptr()
        return b

    def forward_state(self, net_id, s, lengths):
        B = s.shape[0]
        self.ensure_batch(B)
        D = self.cfg["hidden_dim"] * (2 if self.cfg["bidirectional"] else 1)
        h = torch.empty(B, D, device=self.device, dtype=torch.float32)
        N.check(self.lib, self.handle,
                self.lib.rec_forward_state(self.handle, net_id, _ptr(s), _ptr(lengths), B, _ptr(h)),
                "rec_forward_state")
        return h

    def head_logits(self, net_id, head, h):
        self.follow_stream()
        B = h.shape[0]
        V = self.cfg["vocab_hi"] - self.cfg["vocab_lo"]
        out = torch.empty(B, V, device=self.device, dtype=torch.float32)
        N.check(self.lib, self.handle,
                self.lib.rec_head_logits(self.handle, net_id, head, _ptr(h), B, _ptr(out), V), "rec_head_logits")
        return out

    def train_step_supervised(self, batch: N.RecBatch, hp: N.RecTrainHparams, loss_out: torch.Tensor):
        self.ensure_batch(batch.B)
        N.check(self.lib, self.handle,
                self.lib.rec_train_step_supervised(self.handle, C.byref(batch), C.byref(hp), _ptr(loss_out)),
                "rec_train_step_supervised")

    def train_step_q(self, batch: N.RecBatch, hp: N.RecTrainHparams, main_net: int, losses_out: torch.Tensor):
        self.ensure_batch(batch.B)
        N.check(self.lib, self.handle,
                self.lib.rec_train_step_q(self.handle, C.byref(batch), C.byref(hp), main_net, _ptr(losses_out)),
                "rec_train_step_q")

    def train_step_sarm(self, batch: N.RecBatch, hp: N.RecTrainHparams, losses_out: torch.Tensor):
        self.ensure_batch(batch.B)
        N.check(self.lib, self.handle,
                self.lib.rec_train_step_sarm(self.handle, C.byref(batch), C.byref(hp), _ptr(losses_out)),
                "rec_train_step_sarm")

    # host entry points: the batch holds HOST pointers, losses come back as python floats (synchronous)
    def train_step_supervised_host(self, batch: N.RecBatch, hp: N.RecTrainHparams) -> float:
        self.ensure_batch(batch.B)
        N.check(self.lib, self.handle,
                self.lib.rec_train_step_supervised_host(self.handle, C.byref(batch), C.byref(hp), self._host_losses),
                "rec_train_step_supervised_host")
        return self._host_losses[0]

    def train_step_q_host(self, batch: N.RecBatch, hp: N.RecTrainHparams, main_net: int):
        self.ensure_batch(batch.B)
        N.check(self.lib, self.handle,
                self.lib.rec_train_step_q_host(self.handle, C.byref(batch), C.byref(hp), main_net, self._host_losses),
                "rec_train_step_q_host")
        return self._host_losses[0], self._host_losses[1]

    @property
    def host_batch_bytes(self) -> int:
        """Bytes of the H2D copy a host-entry step makes (the engine's whole batch block)."""
        mb, L = self.cfg["max_batch"], self.cfg["state_size"]
        return mb * (2 * L + 3) * 8 + mb * 5

    def eval_hold_params(self, on: bool):
        """Evaluation sweep: parameters are frozen between on=True and on=False (operand images are reused)."""
        N.check(self.lib, self.handle, self.lib.rec_eval_hold_params(self.handle, 1 if on else 0), "rec_eval_hold_params")

    def eval_batch(self, net_id, batch: N.RecBatch, opts: N.RecEvalOpts, acc: N.RecEvalAccum, topk_ids=None,
                   topk_scores=None):
        self.ensure_batch(batch.B)
        N.check(self.lib, self.handle,
                self.lib.rec_eval_batch(self.handle, net_id, C.byref(batch), C.byref(opts), C.byref(acc),
                                        _ptr(topk_ids), _ptr(topk_scores)), "rec_eval_batch")

    # -- vocabulary-sharded (multi-GPU) phases; collectives run by the caller in between --------
    def record_floats(self):
        return int(self.lib.rec_record_floats(self.handle))

    def train_phase_a(self, batch, hp, main_net, records_out):
        self.ensure_batch(batch.B)
        N.check(self.lib, self.handle,
                self.lib.rec_train_phase_a(self.handle, C.byref(batch), C.byref(hp), main_net, _ptr(records_out)),
                "rec_train_phase_a")

    def train_phase_b(self, gathered, n_shards, q_out):
        self.follow_stream()
        N.check(self.lib, self.handle, self.lib.rec_train_phase_b(self.handle, _ptr(gathered), n_shards, _ptr(q_out)),
                "rec_train_phase_b")

    def train_phase_c(self, q_reduced, losses_out, dh_out):
        self.follow_stream()
        N.check(self.lib, self.handle,
                self.lib.rec_train_phase_c(self.handle, _ptr(q_reduced), _ptr(losses_out), _ptr(dh_out)),
                "rec_train_phase_c")

    def train_phase_d(self, dh_reduced):
        self.follow_stream()
        N.check(self.lib, self.handle, self.lib.rec_train_phase_d(self.handle, _ptr(dh_reduced)), "rec_train_phase_d")

    # -- data-parallel trunk of the sharded step (see include/recsys_b200.h) ------------------------
    def dp_packed_bytes(self, B_local):
        return int(self.lib.rec_dp_packed_bytes(self.handle, B_local))

    def dp_grad_floats(self):
        return int(self.lib.rec_dp_grad_floats(self.handle))

    def dp_forward(self, local_batch, main_net, packed_out):
        self.follow_stream()
        N.check(self.lib, self.handle,
                self.lib.rec_dp_forward(self.handle, C.byref(local_batch), main_net, _ptr(packed_out)), "rec_dp_forward")

    def dp_unpack(self, gathered, world, B_local, global_batch):
        self.follow_stream()
        N.check(self.lib, self.handle,
                self.lib.rec_dp_unpack(self.handle, _ptr(gathered), world, B_local, C.byref(global_batch)), "rec_dp_unpack")

    def train_phase_a_heads(self, batch, hp, main_net, records_out):
        self.follow_stream()
        N.check(self.lib, self.handle,
                self.lib.rec_train_phase_a_heads(self.handle, C.byref(batch), C.byref(hp), main_net, _ptr(records_out)),
                "rec_train_phase_a_heads")

    def dp_backward(self, dh_reduced, rank, gru_grads_out, dx_out):
        self.follow_stream()
        N.check(self.lib, self.handle,
                self.lib.rec_dp_backward(self.handle, _ptr(dh_reduced), rank, _ptr(gru_grads_out), _ptr(dx_out)),
                "rec_dp_backward")

    def dp_apply(self, gru_grads_reduced, dx_gathered):
        self.follow_stream()
        N.check(self.lib, self.handle, self.lib.rec_dp_apply(self.handle, _ptr(gru_grads_reduced), _ptr(dx_gathered)),
                "rec_dp_apply")

    def eval_shard_candidates(self, net_id, batch, head_idx, kmax, records_out):
        self.ensure_batch(batch.B)
        N.check(self.lib, self.handle,
                self.lib.rec_eval_shard_candidates(self.handle, net_id, C.byref(batch), head_idx, kmax,
                                                   _ptr(records_out)), "rec_eval_shard_candidates")

    def eval_merge(self, batch, opts, gathered, n_shards, acc, topk_ids=None, topk_scores=None):
        self.follow_stream()
        N.check(self.lib, self.handle,
                self.lib.rec_eval_merge(self.handle, C.byref(batch), C.byref(opts), _ptr(gathered), n_shards,
                                        C.byref(acc), _ptr(topk_ids), _ptr(topk_scores)), "rec_eval_merge")

    def set_tensor_cores(self, on: bool):
        self._tc_on = bool(on)
        self.lib.rec_set_tensor_cores(self.handle, int(on))

    def set_cuda_graphs(self, on: bool):
        """Replay the single-GPU train step as a CUDA graph (default on; REC_NO_GRAPH=1 disables)."""
        self._graphs_on = bool(on)
        self.lib.rec_set_cuda_graphs(self.handle, int(on))

    def set_stream(self, cuda_stream: int):
        """Launch on another stream from now on (used while the sharded step is captured into a torch CUDA graph)."""
        self.lib.rec_set_stream(self.handle, C.c_void_p(cuda_stream))
        self._stream = cuda_stream

    def launch_count(self):
        """Kernels launched by this engine, including those replayed from caller-captured graphs."""
        return int(self.lib.rec_launch_count(self.handle)) + self._replayed_launches

    def enable_kernel_timing(self, on=True):
        """Per-kernel CUDA-event timing: steps run eagerly and serially (no graph replay, no branch overlap)."""
        self.timing = bool(on)
        self.lib.rec_enable_kernel_timing(self.handle, int(on))

    def last_kernel_ms(self, which):
        return float(self.lib.rec_last_kernel_ms(self.handle, which))


class EvalAccumulators:
    """Device accumulators of one evaluation sweep (zeroed at construction)."""

    def __init__(self, device, action_dim):
        self.words = (action_dim + 31) // 32
        self.f64 = torch.zeros(3 * N.REC_MAX_KLIST + 3, dtype=torch.float64, device=device)
        self.cov = torch.zeros(N.REC_MAX_KLIST * self.words, dtype=torch.int32, device=device)
        K = N.REC_MAX_KLIST
        base = self.f64.data_ptr()
        self.struct = N.RecEvalAccum(hits=base, ndcg=base + 8 * K, reps=base + 16 * K, div_sum=base + 24 * K,
                                     nov_sum=base + 24 * K + 8, loss_sum=base + 24 * K + 16,
                                     cov_bits=self.cov.data_ptr())

    def read(self):
        f = self.f64.cpu().numpy()
        K = N.REC_MAX_KLIST
        cov = self.cov.cpu().numpy().view(np.uint32).reshape(K, self.words)
        return dict(hits=f[:K], ndcg=f[K:2 * K], reps=f[2 * K:3 * K], div_sum=f[3 * K], nov_sum=f[3 * K + 1],
                    loss_sum=f[3 * K + 2], cov_bits=cov)
