"""Shared host logic of the drop-in model / trainer classes.

The classes in `recommenders/models/*` keep the reference package's names, constructor arguments,
attributes and `state_dict()` keys; none of them computes anything in PyTorch.  The nn.Embedding /
nn.GRU / nn.Linear sub-modules exist only as *parameter containers* (same names, same seeded init
order as the reference => identical `state_dict`), their storage is handed to the native engine by
pointer.  `forward`, `train_step`, `evaluate` are calls through the C ABI.
"""

from __future__ import annotations

import random
from typing import List, Optional

import torch
import torch.nn as nn

from .. import _native as N
from ..engine import Engine, NetTensors

ADAM_BETAS = (0.9, 0.999)
ADAM_EPS = 1e-8


class NativeSessionNet(nn.Module):
    """embedding -> GRU -> heads, executed by the native engine.

    family: "gru4rec" | "bidir" | "sqn" | "smorl" | "bidir_sqn" | "sarm"
    """

    _TRUNK = {"gru4rec": "gru", "bidir": "gru", "sqn": "base_model", "smorl": "base_model", "bidir_sqn": "base_model",
              "sarm": "base_model"}
    _HEADS = {
        "gru4rec": ["output"],
        "bidir": ["output"],
        "sqn": ["sup_head_output", "q_head_output"],
        "bidir_sqn": ["sup_head_output", "q_head_output"],
        "smorl": ["sup_head_output", "q_head_acc", "q_head_div", "q_head_nov"],
        "sarm": ["q_heads.0", "q_heads.1", "q_heads.2", "q_heads.3", "q_heads.4"],  # nn.ModuleList (sarm.py:58-60)
    }

    def _build(self, family, hidden_dim, embedding_dim, item_num, state_size, action_dim, gru_layers,
               use_packed_seq, train_pad_embed, padding_idx, dropout=0.0):
        self._family = family
        self.hidden_dim = hidden_dim
        self.embedding_dim = embedding_dim
        self.item_num = int(item_num)
        self.state_size = state_size
        self.action_dim = action_dim
        self.use_packed_seq = use_packed_seq
        self.gru_layers = gru_layers
        self._bidirectional = family in ("bidir", "bidir_sqn")
        rl = family in ("sqn", "smorl", "bidir_sqn", "sarm")
        pad = self.item_num if padding_idx is None else padding_idx
        if rl and use_packed_seq:
            train_pad_embed = True  # sqn_gru.py:46-47
        self._frozen_pad_row = -1 if train_pad_embed else int(pad)
        # parameter containers -- construction order == reference (seeded init is bit-identical)
        self.embedding = nn.Embedding(self.item_num + 1, embedding_dim, padding_idx=None if train_pad_embed else pad)
        self.embedding.weight.data.normal_(mean=0, std=0.01)
        if not train_pad_embed:
            with torch.no_grad():
                self.embedding.weight[pad] = torch.zeros(embedding_dim)
        trunk = nn.GRU(input_size=embedding_dim, hidden_size=hidden_dim, num_layers=gru_layers, bias=True,
                       batch_first=True, bidirectional=self._bidirectional)
        setattr(self, self._TRUNK[family], trunk)
        if family == "bidir":
            self.dropout = nn.Dropout(p=dropout)
        d = hidden_dim * (2 if self._bidirectional else 1)
        if family == "sarm":
            self.q_heads = nn.ModuleList([nn.Linear(in_features=d, out_features=action_dim) for _ in range(5)])
        else:
            for name in self._HEADS[family]:
                setattr(self, name, nn.Linear(in_features=d, out_features=action_dim))
        for p in self.parameters():
            p.requires_grad_(False)  # gradients never exist as tensors: Adam is fused into the kernels
        self._engine: Optional[Engine] = None
        self._net_id = 0
        self._opt_m: Optional[NetTensors] = None
        self._opt_v: Optional[NetTensors] = None
        self._vocab_lo, self._vocab_hi = 0, int(action_dim)
        self._group = None
        # row-sharded embedding table (trainer.shard_vocabulary(..., shard_embedding=True)): (rank, world) of the owner
        # split; after a sharded train step the rows this rank does not own are stale until _sync_embedding()
        self._emb_shard = None
        self._emb_stale = self._emb_moments_stale = False

    def shard_vocabulary(self, lo: int, hi: int, group=None):
        """Keep only rows [lo, hi) of every head on this rank (vocabulary sharding, multi-GPU).
        Call on the freshly constructed (CPU) module so every rank starts from the same seeded init."""
        assert 0 <= lo < hi <= self.action_dim
        with torch.no_grad():
            for h in self._head_modules():
                h.weight.data = h.weight.data[lo:hi].clone().contiguous()
                h.bias.data = h.bias.data[lo:hi].clone().contiguous()
        self._vocab_lo, self._vocab_hi, self._group = int(lo), int(hi), group
        self._engine = None

    @property
    def is_sharded(self):
        return (self._vocab_lo, self._vocab_hi) != (0, int(self.action_dim))

    def _sync_embedding(self, with_moments: bool = False):
        """Row-sharded embedding table: make this rank's full-size copy current again -- every owner broadcasts its row
        slice (a COLLECTIVE: all ranks of the group must get here together, as they do in evaluate() / state_dict() of
        a sharded run).  The train step itself never needs this: it refreshes exactly the token rows it reads."""
        if self._emb_shard is None or not (self._emb_stale or (with_moments and self._emb_moments_stale)):
            return
        import torch.distributed as dist
        from ..sharded import shard_bounds
        rank, world = self._emb_shard
        tensors = [self.embedding.weight.data] if self._emb_stale else []
        if with_moments and self._emb_moments_stale and self._opt_m is not None:
            tensors += [self._opt_m.emb, self._opt_v.emb]
        for g in range(world):
            lo, hi = shard_bounds(self.item_num + 1, g, world)
            src = g if self._group is None else dist.get_global_rank(self._group, g)
            for t in tensors:
                dist.broadcast(t[lo:hi], src=src, group=self._group)
        self._emb_stale = False
        if with_moments:
            self._emb_moments_stale = False

    def state_dict(self, *a, **kw):
        self._sync_embedding()
        return super().state_dict(*a, **kw)

    # -- engine plumbing -------------------------------------------------------------------
    @property
    def _trunk(self) -> nn.GRU:
        return getattr(self, self._TRUNK[self._family])

    def _head_modules(self) -> List[nn.Linear]:
        if self._family == "sarm":
            return list(self.q_heads)
        return [getattr(self, n) for n in self._HEADS[self._family]]

    def _net_tensors(self) -> NetTensors:
        g = self._trunk
        sfx = ["", "_reverse"] if self._bidirectional else [""]
        nt = NetTensors(
            self.embedding.weight.data,
            [getattr(g, f"weight_ih_l0{s}").data for s in sfx], [getattr(g, f"weight_hh_l0{s}").data for s in sfx],
            [getattr(g, f"bias_ih_l0{s}").data for s in sfx], [getattr(g, f"bias_hh_l0{s}").data for s in sfx],
            [h.weight.data for h in self._head_modules()], [h.bias.data for h in self._head_modules()])
        nt.m, nt.v = self._opt_m, self._opt_v
        return nt

    def _param_device(self):
        return self.embedding.weight.device

    def _param_objects(self):
        """The Parameter objects behind _net_tensors(), looked up once (module attribute access is slow)."""
        g = self._trunk
        sfx = ["", "_reverse"] if self._bidirectional else [""]
        objs = [self.embedding.weight]
        for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
            objs += [getattr(g, n + x) for x in sfx]
        for h in self._head_modules():
            objs += [h.weight, h.bias]
        return objs

    def _attach(self, engine: Engine, net_id: int):
        self._engine, self._net_id = engine, net_id

    def _ready(self, batch_hint: int = 256) -> Engine:
        dev = self._param_device()
        if dev.type != "cuda":
            raise RuntimeError("this model executes on a B200 through the native engine; move it to a CUDA "
                               "device first (model.to('cuda') / trainer.send_to_device()). No CPU fallback.")
        self._sync_embedding()  # forward / evaluation on this net read arbitrary rows of the table
        if self._engine is None or self._engine.device != dev:
            self._engine = Engine(item_num=self.item_num, action_dim=self.action_dim,
                                  embedding_dim=self.embedding_dim, hidden_dim=self.hidden_dim,
                                  state_size=self.state_size, bidirectional=self._bidirectional,
                                  n_heads=len(self._HEADS[self._family]), n_nets=1,
                                  use_packed_seq=self.use_packed_seq, frozen_pad_row=self._frozen_pad_row,
                                  device=dev, max_batch=max(256, batch_hint), vocab_lo=self._vocab_lo,
                                  vocab_hi=self._vocab_hi)
            self._net_id = 0
        self._engine.bind(self._net_id, self._net_tensors())
        return self._engine

    def load_state_dict(self, *a, **kw):
        out = super().load_state_dict(*a, **kw)
        self._emb_stale = False  # every row was just written
        if self._engine is not None:  # GRU transposes inside the engine are stale now
            self._engine.bind(self._net_id, self._net_tensors(), force=True)
        return out

    def _dev_inputs(self, s, lengths):
        dev = self._param_device()
        if not isinstance(lengths, torch.Tensor):
            lengths = torch.as_tensor(lengths)
        if self.use_packed_seq and lengths.device.type == "cpu" and bool((lengths <= 0).any()):
            # same failure mode as pack_padded_sequence in the reference forward
            raise RuntimeError("Length of all samples has to be greater than 0, but found an element in "
                               "'lengths' that is <= 0")
        s = s.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        lengths = lengths.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        return s, lengths

    def final_state(self, s, lengths):
        eng = self._ready(int(s.shape[0]))
        s, lengths = self._dev_inputs(s, lengths)
        return eng.forward_state(self._net_id, s, lengths)

    def _all_logits(self, s, lengths):
        """Materialised logits of every head; on a sharded module: the LOCAL columns [lo, hi)."""
        eng = self._ready(int(s.shape[0]))
        s, lengths = self._dev_inputs(s, lengths)
        h = eng.forward_state(self._net_id, s, lengths)
        if self._family == "bidir" and self.training and self.dropout.p > 0:
            h = self.dropout(h)  # API-compatibility path only (BidirGRU4Rec/model.py:93); training runs in train_step
        return [eng.head_logits(self._net_id, i, h) for i in range(len(self._HEADS[self._family]))]

    def forward(self, s, lengths):
        outs = self._all_logits(s, lengths)
        if self._family == "sarm":
            return outs  # the reference returns the list of the five heads' outputs (sarm.py:74-75)
        if len(outs) == 1:
            return outs[0]
        if len(outs) == 2:
            return outs[0], outs[1]
        return outs[0], torch.stack(outs[1:], dim=1)


# ---------------------------------------------------------------------------------------------
class BatchStager:
    """Packs one replay-buffer batch into a pinned host buffer and ships it with ONE H2D copy.

    Layout (bytes): int64 s[B,L] | s_next[B,L] | a[B] | true_len[B] | true_next_len[B] | float32 r[B] | uint8 is_end[B]
    (the same layout rec_pack_batch produces).  Host-side packing uses numpy views of the pinned buffer
    (sub-microsecond copies); `depth` slots rotate so that a slot is never rewritten while its copy is in flight."""

    def __init__(self, device, L, depth=4):
        self.device, self.L, self.depth = device, L, depth
        self.cap = 0
        self.slots = []
        self.i = 0
        self.h2d_bytes = 0

    def _alloc(self, B):
        self.cap = B
        L = self.L
        n64 = B * (2 * L + 3)
        self.n_used = n64 * 8 + 5 * B
        self.nbytes = (self.n_used + 15) // 16 * 16
        self.slots = []
        for _ in range(self.depth):
            host = torch.empty(self.nbytes, dtype=torch.uint8).pin_memory()
            dev = torch.empty(self.nbytes, dtype=torch.uint8, device=self.device)
            self.slots.append(dict(host=host, dev=dev, ev=torch.cuda.Event(), hv=self._np_views(host.numpy(), B),
                                   dv=self._views(dev, B)))

    def _views(self, buf, B):
        L = self.L
        i64 = buf[: B * (2 * L + 3) * 8].view(torch.int64)
        o = 0
        s = i64[o:o + B * L].view(B, L); o += B * L
        sn = i64[o:o + B * L].view(B, L); o += B * L
        a = i64[o:o + B]; o += B
        ln = i64[o:o + B]; o += B
        nl = i64[o:o + B]; o += B
        off = B * (2 * L + 3) * 8
        r = buf[off: off + 4 * B].view(torch.float32)
        e = buf[off + 4 * B: off + 5 * B]
        return s, sn, a, ln, nl, r, e

    def _np_views(self, arr, B):
        import numpy as np
        L = self.L
        n64 = B * (2 * L + 3)
        i64 = arr[: n64 * 8].view(np.int64)
        o = 0
        s = i64[o:o + B * L].reshape(B, L); o += B * L
        sn = i64[o:o + B * L].reshape(B, L); o += B * L
        a = i64[o:o + B]; o += B
        ln = i64[o:o + B]; o += B
        nl = i64[o:o + B]; o += B
        r = arr[n64 * 8: n64 * 8 + 4 * B].view(np.float32)
        e = arr[n64 * 8 + 4 * B: n64 * 8 + 5 * B]
        return s, sn, a, ln, nl, r, e

    def stage(self, s, a, true_len, r=None, s_next=None, true_next_len=None, is_end=None):
        import numpy as np
        B = int(s.shape[0])
        if s.is_cuda:  # already resident: no staging, just dtype/contiguity
            f = lambda t, dt: None if t is None else t.to(device=self.device, dtype=dt).contiguous()
            return (f(s, torch.int64), f(s_next, torch.int64), f(a, torch.int64), f(true_len, torch.int64),
                    f(true_next_len, torch.int64), f(r, torch.float32),
                    None if is_end is None else is_end.to(device=self.device, dtype=torch.uint8).contiguous())
        if B != self.cap:
            if self.slots:
                torch.cuda.synchronize(self.device)
            self._alloc(B)
        slot = self.slots[self.i]
        self.i = (self.i + 1) % self.depth
        slot["ev"].synchronize()  # the copy that last used this slot has finished
        hs, hsn, ha, hln, hnl, hr, he = slot["hv"]
        np.copyto(hs, s.numpy(), casting="unsafe"); np.copyto(ha, a.numpy(), casting="unsafe")
        np.copyto(hln, true_len.numpy(), casting="unsafe")
        if r is not None:
            np.copyto(hsn, s_next.numpy(), casting="unsafe"); np.copyto(hnl, true_next_len.numpy(), casting="unsafe")
            np.copyto(hr, r.numpy().reshape(-1), casting="unsafe"); np.copyto(he, is_end.numpy(), casting="unsafe")
        slot["dev"].copy_(slot["host"], non_blocking=True)
        slot["ev"].record()
        self.h2d_bytes = self.n_used
        ds, dsn, da, dln, dnl, dr, de = slot["dv"]
        if r is None:
            return ds, None, da, dln, None, None, None
        return ds, dsn, da, dln, dnl, dr, de


def make_hparams(lr, gamma=0.0, alpha=1.0, q_weights=(1.0, 0.0, 0.0)):
    hp = N.RecTrainHparams()
    hp.lr, hp.beta1, hp.beta2, hp.eps = lr, ADAM_BETAS[0], ADAM_BETAS[1], ADAM_EPS
    hp.gamma, hp.alpha = gamma, alpha
    for i in range(3):
        hp.q_weights[i] = float(q_weights[i]) if i < len(q_weights) else 0.0
    hp.pad_pos_end = 1
    return hp


class NativeTrainerBase:
    """Owns the engine, the twin nets' Adam state and the batch stager."""

    def _setup(self, nets: List[NativeSessionNet], device, learning_rate, torch_rand_seed=None, python_rand_seed=None):
        self._nets = nets
        self.device = device
        self.learning_rate = learning_rate
        self.cross_entropy_loss = nn.CrossEntropyLoss(weight=None, reduction="mean")
        self._engine: Optional[Engine] = None
        self._stager: Optional[BatchStager] = None
        self._loss_dev = None
        self._pending_steps = [0 for _ in nets]
        self._shard = None
        self._fast_bind = None

    @staticmethod
    def _seed(torch_rand_seed, python_rand_seed):
        torch.manual_seed(torch_rand_seed)
        random.seed(python_rand_seed)

    def _drop_engine(self):
        """Forget the native engine (parameter storage moved / was re-sharded).  The Adam step counters live in
        the engine: read them back first, or the next engine would restart bias correction at t = 1 on warm
        moments (and repeat the step-keyed dropout masks)."""
        if self._engine is not None and self._engine.handle is not None:
            for i in range(len(self._nets)):
                self._pending_steps[i] = self._engine.adam_step(i)
        self._engine = None
        self._fast_bind = None

    def send_to_device(self):
        dev = torch.device(self.device)
        if dev.type == "cuda" and dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if all(p.device == dev for n in self._nets for p in n._param_objects()):
            return  # already resident: the engine, its bound pointers and the optimizer state stay valid
        self._drop_engine()
        for n in self._nets:
            n.to(self.device)

    _FULL_CHECK_EVERY = 64

    def _ready(self, B) -> Engine:
        # fast path (a few us): same storage behind every cached Parameter object as at the last full bind.
        # A full re-derivation (module walk + rec_bind_params when anything moved) still runs every
        # _FULL_CHECK_EVERY calls, so even replaced Parameter objects are picked up.
        fast = getattr(self, "_fast_bind", None)
        if fast is not None and self._engine is not None and B <= self._engine.cfg["max_batch"]:
            objs, sig, opt_ids, calls = fast
            if calls < self._FULL_CHECK_EVERY and opt_ids == tuple(id(n._opt_m) for n in self._nets) \
                    and sig == tuple(p.data_ptr() for p in objs):
                fast[3] = calls + 1
                return self._engine
        eng = self._ready_full(B)
        objs = [p for n in self._nets for p in n._param_objects()]
        self._fast_bind = [objs, tuple(p.data_ptr() for p in objs), tuple(id(n._opt_m) for n in self._nets), 0]
        return eng

    def _ready_full(self, B) -> Engine:
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError(f"device={self.device!r}: the native trainers run on a B200 only (no CPU fallback)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if self._nets[0]._param_device() != dev:
            self._drop_engine()
            for n in self._nets:
                n.to(dev)
        if self._engine is None:
            n0 = self._nets[0]
            self._engine = Engine(item_num=n0.item_num, action_dim=n0.action_dim, embedding_dim=n0.embedding_dim,
                                  hidden_dim=n0.hidden_dim, state_size=n0.state_size,
                                  bidirectional=n0._bidirectional, n_heads=len(n0._HEADS[n0._family]),
                                  n_nets=len(self._nets), use_packed_seq=n0.use_packed_seq,
                                  frozen_pad_row=n0._frozen_pad_row, device=dev, max_batch=max(256, B),
                                  vocab_lo=n0._vocab_lo, vocab_hi=n0._vocab_hi)
            for i, n in enumerate(self._nets):
                if n._opt_m is None or n._opt_m.emb.device != dev:
                    base = n._net_tensors()
                    n._opt_m, n._opt_v = base.zeros_like(), base.zeros_like()
                n._attach(self._engine, i)
                self._engine.bind(i, n._net_tensors(), force=True)
                self._engine.set_adam_step(i, self._pending_steps[i])
            if n0._emb_shard is not None:
                from ..sharded import shard_bounds
                self._engine.set_embedding_shard(*shard_bounds(n0.item_num + 1, *n0._emb_shard))
            self._stager = BatchStager(dev, n0.state_size)
            self._loss_dev = torch.zeros(8, dtype=torch.float32, device=dev)
        else:
            for i, n in enumerate(self._nets):
                self._engine.bind(i, n._net_tensors())
        return self._engine

    def release_graphs(self):
        """Drop CUDA graphs that captured NCCL collectives (call before dist.destroy_process_group())."""
        st = getattr(self, "_sharded_step", None)
        if st is not None:
            st.release_graphs()

    def shard_vocabulary(self, rank: int, world: int, group=None, shard_embedding=None):
        """Vocabulary-shard every head of every net over `world` ranks (call before send_to_device).
        shard_embedding: also row-shard the embedding table's Adam sweep (rows [(N+1) r / G, (N+1)(r+1) / G) are
        owned by rank r, token rows travel from their owners before each step).  None: on for world > 1 (measured on cfg4:
        +1.5 % sessions/s at 2 GPUs, +7.5 % at 4), REC_SHARD_EMBEDDING=0 / 1 forces either."""
        import os
        from ..sharded import shard_bounds
        V = self._nets[0].action_dim
        lo, hi = shard_bounds(V, rank, world)
        if shard_embedding is None:
            shard_embedding = os.environ.get("REC_SHARD_EMBEDDING", "1" if world > 1 else "0") == "1"
        if self._nets[0].embedding_dim % 4:
            shard_embedding = False
        self._drop_engine()
        for n in self._nets:
            n.shard_vocabulary(lo, hi, group)
            n._emb_shard = (rank, world) if shard_embedding else None
        self._shard = (rank, world, group)

    def _set_mode(self, train: bool):
        for n in self._nets:
            n.train(train)

    def set_train(self):
        self._set_mode(True)

    def set_eval(self):
        self._set_mode(False)

    # optimizer state for checkpoint/resume (the reference never saves it; offered as an extension)
    def optimizer_state(self):
        out = []
        for i, n in enumerate(self._nets):
            n._sync_embedding(with_moments=True)  # row-sharded table: a collective (see _sync_embedding)
            step = self._engine.adam_step(i) if self._engine is not None else self._pending_steps[i]
            out.append(dict(step=step, exp_avg=None if n._opt_m is None else [t.clone() for t in n._opt_m.all()],
                            exp_avg_sq=None if n._opt_v is None else [t.clone() for t in n._opt_v.all()]))
        return out
