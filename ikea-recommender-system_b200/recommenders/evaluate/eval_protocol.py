"""Drop-in for `recommenders/evaluate/eval_protocol.py` (reference :123-359): `evaluate` and
`update_train_metrics` with the reference's keyword arguments and return types, executed on the B200.

One fused pass per batch (GRU forward -> head GEMM with running top-k and online log-sum-exp ->
metric kernels) replaces the reference's forward + five `torch.topk` calls + a B x V device-to-host copy;
only accumulators, coverage bitmaps and (for update_train_metrics) the [B, k] id list leave the GPU.
Top-k order is (score desc, id asc) -- `torch.topk` leaves ties unspecified.
"""

from __future__ import annotations

import numpy as np
import torch

from ... import _native as N
from ...engine import EvalAccumulators

class _LRU(dict):
    """Small bounded cache of device-side lookup tables (diversity table, unpopular bitmap, token LUT)."""

    MAX = 16

    def __setitem__(self, k, v):
        if len(self) >= self.MAX and k not in self:
            self.pop(next(iter(self)))
        super().__setitem__(k, v)


_cache = _LRU()


def _as_div_table(diversity_embedding, device):
    """Frozen nn.Embedding (or tensor) -> contiguous fp32 table on `device` (cached by storage)."""
    w = diversity_embedding.weight if hasattr(diversity_embedding, "weight") else diversity_embedding
    if w.is_cuda and w.device == torch.device(device) and w.dtype == torch.float32 and w.is_contiguous():
        return w.detach()  # used in place: no copy to go stale
    key = ("div", w.data_ptr(), int(getattr(w, "_version", 0)), tuple(w.shape), str(device))
    if key not in _cache:
        _cache[key] = (w, w.detach().to(device=device, dtype=torch.float32).contiguous())
    return _cache[key][1]


def _unpopular_bitmap(unpopular_actions_set, action_dim, device, packed=False):
    """python set -> uint8[V] membership table on the device (novelty.py:5-9 `in unpopular_items`); packed=True: the
    same set as uint32 bit words on the host (bit i of word w = action 32 w + i), the layout of the coverage bitmaps."""
    key = ("unpop", id(unpopular_actions_set), len(unpopular_actions_set), action_dim, str(device))
    if key not in _cache:
        bm = np.zeros(action_dim, dtype=np.uint8)
        ids = np.fromiter((int(i) for i in unpopular_actions_set), dtype=np.int64, count=len(unpopular_actions_set))
        ids = ids[(ids >= 0) & (ids < action_dim)]
        bm[ids] = 1
        bits = np.packbits(bm, bitorder="little")
        words = np.concatenate([bits, np.zeros((-len(bits)) % 4, dtype=np.uint8)]).view(np.uint32)
        _cache[key] = (unpopular_actions_set, torch.from_numpy(bm).to(device), words)
    return _cache[key][2 if packed else 1]


def _token_lut(input_tokenizer, output_tokenizer, action_dim, device):
    """`input_tokenizer.stoi(output_tokenizer.itos(x))` (diversity.py:55-60) as an int64[V] table."""
    if input_tokenizer is None:
        return None
    key = ("lut", id(input_tokenizer), id(output_tokenizer), action_dim, str(device))
    if key not in _cache:
        lut = np.fromiter((input_tokenizer.stoi(output_tokenizer.itos(x)) for x in range(action_dim)),
                          dtype=np.int64, count=action_dim)
        _cache[key] = ((input_tokenizer, output_tokenizer), torch.from_numpy(lut).to(device))
    return _cache[key][1]


def _opts(model, device, head_idx, topk_hr_ndcg, topk_div, topk_nov, topk_cov, nov_rew, padding_pos,
          diversity_embedding, unpopular_actions_set, input_tokenizer, output_tokenizer):
    if len(topk_hr_ndcg) > N.REC_MAX_KLIST or len(topk_cov) > N.REC_MAX_KLIST:
        raise ValueError(f"at most {N.REC_MAX_KLIST} entries per top-k list")
    kmax = max([*topk_hr_ndcg, *topk_cov, topk_div, topk_nov])
    if kmax > N.REC_MAX_TOPK:
        raise ValueError(f"top-k up to {N.REC_MAX_TOPK} is supported, got {kmax}")
    o = N.RecEvalOpts()
    o.head_idx = head_idx
    o.n_k = len(topk_hr_ndcg)
    for i, k in enumerate(topk_hr_ndcg):
        o.ks[i] = int(k)
    o.n_cov = len(topk_cov)
    for i, k in enumerate(topk_cov):
        o.cov_ks[i] = int(k)
    o.topk_div, o.topk_nov, o.nov_reward = int(topk_div), int(topk_nov), float(nov_rew)
    div = _as_div_table(diversity_embedding, device)
    unpop = _unpopular_bitmap(unpopular_actions_set, model.action_dim, device)
    lut = _token_lut(input_tokenizer, output_tokenizer, model.action_dim, device)
    o.div_emb, o.div_dim = div.data_ptr(), int(div.shape[1])
    o.unpopular = unpop.data_ptr()
    o.out_to_in = None if lut is None else lut.data_ptr()
    o.pad_pos_end = 1 if padding_pos == "end" else 0
    return o, kmax, (div, unpop, lut)


def _check_loss(loss_function):
    if loss_function is not None and not isinstance(loss_function, torch.nn.CrossEntropyLoss):
        raise NotImplementedError("the fused evaluation computes nn.CrossEntropyLoss(reduction='mean') only")


def get_preds(states, true_len, model, head_idx):
    """reference :103-120 (materialises logits; API compatibility only)."""
    out = model(states, true_len)
    return out[head_idx] if isinstance(out, tuple) else out


def _coverage(cov_bits, topk_cov, unpop_words, num_actions, n_unpop):
    """{k: (share of the unpopular actions covered, share of all actions covered)} (coverage.py:24-53) from the device
    bitmaps: population counts on the packed words (unpacking 4 x 1 M bits cost 5 ms per evaluate() at 1 M items)."""
    res = {}
    n_words = (num_actions + 31) // 32
    tail = np.uint32(0xFFFFFFFF >> ((32 - num_actions % 32) % 32))  # bits of the last word that are actions
    for i, k in enumerate(topk_cov):
        w = cov_bits[i][:n_words].copy()
        w[-1] &= tail
        res[k] = (int(np.bitwise_count(w & unpop_words[:n_words]).sum()) / n_unpop, int(np.bitwise_count(w).sum()) / num_actions)
    return res


def _eval_one_batch(model, eng, batch, o, acc, kmax, topk_ids=None, virtual_gather=None):
    """One fused evaluation batch.  On a vocabulary-sharded model (SURVEY 8e "collectives for eval", BASELINE
    configs[4]): every rank scores the batch against its vocabulary slice and publishes ONE record per row
    (max, sum-exp, target logit, its fp32-exact top-k candidates) -> one all-gather of the records -> every rank
    merges them ((score desc, id asc), log-sum-exp combine) and accumulates the metrics.  Inputs and therefore the
    merged results are replicated, so the coverage bitmaps / accumulators need no further reduction."""
    if not model.is_sharded:
        eng.eval_batch(model._net_id, batch, o, acc.struct, topk_ids=topk_ids)
        return
    import torch.distributed as dist
    B, rec = int(batch.B), eng.record_floats()
    dev = model._param_device()
    records = torch.empty(B, rec, dtype=torch.float32, device=dev)
    eng.eval_shard_candidates(model._net_id, batch, o.head_idx, kmax, records)
    if virtual_gather is not None:  # tests: several shards living on one GPU
        gathered = virtual_gather(records)
    else:
        world = dist.get_world_size(model._group)
        gathered = torch.empty(world, B, rec, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(gathered, records, group=model._group)
    eng.eval_merge(batch, o, gathered, int(gathered.shape[0]), acc.struct, topk_ids=topk_ids)


class _FullHeadReplica:
    """Evaluation of a vocabulary-sharded net sharded by SESSIONS instead (sessions are independent): every
    rank gets an unsharded copy of the ONE head `evaluate` scores (one all-gather of the shards' rows per sweep: 256 MB at
    1 M items, a fraction of a millisecond over NVLink), binds it with the replicated embedding + GRU to a second,
    unsharded engine and evaluates every world-th batch against the full catalogue with the single-GPU kernels;
    accumulators are summed and coverage bitmaps OR-ed once at the end.  Per-batch collectives and the replicated work of
    the vocabulary-sharded sweep (GRU forward, exact scoring, merge, metrics on every rank for every batch) disappear, so
    the sweep scales with the number of GPUs.  Duck-types what `evaluate` needs of a NativeSessionNet."""

    is_sharded = False
    _net_id = 0

    def __init__(self, model, head_idx):
        import torch.distributed as dist
        from ...sharded import shard_bounds
        self.src, self.head_idx, self.group = model, head_idx, model._group
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.action_dim = model.action_dim
        self.bounds = [shard_bounds(model.action_dim, g, self.world) for g in range(self.world)]
        dev = model._param_device()
        D = model._head_modules()[head_idx].weight.shape[1]
        self.rows_max = max(hi - lo for lo, hi in self.bounds)
        self.send = torch.zeros(self.rows_max, D + 1, dtype=torch.float32, device=dev)
        self.recv = torch.empty(self.world, self.rows_max, D + 1, dtype=torch.float32, device=dev)
        self.W = torch.empty(model.action_dim, D, dtype=torch.float32, device=dev)
        self.b = torch.empty(model.action_dim, dtype=torch.float32, device=dev)
        self._engine = None

    @staticmethod
    def plan(model, head_idx):
        """The replica of `model` for this sweep, or None when the sweep runs on `model` itself (unsharded model, no
        process group, REC_EVAL_SHARD=vocab, or shards that are not the balanced split of the group)."""
        import os
        import torch.distributed as dist
        from ...sharded import shard_bounds
        if not model.is_sharded or os.environ.get("REC_EVAL_SHARD", "sessions") == "vocab":
            return None
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world, rank = dist.get_world_size(model._group), dist.get_rank(model._group)
        if world < 2 or shard_bounds(model.action_dim, rank, world) != (model._vocab_lo, model._vocab_hi):
            return None
        rep = getattr(model, "_eval_replica", None)
        if rep is None or rep.head_idx != head_idx or rep.world != world or rep.W.device != model._param_device():
            rep = model._eval_replica = _FullHeadReplica(model, head_idx)
        rep.refresh()
        return rep

    def refresh(self):
        """All-gather the scored head's rows (weights | bias) of every shard into the full [V, D] / [V] copies."""
        import torch.distributed as dist
        h = self.src._head_modules()[self.head_idx]
        n, D = h.weight.shape
        self.send[:n, :D].copy_(h.weight.data)
        self.send[:n, D].copy_(h.bias.data)
        dist.all_gather_into_tensor(self.recv.view(-1), self.send.view(-1), group=self.group)
        for g, (lo, hi) in enumerate(self.bounds):
            self.W[lo:hi].copy_(self.recv[g, :hi - lo, :D])
            self.b[lo:hi].copy_(self.recv[g, :hi - lo, D])

    def mine(self, batch_index):
        return batch_index % self.world == self.rank

    def _param_device(self):
        return self.src._param_device()

    def _dev_inputs(self, s, lengths):
        return self.src._dev_inputs(s, lengths)

    def _ready(self, batch_hint=256):
        from ...engine import Engine
        m = self.src
        if self._engine is None:
            self._engine = Engine(item_num=m.item_num, action_dim=m.action_dim, embedding_dim=m.embedding_dim,
                                  hidden_dim=m.hidden_dim, state_size=m.state_size, bidirectional=m._bidirectional,
                                  n_heads=len(m._HEADS[m._family]), n_nets=1, use_packed_seq=m.use_packed_seq,
                                  frozen_pad_row=m._frozen_pad_row, device=self._param_device(),
                                  max_batch=max(256, batch_hint))
        nt = m._net_tensors()  # embedding + GRU: the replicated tensors themselves; every head slot: the scored head
        nt.head_w, nt.head_b = [self.W for _ in nt.head_w], [self.b for _ in nt.head_b]
        nt.m = nt.v = None
        self._engine.bind(0, nt)
        return self._engine

    def reduce(self, acc):
        """Sum of the ranks' accumulators, OR of their coverage bitmaps (in place, identical on every rank)."""
        import torch.distributed as dist
        dist.all_reduce(acc.f64, group=self.group)
        allcov = torch.empty(self.world, acc.cov.numel(), dtype=torch.int32, device=acc.cov.device)
        dist.all_gather_into_tensor(allcov.view(-1), acc.cov.view(-1), group=self.group)
        out = allcov[0]
        for g in range(1, self.world):
            out = torch.bitwise_or(out, allcov[g])
        acc.cov.copy_(out)


def evaluate(evaluation_data_loader, model, device, loss_function, padding_pos, diversity_embedding,
             unpopular_actions_set, head_idx=0, topk_hr_ndcg=[5, 10, 20], topk_to_consider_div=1,
             topk_to_consider_nov=1, topk_to_consider_cov=[1, 5, 10], novelty_rew_signal=1, input_tokenizer=None,
             output_tokenizer=None):
    """Returns (loss, hr, ndcg, coverage_res, avg_diversity_rew, avg_novelty_rew, repetitions) with the
    reference's types: torch scalar, np.ndarray, np.ndarray, {k: (unpop_cov, all_cov)}, torch scalar,
    np.float64, np.ndarray."""
    _check_loss(loss_function)
    model.eval()
    eng = model._ready()
    dev = model._param_device()
    o, kmax, keep = _opts(model, dev, head_idx, topk_hr_ndcg, topk_to_consider_div, topk_to_consider_nov,
                          topk_to_consider_cov, novelty_rew_signal, padding_pos, diversity_embedding,
                          unpopular_actions_set, input_tokenizer, output_tokenizer)
    acc = EvalAccumulators(dev, model.action_dim)
    n_total, n_batches = 0, 0
    held = None  # the engine that was told "parameters are frozen for this sweep" (model.eval(): nothing trains here)
    # a vocabulary-sharded model under a process group evaluates sharded by SESSIONS: this rank scores every world-th
    # batch against the full catalogue (REC_EVAL_SHARD=vocab keeps the per-batch candidate exchange)
    import os
    import time
    trace = [("start", time.perf_counter())] if os.environ.get("REC_EVAL_TRACE") else None  # per-phase wall clock (debug)

    def stamp(what):
        if trace is not None:
            torch.cuda.synchronize(dev)
            trace.append((what, time.perf_counter()))

    replica = _FullHeadReplica.plan(model, head_idx)
    stamp("plan")
    net = model if replica is None else replica
    try:
        for s, a, s_len in evaluation_data_loader:
            B = int(s.shape[0])
            if replica is None or replica.mine(n_batches):
                ds, dl = net._dev_inputs(s, s_len)
                da = a.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
                eng = net._ready(B)
                if (eng, eng.handle) != held:  # first batch, or the engine was re-created for a larger batch
                    eng.eval_hold_params(True)
                    held = (eng, eng.handle)
                _eval_one_batch(net, eng, eng._batch(B, ds, da, dl), o, acc, kmax)
                stamp(f"batch{n_batches}")
            n_total += B
            n_batches += 1
    finally:
        if held is not None and held[0].handle == held[1]:
            held[0].eval_hold_params(False)
    if replica is not None:
        replica.reduce(acc)
        stamp("reduce")
    r = acc.read()
    if trace is not None:
        stamp("read")
        print("REC_EVAL_TRACE", " ".join(f"{w}={1e3 * (t1 - t0):.3f}ms" for (_, t0), (w, t1) in zip(trace, trace[1:])), flush=True)
    nk = len(topk_hr_ndcg)
    hr = r["hits"][:nk] / n_total
    ndcg = r["ndcg"][:nk] / n_total
    reps = r["reps"][:nk] / n_total
    loss = torch.tensor(r["loss_sum"] / n_batches, dtype=torch.float32, device=dev)
    avg_div = torch.tensor(r["div_sum"] / n_total, dtype=torch.float32, device=dev)
    avg_nov = np.float64(r["nov_sum"] / n_total)
    unpop_words = _unpopular_bitmap(unpopular_actions_set, model.action_dim, dev, packed=True)
    cov = _coverage(r["cov_bits"], topk_to_consider_cov, unpop_words, model.action_dim, len(unpopular_actions_set))
    return loss, hr, ndcg, cov, avg_div, avg_nov, reps


def update_train_metrics(s, a, s_len, model, device, padding_pos, diversity_embedding, unpopular_actions_set,
                         actions_covered_topk_dict, head_idx=0, topk_hr_ndcg=[5, 10, 20], topk_to_consider_div=1,
                         topk_to_consider_nov=1, topk_to_consider_cov=[1, 5, 10], novelty_rew_signal=1,
                         input_tokenizer=None, output_tokenizer=None):
    """One batch (reference :266-359): returns (hr_batch, ndcg_batch, actions_covered_topk_dict,
    batch_div_rew, batch_nov_rew, batch_repetitions) -- sums, not means."""
    model.eval()
    B = int(s.shape[0])
    eng = model._ready(B)
    dev = model._param_device()
    o, kmax, keep = _opts(model, dev, head_idx, topk_hr_ndcg, topk_to_consider_div, topk_to_consider_nov,
                          topk_to_consider_cov, novelty_rew_signal, padding_pos, diversity_embedding,
                          unpopular_actions_set, input_tokenizer, output_tokenizer)
    acc = EvalAccumulators(dev, model.action_dim)
    ds, dl = model._dev_inputs(s, s_len)
    da = a.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
    ids = torch.empty(B, kmax, dtype=torch.int32, device=dev)
    _eval_one_batch(model, eng, eng._batch(B, ds, da, dl), o, acc, kmax, topk_ids=ids)
    r = acc.read()
    nk = len(topk_hr_ndcg)
    ids_h = ids.cpu().numpy()
    for k in topk_to_consider_cov:  # python sets, as the reference returns them (coverage.py:46-51)
        actions_covered_topk_dict[k] = actions_covered_topk_dict[k].union(ids_h[:, :k].flatten().tolist())
    div = torch.tensor(r["div_sum"], dtype=torch.float32, device=dev)
    return (r["hits"][:nk].copy(), r["ndcg"][:nk].copy(), actions_covered_topk_dict, div, np.float64(r["nov_sum"]),
            r["reps"][:nk].copy())
