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
    try:
        for s, a, s_len in evaluation_data_loader:
            B = int(s.shape[0])
            ds, dl = model._dev_inputs(s, s_len)
            da = a.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
            eng = model._ready(B)
            if (eng, eng.handle) != held:  # first batch, or the engine was re-created for a larger batch
                eng.eval_hold_params(True)
                held = (eng, eng.handle)
            _eval_one_batch(model, eng, eng._batch(B, ds, da, dl), o, acc, kmax)
            n_total += B
            n_batches += 1
    finally:
        if held is not None and held[0].handle == held[1]:
            held[0].eval_hold_params(False)
    r = acc.read()
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
