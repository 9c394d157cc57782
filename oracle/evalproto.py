"""Oracle evaluation protocol: HR/NDCG@k, coverage, diversity, novelty, repetitions.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates recommenders/evaluate/{eval_protocol,coverage,diversity,novelty,repetitiveness}.py
with ONE tie-stable top-k per batch instead of the reference's five `torch.topk` calls
(eval_protocol.py:75, coverage.py:43, diversity.py:52, novelty.py:35, repetitiveness.py:39).
`torch.topk` leaves the order among equal scores unspecified; the oracle (and the CUDA
path) order by (score desc, id asc).  The tokenizer round trip
`input_tokenizer.stoi(output_tokenizer.itos(x))` (diversity.py:55-60, repetitiveness.py:42-46)
is restated as an integer look-up table `out_to_in[V]`.
"""

from __future__ import annotations

import numpy as np
import torch


# ----------------------------------------------------------------------------- top-k
def stable_topk(scores: torch.Tensor, k: int) -> torch.Tensor:
    """ids [B, k] ordered by (score desc, id asc)."""
    scores = scores.detach().to("cpu")
    B, V = scores.shape
    k = min(k, V)
    if V <= 4096:
        order = torch.sort(-scores, dim=1, stable=True).indices
        return order[:, :k].contiguous()
    k2 = min(V, k + 64)
    vals, ids = torch.topk(scores, k2, dim=1)  # contains every element > k-th value
    vals_n, ids_n = vals.numpy(), ids.numpy()
    out = np.empty((B, k), dtype=np.int64)
    for b in range(B):
        if k2 < V and vals_n[b, k2 - 1] == vals_n[b, k - 1]:
            # a tie group straddles the slab edge: resolve it on the full row
            row = scores[b].numpy()
            cand = np.nonzero(row >= vals_n[b, k - 1])[0]
            o = np.lexsort((cand, -row[cand]))
            out[b] = cand[o][:k]
        else:
            o = np.lexsort((ids_n[b], -vals_n[b]))
            out[b] = ids_n[b][o][:k]
    return torch.from_numpy(out)


# ----------------------------------------------------------------------------- per-batch helpers
def last_action(s, padding_pos, s_len=None):
    """diversity.py:4-12"""
    if padding_pos == "end":
        return s.gather(1, (s_len.to(s.device) - 1).unsqueeze(1)).squeeze(1)
    return s[:, -1]


def _as_embedding(e):
    if isinstance(e, torch.nn.Embedding):
        return e
    return torch.nn.Embedding.from_pretrained(torch.as_tensor(e), freeze=True)


def diversity_rewards(s, preds, len_states, padding_pos, topk, embedding, out_to_in=None, topk_ids=None):
    """1 - mean_k cos(E_div[last(s)], E_div[top-k]) -- diversity.py:15-73 (eps 1e-6, dim 2)."""
    emb = _as_embedding(embedding)
    ids = stable_topk(preds, topk) if topk_ids is None else topk_ids[:, :topk]
    if out_to_in is not None:
        ids = torch.as_tensor(out_to_in)[ids]
    la = last_action(s, padding_pos, len_states)
    cos = torch.nn.CosineSimilarity(dim=2, eps=1e-6)
    sim = cos(emb(la).unsqueeze(1), emb(ids))
    return 1 - torch.mean(sim, dim=1)


def novelty_rewards(preds, unpopular_items, reward=1, topk=1, topk_ids=None):
    """mean over top-k of [id in unpopular]*reward, float64 numpy [B] -- novelty.py:12-47."""
    ids = (stable_topk(preds, topk) if topk_ids is None else topk_ids[:, :topk]).numpy()
    unpop = np.fromiter(unpopular_items, dtype=np.int64, count=len(unpopular_items))
    flags = np.isin(ids, unpop).astype(int) * reward
    return np.mean(flags, axis=1)


def hits_and_ndcg(preds, true_idx, top_k=(5, 10, 20), topk_ids=None):
    """(#hits per k, sum of 1/log2(rank+1) per k) -- eval_protocol.py:26-100."""
    kmax = max(top_k)
    ids = (stable_topk(preds, kmax) if topk_ids is None else topk_ids[:, :kmax]).numpy()
    truth = true_idx.reshape(-1, 1).numpy()
    hits = np.zeros(len(top_k))
    ndcg = np.zeros(len(top_k))
    for i, k in enumerate(top_k):
        match = ids[:, :k] == truth
        found = match.any(1)
        rank = np.where(found, match.argmax(1) + 1, 0).astype(np.float64)
        hits[i] = found.sum()
        with np.errstate(divide="ignore"):
            gain = np.where(found, 1.0 / np.log2(rank + 1), 0.0)
        ndcg[i] = gain.sum()
    return hits, ndcg


def repetitions(s, preds, topk=(1,), out_to_in=None, topk_ids=None):
    """per k: sum_b sum_{i<k} sum_t [s[b,t] == top_i(b)] -- repetitiveness.py:21-57."""
    kmax = max(topk)
    ids = stable_topk(preds, kmax) if topk_ids is None else topk_ids[:, :kmax]
    if out_to_in is not None:
        ids = torch.as_tensor(out_to_in)[ids]
    ids = ids.numpy()
    st = s.cpu().numpy()[:, :, None]
    res = np.zeros(len(topk))
    for i, k in enumerate(topk):
        res[i] = (st == ids[:, None, :k]).sum()
    return res


def coverage_update(covered, preds, top_k, topk_ids=None):
    """set-union of top-k ids per k -- coverage.py:24-53 (ids are OUTPUT ids, no remap)."""
    kmax = max(top_k)
    ids = stable_topk(preds, kmax) if topk_ids is None else topk_ids[:, :kmax]
    for k in top_k:
        covered[k] = covered[k].union(ids[:, :k].flatten().tolist())
    return covered


def coverage_result(covered, unpopular_set, num_actions, topk):
    """k -> (|set & unpop|/|unpop|, |set|/V) -- coverage.py:4-21,56-74."""
    return {k: (len(covered[k].intersection(unpopular_set)) / len(unpopular_set),
                len(covered[k]) / num_actions) for k in topk}


def _preds(model, s, true_len, head_idx):
    out = model(s, true_len)  # eval_protocol.py:103-120
    return out[head_idx] if isinstance(out, tuple) else out


# ----------------------------------------------------------------------------- orchestration
def evaluate(batches, model, loss_function, padding_pos, diversity_embedding, unpopular_actions_set,
             head_idx=0, topk_hr_ndcg=(5, 10, 20), topk_to_consider_div=1, topk_to_consider_nov=1,
             topk_to_consider_cov=(1, 5, 10), novelty_rew_signal=1, out_to_in=None):
    """eval_protocol.py:123-263.  `batches` is any iterable of (s, a, s_len) with a len()."""
    model.eval()
    kmax = max(max(topk_hr_ndcg), topk_to_consider_div, topk_to_consider_nov, max(topk_to_consider_cov))
    with torch.no_grad():
        n = 0
        nov_total = 0
        div_total = 0
        hr = np.zeros(len(topk_hr_ndcg))
        ndcg = np.zeros(len(topk_hr_ndcg))
        reps = np.zeros(len(topk_hr_ndcg))
        covered = {k: set() for k in topk_to_consider_cov}
        loss = 0
        n_batches = 0
        for s, a, s_len in batches:
            n_batches += 1
            preds = _preds(model, s, s_len, head_idx)
            loss += loss_function(preds, a)  # mean of batch means (q4)
            ids = stable_topk(preds, kmax)
            div_total += torch.sum(diversity_rewards(s, preds, s_len, padding_pos, topk_to_consider_div,
                                                     diversity_embedding, out_to_in, topk_ids=ids))
            nov_total += novelty_rewards(preds, unpopular_actions_set, novelty_rew_signal,
                                         topk_to_consider_nov, topk_ids=ids).sum()
            h, g = hits_and_ndcg(preds, a, topk_hr_ndcg, topk_ids=ids)
            hr += h
            ndcg += g
            n += len(a)
            covered = coverage_update(covered, preds, topk_to_consider_cov, topk_ids=ids)
            reps += repetitions(s, preds, topk_hr_ndcg, out_to_in, topk_ids=ids)
        cov = coverage_result(covered, unpopular_actions_set, model.action_dim, topk_to_consider_cov)
    return loss / n_batches, hr / n, ndcg / n, cov, div_total / n, nov_total / n, reps / n


def update_train_metrics(s, a, s_len, model, padding_pos, diversity_embedding, unpopular_actions_set,
                         actions_covered_topk_dict, head_idx=0, topk_hr_ndcg=(5, 10, 20),
                         topk_to_consider_div=1, topk_to_consider_nov=1, topk_to_consider_cov=(1, 5, 10),
                         novelty_rew_signal=1, out_to_in=None):
    """eval_protocol.py:266-359 (one batch; sums, not means)."""
    model.eval()
    kmax = max(max(topk_hr_ndcg), topk_to_consider_div, topk_to_consider_nov, max(topk_to_consider_cov))
    with torch.no_grad():
        preds = _preds(model, s, s_len, head_idx)
        ids = stable_topk(preds, kmax)
        div = torch.sum(diversity_rewards(s, preds, s_len, padding_pos, topk_to_consider_div,
                                          diversity_embedding, out_to_in, topk_ids=ids))
        nov = novelty_rewards(preds, unpopular_actions_set, novelty_rew_signal, topk_to_consider_nov,
                              topk_ids=ids).sum()
        h, g = hits_and_ndcg(preds, a, topk_hr_ndcg, topk_ids=ids)
        covered = coverage_update(actions_covered_topk_dict, preds, topk_to_consider_cov, topk_ids=ids)
        reps = repetitions(s, preds, topk_hr_ndcg, out_to_in, topk_ids=ids)
    return h, g, covered, div, nov, reps
