// Small fused kernels around the heads: SMORL online rewards, double-Q TD target/loss, and the
// on-device evaluation metrics (HR/NDCG@k, coverage bitmaps, diversity, novelty, repetitions).
//
// Reference semantics: evaluate/diversity.py:4-73, evaluate/novelty.py:12-47,
// evaluate/eval_protocol.py:26-100,123-263, evaluate/coverage.py:24-74,
// evaluate/repetitiveness.py:21-57, models/SQN/sqn_gru.py:231-245, models/SMORL/smorl_gru.py:291-325,
// utils/tensor_operations.py:36-47.
#include "common.cuh"

#define ROW_STRIDE 8

#include "rewards.cuh"

// ---- training-time rewards + TD target + Q loss gradient ------------------------------------------
// One warp per row.  rewards r = [r_acc, r_div, r_nov] (SMORL, n_q = 3) or [r] (SQN, n_q = 1).
__global__ void __launch_bounds__(256) td_kernel(int B, int L, int N, int V, int n_q, const float *__restrict__ r_acc,
                                                 const uint8_t *__restrict__ is_end,
                                                 const float *__restrict__ q_sa, const float *__restrict__ q_boot,
                                                 const int64_t *__restrict__ s, const int64_t *__restrict__ div_lens,
                                                 const int32_t *__restrict__ row_ids, rec_train_hparams hp,
                                                 float alpha_eff, float *__restrict__ dq,
                                                 float *__restrict__ q_loss_rows, float *__restrict__ rewards) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float r[3] = {r_acc[b], 0.f, 0.f};
  if (n_q == 3) {
    const int32_t *ids = row_ids + (int64_t)b * REC_MAX_TOPK;
    int last = last_action_of(s, div_lens, b, L, N, hp.pad_pos_end);
    r[1] = diversity_reward_warp(hp.div_emb, hp.div_dim, last, ids, hp.topk_div, hp.out_to_in, N, lane, V);
    float nov = 0.f;
    for (int j = 0; j < hp.topk_nov; ++j) nov += ((unsigned)ids[j] < (unsigned)V && hp.unpopular[ids[j]]) ? hp.nov_reward : 0.f;
    r[2] = nov / (float)hp.topk_nov;
  }
  if (lane == 0) {
    const bool end = is_end[b] != 0;
    float loss = 0.f;
    for (int j = 0; j < n_q; ++j) {
      float boot = end ? 0.f : q_boot[b * 3 + j];
      float y = r[j] + hp.gamma * boot;
      float diff = y - q_sa[b * 3 + j];
      float w = (n_q == 3) ? hp.q_weights[j] : 1.f;
      loss += diff * diff * w;
      dq[b * 3 + j] = alpha_eff * w * 2.f * (-diff) / (float)B;
      rewards[b * 3 + j] = r[j];
    }
    q_loss_rows[b] = loss;
  }
}

// out[0] = mean(ce_row), out[1] = mean(q_loss_rows): single CTA, fixed-order tree (deterministic)
__global__ void __launch_bounds__(256) loss_reduce_kernel(const float *__restrict__ row_stats,
                                                          const float *__restrict__ q_loss_rows, int B,
                                                          float *__restrict__ out) {
  __shared__ float sh[2][256];
  float a = 0.f, q = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) {
    a += row_stats[(int64_t)b * ROW_STRIDE + 4];
    if (q_loss_rows) q += q_loss_rows[b];
  }
  sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = sh[0][0] / (float)B; out[1] = sh[1][0] / (float)B; }
}

// ---- dropout on the final GRU state (BidirGRU4Rec/model.py:60,93) ---------------------------------------
// keep = injected mask, or a counter-based draw from (seed, Adam step of this net, element): no RNG state to
// carry, the step counter lives in device memory so that a replayed graph draws a fresh mask every step.
__device__ __forceinline__ float u01_hash(unsigned long long seed, unsigned long long t, unsigned long long i) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (t * 0x100000001B3ull + i + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;   // splitmix64 finaliser
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}
__global__ void dropout_fwd_kernel(float *__restrict__ h, int n, float p, unsigned long long seed,
                                   const long long *__restrict__ d_step, const uint8_t *__restrict__ injected,
                                   uint8_t *__restrict__ mask_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool keep = injected ? injected[i] != 0 : u01_hash(seed, (unsigned long long)*d_step, (unsigned long long)i) >= p;
  mask_out[i] = keep ? 1 : 0;
  h[i] = keep ? h[i] / (1.f - p) : 0.f;
}
__global__ void dropout_bwd_kernel(float *__restrict__ dh, int n, float p, const uint8_t *__restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dh[i] = mask[i] ? dh[i] / (1.f - p) : 0.f;
}
// mask_row0: first row of the saved keep mask that `dh` corresponds to (data-parallel trunk: a rank back-propagates
// only its own rows of the global batch).
int launch_dropout(rec_engine *e, int net_id, float *h, float *dh, int B, const rec_train_hparams *hp, bool backward,
                   int mask_row0) {
  const int n = B * e->D;
  if (backward)
    dropout_bwd_kernel<<<cdiv(n, 256), 256, 0, e->stream>>>(dh, n, hp->dropout_p, e->drop_mask + (size_t)mask_row0 * e->D);
  else dropout_fwd_kernel<<<cdiv(n, 256), 256, 0, e->stream>>>(h, n, hp->dropout_p, hp->dropout_seed, e->d_step + net_id,
                                                              hp->dropout_mask, e->drop_mask);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

// ---- evaluation metrics ----------------------------------------------------------------------------
// One warp per row -> per-row metric record rowm[b][0..M): hits[n_k], ndcg[n_k], reps[n_k], div, nov, ce
__global__ void __launch_bounds__(256) eval_rows_kernel(int B, int L, int N, int V, const int64_t *__restrict__ s,
                                                        const int64_t *__restrict__ a,
                                                        const int64_t *__restrict__ lens,
                                                        const int32_t *__restrict__ row_ids,
                                                        const float *__restrict__ row_stats, rec_eval_opts o,
                                                        int kmax, uint32_t *__restrict__ cov_bits, int cov_words,
                                                        double *__restrict__ rowm, int M) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int32_t *ids = row_ids + (int64_t)b * REC_MAX_TOPK;
  int my = lane < kmax ? ids[lane] : -1;
  if (my >= V) my = -1;  // exhausted top-k slot (fewer than kmax actions exist)
  double *out = rowm + (int64_t)b * M;
  // HR / NDCG: rank of the true action inside the top-k list
  const int64_t truth = a[b];
  unsigned hit = __ballot_sync(0xffffffffu, lane < kmax && (int64_t)my == truth);
  int pos = hit ? (__ffs(hit) - 1) : 1 << 30;
  // repetitions: matches between every state token and the (remapped) top-k ids
  int64_t mapped = my;
  if (my >= 0 && o.out_to_in) mapped = o.out_to_in[my];
  int reps[REC_MAX_KLIST];
#pragma unroll
  for (int i = 0; i < REC_MAX_KLIST; ++i) reps[i] = 0;
  for (int t = 0; t < L; ++t) {
    int64_t sv = s[(int64_t)b * L + t];
    unsigned mm = __ballot_sync(0xffffffffu, my >= 0 && mapped == sv);
#pragma unroll
    for (int i = 0; i < REC_MAX_KLIST; ++i)
      if (i < o.n_k) reps[i] += __popc(mm & (o.ks[i] >= 32 ? 0xffffffffu : ((1u << o.ks[i]) - 1u)));
  }
  // coverage bitmaps (integer OR: order independent)
  for (int i = 0; i < o.n_cov; ++i)
    if (lane < o.cov_ks[i] && my >= 0) atomicOr(cov_bits + (int64_t)i * cov_words + (my >> 5), 1u << (my & 31));
  // diversity / novelty
  float div = 0.f;
  if (o.div_emb) {
    int last = last_action_of(s, lens, b, L, N, o.pad_pos_end);
    div = diversity_reward_warp(o.div_emb, o.div_dim, last, ids, o.topk_div, o.out_to_in, N, lane, V);
  }
  if (lane == 0) {
    for (int i = 0; i < o.n_k; ++i) {
      bool in = pos < o.ks[i];
      out[i] = in ? 1.0 : 0.0;
      out[o.n_k + i] = in ? 1.0 / log2((double)pos + 2.0) : 0.0;
      out[2 * o.n_k + i] = (double)reps[i];
    }
    double nov = 0.0;
    if (o.unpopular) {
      int cnt = 0;
      for (int j = 0; j < o.topk_nov; ++j) cnt += (ids[j] >= 0 && ids[j] < V && o.unpopular[ids[j]]) ? 1 : 0;
      nov = (double)cnt * (double)o.nov_reward / (double)o.topk_nov;
    }
    out[3 * o.n_k] = (double)div;
    out[3 * o.n_k + 1] = nov;
    out[3 * o.n_k + 2] = (double)row_stats[(int64_t)b * ROW_STRIDE + 4];
  }
}

// column sums of rowm in a fixed order, added to the sweep accumulators. grid = M, block 256.
__global__ void __launch_bounds__(256) eval_reduce_kernel(const double *__restrict__ rowm, int B, int M, int n_k,
                                                          rec_eval_accum acc) {
  __shared__ double sh[256];
  const int m = blockIdx.x;
  double v = 0.0;
  for (int b = threadIdx.x; b < B; b += 256) v += rowm[(int64_t)b * M + m];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double tot = sh[0];
    if (m < n_k) acc.hits[m] += tot;
    else if (m < 2 * n_k) acc.ndcg[m - n_k] += tot;
    else if (m < 3 * n_k) acc.reps[m - 2 * n_k] += tot;
    else if (m == 3 * n_k) acc.div_sum[0] += tot;
    else if (m == 3 * n_k + 1) acc.nov_sum[0] += tot;
    else acc.loss_sum[0] += tot / (double)B;  // mean of batch means (eval_protocol.py:182,250)
  }
}

__global__ void copy_topk_kernel(const int32_t *__restrict__ row_ids, const float *__restrict__ row_topv, int B,
                                 int kmax, int32_t *__restrict__ ids_out, float *__restrict__ sc_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * kmax) return;
  int b = i / kmax, j = i - b * kmax;
  if (ids_out) ids_out[i] = row_ids[(int64_t)b * REC_MAX_TOPK + j];
  if (sc_out) sc_out[i] = row_topv[(int64_t)b * REC_MAX_TOPK + j];
}

// ------------------------------------------------------------------------------------------------
int launch_td(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp, int n_q, float alpha_eff,
              float *q_loss_rows) {
  td_kernel<<<cdiv(b->B, 8), 256, 0, e->stream>>>(b->B, e->cfg.state_size, e->cfg.item_num, e->cfg.action_dim, n_q, b->r, b->is_end, e->q_sa,
                                                 e->q_boot, b->s, b->true_next_len, e->row_ids, *hp, alpha_eff, e->dq,
                                                 q_loss_rows, e->rewards);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

int launch_loss_reduce(rec_engine *e, int B, const float *q_loss_rows, float *out) {
  loss_reduce_kernel<<<1, 256, 0, e->stream>>>(e->row_stats, q_loss_rows, B, out);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

int launch_eval_metrics(rec_engine *e, const rec_batch *b, const rec_eval_opts *o, int kmax,
                        const rec_eval_accum *acc, double *rowm, int32_t *topk_ids, float *topk_scores) {
  const int M = 3 * o->n_k + 3;
  const int cov_words = (e->cfg.action_dim + 31) / 32;
  eval_rows_kernel<<<cdiv(b->B, 8), 256, 0, e->stream>>>(b->B, e->cfg.state_size, e->cfg.item_num, e->cfg.action_dim, b->s,
                                                        b->a, b->true_len, e->row_ids, e->row_stats, *o, kmax,
                                                        acc->cov_bits, cov_words, rowm, M);
  REC_LAUNCH_CHECK(e);
  eval_reduce_kernel<<<M, 256, 0, e->stream>>>(rowm, b->B, M, o->n_k, *acc);
  REC_LAUNCH_CHECK(e);
  if (topk_ids || topk_scores) {
    copy_topk_kernel<<<cdiv(b->B * kmax, 256), 256, 0, e->stream>>>(e->row_ids, e->row_topv, b->B, kmax, topk_ids,
                                                                   topk_scores);
    REC_LAUNCH_CHECK(e);
  }
  return REC_OK;
}

// ---- multi-GPU plumbing: packed batch <-> field arrays -------------------------------------------------
// Packed layout of one rank's batch (bytes): int64 s[B,L] | s_next[B,L] | a[B] | true_len[B] | true_next_len[B]
// | float r[B] | uint8 is_end[B], padded to 16.  One all-gather of this buffer replaces seven.
__global__ void pack_batch_kernel(rec_batch b, int L, uint8_t *__restrict__ out) {
  const int B = b.B;
  int64_t *o64 = reinterpret_cast<int64_t *>(out);
  const int n64 = B * (2 * L + 3);
  float *of = reinterpret_cast<float *>(out + (size_t)n64 * 8);
  uint8_t *oe = out + (size_t)n64 * 8 + (size_t)B * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n64 + 2 * B; i += gridDim.x * blockDim.x) {
    if (i < B * L) o64[i] = b.s[i];
    else if (i < 2 * B * L) o64[i] = b.s_next ? b.s_next[i - B * L] : 0;
    else if (i < 2 * B * L + B) o64[i] = b.a[i - 2 * B * L];
    else if (i < 2 * B * L + 2 * B) o64[i] = b.true_len[i - 2 * B * L - B];
    else if (i < n64) o64[i] = b.true_next_len ? b.true_next_len[i - 2 * B * L - 2 * B] : 0;
    else if (i < n64 + B) of[i - n64] = b.r ? b.r[i - n64] : 0.f;
    else oe[i - n64 - B] = b.is_end ? b.is_end[i - n64 - B] : 0;
  }
}

// gathered[G][stride bytes] (each segment in the packed layout with Bl rows) -> field arrays of G*Bl rows
__global__ void unpack_batch_kernel(const uint8_t *__restrict__ gathered, int G, int Bl, int L, size_t stride,
                                    int64_t *__restrict__ s, int64_t *__restrict__ s_next, int64_t *__restrict__ a,
                                    int64_t *__restrict__ ln, int64_t *__restrict__ nl, float *__restrict__ r,
                                    uint8_t *__restrict__ e) {
  const int n64 = Bl * (2 * L + 3);
  const int per = n64 + 2 * Bl;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per * G; i += gridDim.x * blockDim.x) {
    const int g = i / per, j = i - g * per;
    const uint8_t *seg = gathered + (size_t)g * stride;
    const int64_t *i64 = reinterpret_cast<const int64_t *>(seg);
    if (j < Bl * L) s[(int64_t)g * Bl * L + j] = i64[j];
    else if (j < 2 * Bl * L) s_next[(int64_t)g * Bl * L + (j - Bl * L)] = i64[j];
    else if (j < 2 * Bl * L + Bl) a[g * Bl + (j - 2 * Bl * L)] = i64[j];
    else if (j < 2 * Bl * L + 2 * Bl) ln[g * Bl + (j - 2 * Bl * L - Bl)] = i64[j];
    else if (j < n64) nl[g * Bl + (j - 2 * Bl * L - 2 * Bl)] = i64[j];
    else if (j < n64 + Bl) r[g * Bl + (j - n64)] = reinterpret_cast<const float *>(seg + (size_t)n64 * 8)[j - n64];
    else e[g * Bl + (j - n64 - Bl)] = seg[(size_t)n64 * 8 + (size_t)Bl * 4 + (j - n64 - Bl)];
  }
}

// Replay-buffer sampling on the device (ikea/data_utils/replay_buffer.py:65-74 + the DataLoader's collate):
// row idx[b] of every column of the device-resident buffer -> row b of the batch.  One thread per 8-byte element of
// the state rows, the scalar columns ride along; indices are clamped into [0, n_rows).
__global__ void gather_batch_kernel(rec_batch col, int64_t n_rows, const int64_t *__restrict__ idx, int B, int L,
                                    int64_t *__restrict__ o_s, int64_t *__restrict__ o_sn, int64_t *__restrict__ o_a,
                                    int64_t *__restrict__ o_len, int64_t *__restrict__ o_nlen, float *__restrict__ o_r,
                                    uint8_t *__restrict__ o_end) {
  const int n = B * L;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int b = i / L, t = i - b * L;
    int64_t row = idx[b];
    row = row < 0 ? 0 : (row >= n_rows ? n_rows - 1 : row);
    o_s[i] = col.s[row * L + t];
    if (col.s_next && o_sn) o_sn[i] = col.s_next[row * L + t];
    if (t == 0) {
      o_a[b] = col.a[row];
      o_len[b] = col.true_len[row];
      if (col.true_next_len && o_nlen) o_nlen[b] = col.true_next_len[row];
      if (col.r && o_r) o_r[b] = col.r[row];
      if (col.is_end && o_end) o_end[b] = col.is_end[row];
    }
  }
}
int launch_gather_batch(rec_engine *e, const rec_batch *columns, int64_t n_rows, const int64_t *idx, int B,
                        const rec_batch *out) {
  const int L = e->cfg.state_size, n = B * L;
  gather_batch_kernel<<<cdiv(n, 256), 256, 0, e->stream>>>(*columns, n_rows, idx, B, L, (int64_t *)out->s, (int64_t *)out->s_next,
                                                          (int64_t *)out->a, (int64_t *)out->true_len,
                                                          (int64_t *)out->true_next_len, (float *)out->r, (uint8_t *)out->is_end);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

// Replay-buffer construction from a raw event log (recommenders/data_utils/preprocessing.py:5-29,143-170,199-268):
// one row per event; thread (row, t) finds the row's session by binary search in the CSR offsets and writes element t
// of the padded sliding windows `state` (the <= L items before the event) and `next_state` (... including it).
__global__ void build_rows_kernel(const int64_t *__restrict__ off, int64_t n_sessions, const int64_t *__restrict__ items,
                                  const float *__restrict__ rewards, int64_t n, int L, int64_t pad_id, int pad_end,
                                  int64_t *__restrict__ o_s, int64_t *__restrict__ o_sn, int64_t *__restrict__ o_a,
                                  int64_t *__restrict__ o_len, int64_t *__restrict__ o_nlen, float *__restrict__ o_r,
                                  uint8_t *__restrict__ o_end) {
  const int64_t total = n * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / L;
    const int t = (int)(i - row * L);
    int64_t lo = 0, hi = n_sessions;  // off[lo] <= row < off[hi]
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (off[mid] <= row) lo = mid; else hi = mid;
    }
    const int64_t start = off[lo], stop = off[lo + 1];
    const int64_t nb = row - start;  // n_items_bef
    {
      const int64_t cnt = nb < L ? nb : L, first = start + nb - cnt, padn = L - cnt;
      int64_t v = pad_id;
      if (pad_end) { if (t < cnt) v = items[first + t]; }
      else if (t >= padn) v = items[first + t - padn];
      o_s[i] = v;
    }
    {
      const int64_t cnt = nb + 1 < L ? nb + 1 : L, first = start + nb + 1 - cnt, padn = L - cnt;
      int64_t v = pad_id;
      if (pad_end) { if (t < cnt) v = items[first + t]; }
      else if (t >= padn) v = items[first + t - padn];
      o_sn[i] = v;
    }
    if (t == 0) {
      o_a[row] = items[row];
      o_len[row] = nb < 1 ? 1 : (nb > L ? L : nb);
      o_nlen[row] = nb + 1 > L ? L : nb + 1;
      o_end[row] = row == stop - 1 ? 1 : 0;
      if (o_r) o_r[row] = rewards ? rewards[row] : 0.f;
    }
  }
}
int launch_build_rows(rec_engine *e, const int64_t *off, int64_t n_sessions, const int64_t *items, const float *rewards,
                      int64_t n, int64_t pad_id, int pad_end, const rec_batch *out) {
  const int L = e->cfg.state_size;
  const int64_t total = n * L;
  const int blocks = (int)(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  build_rows_kernel<<<blocks, 256, 0, e->stream>>>(off, n_sessions, items, rewards, n, L, pad_id, pad_end, (int64_t *)out->s,
                                                  (int64_t *)out->s_next, (int64_t *)out->a, (int64_t *)out->true_len,
                                                  (int64_t *)out->true_next_len, (float *)out->r, (uint8_t *)out->is_end);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

int launch_pack_batch(rec_engine *e, const rec_batch *b, uint8_t *out) {
  const int n = b->B * (2 * e->cfg.state_size + 5);
  pack_batch_kernel<<<cdiv(n, 256), 256, 0, e->stream>>>(*b, e->cfg.state_size, out);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}
int launch_unpack_batch(rec_engine *e, const uint8_t *gathered, int G, int Bl, size_t stride, const rec_batch *out) {
  const int n = G * Bl * (2 * e->cfg.state_size + 5);
  unpack_batch_kernel<<<cdiv(n, 256), 256, 0, e->stream>>>(gathered, G, Bl, e->cfg.state_size, stride, (int64_t *)out->s,
                                                          (int64_t *)out->s_next, (int64_t *)out->a, (int64_t *)out->true_len,
                                                          (int64_t *)out->true_next_len, (float *)out->r, (uint8_t *)out->is_end);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}
