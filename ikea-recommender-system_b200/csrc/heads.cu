// Full-vocabulary heads, CUDA-core (fp32 FFMA) path: generic in B, D and V.
//
// Replaces nn.Linear heads + CrossEntropyLoss + argmax/gather/topk of the reference
// (models/SQN/sqn_gru.py:78-85,107-110,221-242; models/SMORL/smorl_gru.py:85-137,272-295;
// utils/tensor_operations.py:4-84; evaluate/*.py `torch.topk` call sites).  Logits are never
// written to HBM on the hot path: every consumer is an epilogue of the tile GEMM
//   * head_stats_kernel  : online (max, sum-exp) + target logit, running top-k (score desc, id asc),
//                          running argmax of sum_h w_h Q_h  -> per-(vocab split,row) partials
//   * head_merge_kernel  : combines the partials of all splits (and, multi-GPU, all shards)
//   * head_bwd_adam_kernel: recomputes the tile, forms dlogits = (softmax - onehot)/B on the fly,
//                          dW/db tile + dh partial, and applies Adam to the tile in the same pass
//                          (W, m, v are read once and written once per step: 24 B/param).
#include "common.cuh"
#include "rewards.cuh"

#define TM 64   // batch rows per tile
#define TN 64   // vocabulary rows per tile
#define KC 32   // k-chunk of the logits GEMM
#define KP 36   // padded k stride (floats) -> conflict-free LDS.128
#define PART_TOPK_OFF 5

// ---- tile GEMM: acc[a][c] += sum_k A[ty+16a][k] * B[tx+16c][k] over one staged chunk ------------
__device__ __forceinline__ void tile_fma(const float (*As)[KP], const float (*Bs)[KP], int ty, int tx,
                                         float acc[4][4]) {
#pragma unroll
  for (int k4 = 0; k4 < KC / 4; ++k4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      a[i] = *reinterpret_cast<const float4 *>(&As[ty + 16 * i][k4 * 4]);
      b[i] = *reinterpret_cast<const float4 *>(&Bs[tx + 16 * i][k4 * 4]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
        acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
        acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
        acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
      }
  }
}

// stage rows [r0, r0+64) x cols [k0, k0+32) of a row-major [nrows, ld] matrix (zero fill outside)
__device__ __forceinline__ void stage_chunk(float (*dst)[KP], const float *__restrict__ src, int r0, int nrows,
                                            int ld, int k0, int kmax, int tid) {
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    int e = tid + q * 256;
    int r = e >> 3, k4 = e & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    int k = k0 + k4 * 4;
    if (r0 + r < nrows && k < kmax) v = __ldg(reinterpret_cast<const float4 *>(src + (int64_t)(r0 + r) * ld + k));
    *reinterpret_cast<float4 *>(&dst[r][k4 * 4]) = v;
  }
}

// logits tile of one head: acc = h[bb*64.., :] . W[v0.., :]^T  (D multiple of 4)
__device__ __forceinline__ void logits_tile(float (*As)[KP], float (*Bs)[KP], const float *__restrict__ h, int b0,
                                            int B, const float *__restrict__ W, int v0, int Vloc, int D, int tid,
                                            int ty, int tx, float acc[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < D; k0 += KC) {
    stage_chunk(As, h, b0, B, D, k0, D, tid);
    stage_chunk(Bs, W, v0, Vloc, D, k0, D, tid);
    __syncthreads();
    tile_fma(As, Bs, ty, tx, acc);
    __syncthreads();
  }
}

struct HeadPtrs {
  const float *w[REC_MAX_HEADS];
  const float *b[REC_MAX_HEADS];
};

// ------------------------------------------------------------------------------------------------
// Statistics pass.  grid = (n_split, ceil(B/64)), block 256.  CTA (sp, bb) walks vocabulary tiles
// [t_lo, t_hi) of the local shard for batch block bb and keeps running per-row results.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_stats_kernel(HeadPtrs hp, const float *__restrict__ h, int B, int D,
                                                         int Vloc, int vocab_lo, int n_tiles, int do_stats,
                                                         int stats_head, const int64_t *__restrict__ target,
                                                         int topk, int n_arg, float w0, float w1, float w2,
                                                         float *__restrict__ part, int part_stride) {
  __shared__ __align__(16) float As[TM][KP];
  __shared__ __align__(16) float Bs[TN][KP];
  __shared__ float Ct[TM][TN + 1];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
  const int sp = blockIdx.x, n_split = gridDim.x, bb = blockIdx.y;
  const int b0 = bb * TM;
  const int per = (n_tiles + n_split - 1) / n_split;
  const int t_lo = sp * per, t_hi = min(n_tiles, t_lo + per);

  float m_run[4], s_run[4], tgt[4], av[4];
  int ai[4];
  int64_t trow[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m_run[a] = REC_NEG_INF; s_run[a] = 0.f; tgt[a] = REC_NEG_INF; av[a] = REC_NEG_INF; ai[a] = 0x7fffffff;
    int row = b0 + ty + 16 * a;
    trow[a] = (do_stats && target && row < B) ? target[row] - vocab_lo : -1;
  }
  float lv[8];
  int li[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) { lv[r] = REC_NEG_INF; li[r] = 0x7fffffff; }
  const float wq[3] = {w0, w1, w2};

  for (int t = t_lo; t < t_hi; ++t) {
    const int v0 = t * TN;
    float acc[4][4];
    if (do_stats || topk > 0) {
      logits_tile(As, Bs, h, b0, B, hp.w[stats_head], v0, Vloc, D, tid, ty, tx, acc);
      float bias[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int col = v0 + tx + 16 * c;
        bias[c] = col < Vloc ? __ldg(hp.b[stats_head] + col) : 0.f;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float l[4];
        float tmax = REC_NEG_INF;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          int col = v0 + tx + 16 * c;
          l[c] = col < Vloc ? acc[a][c] + bias[c] : REC_NEG_INF;
          tmax = fmaxf(tmax, l[c]);
          if ((int64_t)col == trow[a]) tgt[a] = l[c];
          if (topk > 0) Ct[ty + 16 * a][tx + 16 * c] = l[c];
        }
        if (do_stats) {
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
          float nm = fmaxf(m_run[a], tmax);
          float ps = 0.f;
#pragma unroll
          for (int c = 0; c < 4; ++c) ps += (l[c] > REC_NEG_INF) ? __expf(l[c] - nm) : 0.f;
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
          s_run[a] = s_run[a] * __expf(m_run[a] - nm) + ps;
          m_run[a] = nm;
        }
      }
      if (topk > 0) {
        __syncthreads();
        // warp w owns rows 8w..8w+7; lane r of the warp holds the r-th best (value,id) of a row
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const int row = warp * 8 + r;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            int col = v0 + lane + 32 * half;
            float cand = (col < Vloc) ? Ct[row][lane + 32 * half] : REC_NEG_INF;
            float tau = __shfl_sync(0xffffffffu, lv[r], topk - 1);
            unsigned mask = __ballot_sync(0xffffffffu, cand > tau);
            while (mask) {
              int l = __ffs(mask) - 1;
              mask &= mask - 1;
              float cv = __shfl_sync(0xffffffffu, cand, l);
              int ci = vocab_lo + v0 + l + 32 * half;
              tau = __shfl_sync(0xffffffffu, lv[r], topk - 1);
              if (cv > tau) {
                int pos = __popc(__ballot_sync(0xffffffffu, lv[r] >= cv));
                float uv = __shfl_up_sync(0xffffffffu, lv[r], 1);
                int ui = __shfl_up_sync(0xffffffffu, li[r], 1);
                if (lane > pos) { lv[r] = uv; li[r] = ui; }
                if (lane == pos) { lv[r] = cv; li[r] = ci; }
                if (lane >= topk) { lv[r] = REC_NEG_INF; li[r] = 0x7fffffff; }
              }
            }
          }
        }
        __syncthreads();
      }
    }
    if (n_arg > 0) {
      float q[4][4];
      for (int hq = 0; hq < n_arg; ++hq) {
        logits_tile(As, Bs, h, b0, B, hp.w[1 + hq], v0, Vloc, D, tid, ty, tx, acc);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          int col = v0 + tx + 16 * c;
          float bias = col < Vloc ? __ldg(hp.b[1 + hq] + col) : 0.f;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            float val = acc[a][c] + bias;
            if (n_arg > 1) val *= wq[hq];  // get_weighted_q_target: sum_h q_h * w_h in head order
            q[a][c] = (hq == 0) ? val : q[a][c] + val;
          }
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        float bv = REC_NEG_INF;
        int bi = 0x7fffffff;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          int col = v0 + tx + 16 * c;
          if (col < Vloc && better(q[a][c], vocab_lo + col, bv, bi)) { bv = q[a][c]; bi = vocab_lo + col; }
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (better(bv, bi, av[a], ai[a])) { av[a] = bv; ai[a] = bi; }
      }
    }
  }
  // publish partials
  if (tx == 0) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      int row = b0 + ty + 16 * a;
      if (row < B) {
        float *o = part + ((int64_t)sp * B + row) * part_stride;
        o[0] = m_run[a]; o[1] = s_run[a]; o[3] = av[a]; o[4] = __int_as_float(ai[a]);
      }
    }
  }
  // the target logit lives in exactly one lane of the 16: reduce with max
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float tv = tgt[a];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) tv = fmaxf(tv, __shfl_xor_sync(0xffffffffu, tv, o));
    int row = b0 + ty + 16 * a;
    if (tx == 0 && row < B) part[((int64_t)sp * B + row) * part_stride + 2] = tv;
  }
  if (topk > 0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int row = b0 + warp * 8 + r;
      if (row < B && lane < topk) {
        float *o = part + ((int64_t)sp * B + row) * part_stride + PART_TOPK_OFF;
        o[lane] = lv[r];
        o[REC_MAX_TOPK + lane] = __int_as_float(li[r]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Merge of n_split partial records per row (one warp per row).
// row_stats[b] = {lse, target_logit, argmax_val, argmax_id(bits), ce_row, max, sumexp, -}
// ------------------------------------------------------------------------------------------------
#define ROW_STRIDE 8

// ---- fp32 re-score of candidates (one warp per session row, lane c owns candidate c) -----------------------
// exact[c] = sum_k h[k] * W[id_c][k] in ascending k with ONE fp32 accumulator, + bias -- the arithmetic of nn.Linear
// in fp32 up to summation order.  h[k] is broadcast from the lane that loaded it; every lane runs the loop (lanes
// without a candidate compute on row 0 and are ignored by the caller).
struct RescoreSrc {
  const float *w[3], *b[3];   // heads scored: [0] for top-k, [0..n_arg) for the greedy action
  const float *h;             // [B, D] states the candidates were scored on
  int n_arg;                  // >0: greedy-action candidates (sum_j wq[j] * Q_j), else top-k of head w[0]
  float wq[3];
  int approx;                 // 1: records hold approximate (tensor-core) scores -> re-score
  int kpub, apub;             // candidates per record: top-k slots / greedy-action slots
};

// Warp-cooperative: the warp walks the candidates four at a time; for each one every lane loads ONE float4 of the
// candidate's weight row per 128 columns (coalesced, all loads of a group independent -> one memory round trip per
// group instead of a chain of dependent per-lane loads), multiplies it with its float4 of h and the partial sums are
// combined by the fixed butterfly of warp_sum -- a deterministic fp32 evaluation of h.W[id] whose rounding error
// (~1e-7 relative) is that of any fp32 dot product.  `n_cand` is warp-uniform; lane c < n_cand passes its candidate's
// local row `loc` (any valid row for the other lanes) and receives its candidate's score.
template <int NH>
struct RowSrc { const float *p[NH]; };
template <int NH>
__device__ __forceinline__ void exact_rows_coop(const RowSrc<NH> W, int64_t loc, int n_cand,
                                                const float *__restrict__ hrow, int D, int lane, float *out) {
#pragma unroll
  for (int j = 0; j < NH; ++j) out[j] = 0.f;
  for (int c0 = 0; c0 < n_cand; c0 += 4) {
    float part[NH][4];
#pragma unroll
    for (int j = 0; j < NH; ++j)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) part[j][cc] = 0.f;
    int64_t locs[4];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) locs[cc] = __shfl_sync(0xffffffffu, loc, (c0 + cc) & 31);
    for (int k0 = lane * 4; k0 < D; k0 += 128) {
      const float4 hv = *reinterpret_cast<const float4 *>(hrow + k0);
      float4 w4[NH][4];
#pragma unroll
      for (int j = 0; j < NH; ++j)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
          w4[j][cc] = (c0 + cc < n_cand) ? __ldg(reinterpret_cast<const float4 *>(W.p[j] + locs[cc] * D + k0))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < NH; ++j)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          float a = part[j][cc];
          a = fmaf(hv.x, w4[j][cc].x, a); a = fmaf(hv.y, w4[j][cc].y, a);
          a = fmaf(hv.z, w4[j][cc].z, a); a = fmaf(hv.w, w4[j][cc].w, a);
          part[j][cc] = a;
        }
    }
#pragma unroll
    for (int j = 0; j < NH; ++j)
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const float sfull = warp_sum(part[j][cc]);
        if (lane == c0 + cc) out[j] = sfull;
      }
  }
}

// exact score of head (W, bias) for the candidate this lane owns (lanes >= n_cand: unspecified)
__device__ __forceinline__ float exact_row_dot(const float *__restrict__ W, const float *__restrict__ bias, int64_t loc,
                                               int n_cand, const float *__restrict__ hrow, int D, int lane) {
  RowSrc<1> Ws; Ws.p[0] = W;
  float o[1];
  exact_rows_coop<1>(Ws, loc, n_cand, hrow, D, lane, o);
  return o[0] + __ldg(bias + loc);
}

// greedy-action score of candidate `id` in the oracle's order: ((q0*w0 + q1*w1) + q2*w2), plain q0 for one head
__device__ __forceinline__ float exact_arg_score(const RescoreSrc &R, int id, int n_cand, int vocab_lo, int Vloc,
                                                 const float *hrow, int D, int lane) {
  int64_t loc = (int64_t)id - vocab_lo;
  const bool ok = loc >= 0 && loc < Vloc;
  if (!ok) loc = 0;
  float sc = 0.f;
  if (R.n_arg == 3) {
    RowSrc<3> Ws; Ws.p[0] = R.w[0]; Ws.p[1] = R.w[1]; Ws.p[2] = R.w[2];
    float q[3];
    exact_rows_coop<3>(Ws, loc, n_cand, hrow, D, lane, q);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float qj = (q[j] + __ldg(R.b[j] + loc)) * R.wq[j];
      sc = (j == 0) ? qj : sc + qj;
    }
  } else {
    for (int j = 0; j < R.n_arg; ++j) {
      float q = exact_row_dot(R.w[j], R.b[j], loc, n_cand, hrow, D, lane);
      if (R.n_arg > 1) q *= R.wq[j];
      sc = (j == 0) ? q : sc + q;
    }
  }
  return ok ? sc : REC_NEG_INF;
}

// Greedy action from `n_split` records of one row: best REC_ARG_CAND candidates by approximate score, re-scored in
// fp32, best by (score desc, id asc).  All lanes return the winner.
__device__ __forceinline__ void merge_arg_rescored(const float *__restrict__ part, int part_stride, int n_split, int B,
                                                   int row, const RescoreSrc &R, int vocab_lo, int Vloc, int D, int lane,
                                                   float &out_v, int &out_i) {
  const int n = n_split * R.apub;
  float lastv = 3.402823466e+38f, cv = REC_NEG_INF;
  int lasti = -1, ci = 0x7fffffff;
  for (int k = 0; k < REC_ARG_CAND; ++k) {
    float bv = REC_NEG_INF;
    int bi = 0x7fffffff;
    for (int c = lane; c < n; c += 32) {
      const int sp = c / R.apub, j = c - sp * R.apub;
      const float *o = part + ((int64_t)sp * B + row) * part_stride + PART_TOPK_OFF;
      const float v = o[j];
      const int i = __float_as_int(o[REC_MAX_TOPK + j]);
      if (i != 0x7fffffff && better(lastv, lasti, v, i) && better(v, i, bv, bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == k) { cv = bv; ci = bi; }
    lastv = bv; lasti = bi;
  }
  const float ex = exact_arg_score(R, ci == 0x7fffffff ? vocab_lo : ci, REC_ARG_CAND, vocab_lo, Vloc, R.h + (int64_t)row * D, D, lane);
  float bv = (lane < REC_ARG_CAND && ci != 0x7fffffff) ? ex : REC_NEG_INF;
  int bi = (lane < REC_ARG_CAND) ? ci : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  out_v = bv; out_i = bi;
}

// Orders the (<= 32) re-scored candidates of a row, one per lane, by (score desc, id asc) and writes the best
// `topk` of them; slots without a candidate get (-FLT_MAX, 0x7fffffff) like an exhausted selection round.
__device__ __forceinline__ void write_ranked(float v, int i, bool valid, int topk, int lane, int64_t row,
                                             int32_t *__restrict__ row_ids, float *__restrict__ row_topv, float *sm) {
  if (!valid) { v = REC_NEG_INF; i = 0x7fffffff; }
  int rank = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float ov = __shfl_sync(0xffffffffu, v, j);
    const int oi = __shfl_sync(0xffffffffu, i, j);
    rank += better(ov, oi, v, i) ? 1 : 0;
  }
  const int n_valid = __popc(__ballot_sync(0xffffffffu, valid));
  int slot = -1;
  if (valid && rank < topk) slot = rank;
  else if (!valid && lane >= n_valid && lane < topk) slot = lane;
  if (slot >= 0) {
    row_ids[row * REC_MAX_TOPK + slot] = i; row_topv[row * REC_MAX_TOPK + slot] = v;
    if (sm) { sm[PART_TOPK_OFF + slot] = v; sm[PART_TOPK_OFF + REC_MAX_TOPK + slot] = __int_as_float(i); }
  }
}

__global__ void __launch_bounds__(256) head_merge_kernel(const float *__restrict__ part, int part_stride, int n_split,
                                                         int B, int topk, int has_stats, int has_arg,
                                                         float *__restrict__ row_stats, int32_t *__restrict__ row_ids,
                                                         float *__restrict__ row_topv, int32_t *__restrict__ astar,
                                                         float *__restrict__ summary, RescoreSrc R, int D, int vocab_lo,
                                                         int Vloc) {
  // `summary` (optional): the merged result re-packed as ONE record per row in the partial-record
  // format, so that a second merge over the all-gathered summaries of all vocabulary shards
  // (n_split = number of shards) finishes the reduction across GPUs with this same kernel.
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  float *rs = row_stats + (int64_t)row * ROW_STRIDE;
  float *sm = summary ? summary + (int64_t)row * part_stride : nullptr;
  if (has_stats) {
    float m = REC_NEG_INF, tg = REC_NEG_INF;
    for (int sp = lane; sp < n_split; sp += 32) {
      const float *o = part + ((int64_t)sp * B + row) * part_stride;
      m = fmaxf(m, o[0]);
      tg = fmaxf(tg, o[2]);
    }
    m = warp_max(m);
    tg = warp_max(tg);
    float ssum = 0.f;
    for (int sp = lane; sp < n_split; sp += 32) {
      const float *o = part + ((int64_t)sp * B + row) * part_stride;
      if (o[1] > 0.f) ssum += o[1] * __expf(o[0] - m);
    }
    ssum = warp_sum(ssum);
    if (lane == 0) {
      float lse = m + logf(ssum);
      rs[0] = lse; rs[1] = tg; rs[4] = lse - tg; rs[5] = m; rs[6] = ssum;
      if (sm) { sm[0] = m; sm[1] = ssum; sm[2] = tg; }
    }
  }
  if (has_arg) {
    float bv = REC_NEG_INF;
    int bi = 0x7fffffff;
    if (R.approx && R.apub > 0) {
      merge_arg_rescored(part, part_stride, n_split, B, row, R, vocab_lo, Vloc, D, lane, bv, bi);
    } else {
      for (int sp = lane; sp < n_split; sp += 32) {
        const float *o = part + ((int64_t)sp * B + row) * part_stride;
        float v = o[3];
        int i = __float_as_int(o[4]);
        if (better(v, i, bv, bi)) { bv = v; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
    }
    if (lane == 0) {
      rs[2] = bv; rs[3] = __int_as_float(bi); astar[row] = bi;
      if (sm) { sm[3] = bv; sm[4] = __int_as_float(bi); }
    }
  }
  if (topk > 0) {
    // selection rounds over n_split*kpub candidates; "taken" = not after the last pick in the order.
    // Approximate scores: rec_kt(topk) rounds, lane k keeps the k-th pick for the fp32 re-score.
    float lastv = 3.402823466e+38f, cv = REC_NEG_INF;
    int lasti = -1, ci = 0x7fffffff;
    const int kpub = R.kpub, n = n_split * kpub;
    const int kt = R.approx ? rec_kt(topk) : topk;
    for (int k = 0; k < kt; ++k) {
      float bv = REC_NEG_INF;
      int bi = 0x7fffffff;
      for (int c = lane; c < n; c += 32) {
        int sp = c / kpub, j = c - sp * kpub;
        const float *o = part + ((int64_t)sp * B + row) * part_stride + PART_TOPK_OFF;
        float v = o[j];
        int i = __float_as_int(o[REC_MAX_TOPK + j]);
        if (i != 0x7fffffff && better(lastv, lasti, v, i) && better(v, i, bv, bi)) { bv = v; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
      if (R.approx) {
        if (lane == k) { cv = bv; ci = bi; }
      } else if (lane == 0) {
        row_ids[(int64_t)row * REC_MAX_TOPK + k] = bi; row_topv[(int64_t)row * REC_MAX_TOPK + k] = bv;
        if (sm) { sm[PART_TOPK_OFF + k] = bv; sm[PART_TOPK_OFF + REC_MAX_TOPK + k] = __int_as_float(bi); }
      }
      lastv = bv; lasti = bi;
    }
    if (R.approx) {
      const bool valid = lane < kt && ci != 0x7fffffff;
      const float ex = exact_row_dot(R.w[0], R.b[0], valid ? (int64_t)ci - vocab_lo : 0, kt, R.h + (int64_t)row * D, D, lane);
      write_ranked(ex, ci, valid, topk, lane, row, row_ids, row_topv, sm);
    }
  }
}

// Training-shaped merge (n_split <= 160 records per row, top-k <= 2): every lane requests ALL its records up front
// (one memory round trip instead of one per reduction pass), then the same reductions run from registers.
// Results are bit-identical to head_merge_kernel (same per-lane orders, same warp reductions).
__global__ void __launch_bounds__(128) head_merge_small_kernel(const float *__restrict__ part, int part_stride, int n_split,
                                                               int B, int topk, int has_stats, int has_arg,
                                                               float *__restrict__ row_stats, int32_t *__restrict__ row_ids,
                                                               float *__restrict__ row_topv, int32_t *__restrict__ astar,
                                                               float *__restrict__ summary, RescoreSrc R, int D,
                                                               int vocab_lo, int Vloc) {
  constexpr int R5 = 5, KP4 = 4;  // records per lane, candidate slots per record held in registers
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B) return;
  float *rs = row_stats + (int64_t)row * ROW_STRIDE;
  float *sm = summary ? summary + (int64_t)row * part_stride : nullptr;
  const bool arg_cand = has_arg && R.approx && R.apub > 0;
  const int np = arg_cand ? R.apub : (topk > 0 ? R.kpub : 0);  // <= KP4 (checked by the launcher)
  float4 a[R5];
  float a4[R5], tv[R5][KP4];
  int ti[R5][KP4];
#pragma unroll
  for (int r = 0; r < R5; ++r) {
    const int sp = lane + 32 * r;
    a[r] = make_float4(REC_NEG_INF, 0.f, REC_NEG_INF, REC_NEG_INF);
    a4[r] = __int_as_float(0x7fffffff);
#pragma unroll
    for (int j = 0; j < KP4; ++j) { tv[r][j] = REC_NEG_INF; ti[r][j] = 0x7fffffff; }
    if (sp < n_split) {
      const float *o = part + ((int64_t)sp * B + row) * part_stride;
      a[r] = *reinterpret_cast<const float4 *>(o);
      a4[r] = o[4];
#pragma unroll
      for (int j = 0; j < KP4; ++j)
        if (j < np) { tv[r][j] = o[PART_TOPK_OFF + j]; ti[r][j] = __float_as_int(o[PART_TOPK_OFF + REC_MAX_TOPK + j]); }
    }
  }
  // one selection round over the register-held candidates: best entry strictly after (lastv, lasti) in the order
  auto select_next = [&](float lastv, int lasti, float &bv, int &bi) {
    bv = REC_NEG_INF; bi = 0x7fffffff;
#pragma unroll
    for (int r = 0; r < R5; ++r)
#pragma unroll
      for (int j = 0; j < KP4; ++j) {
        const float v = tv[r][j];
        const int i = ti[r][j];
        if (i != 0x7fffffff && better(lastv, lasti, v, i) && better(v, i, bv, bi)) { bv = v; bi = i; }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
  };
  if (has_stats) {
    float m = REC_NEG_INF, tg = REC_NEG_INF;
#pragma unroll
    for (int r = 0; r < R5; ++r) { m = fmaxf(m, a[r].x); tg = fmaxf(tg, a[r].z); }
    m = warp_max(m);
    tg = warp_max(tg);
    float ssum = 0.f;
#pragma unroll
    for (int r = 0; r < R5; ++r)
      if (lane + 32 * r < n_split && a[r].y > 0.f) ssum += a[r].y * __expf(a[r].x - m);
    ssum = warp_sum(ssum);
    if (lane == 0) {
      float lse = m + logf(ssum);
      rs[0] = lse; rs[1] = tg; rs[4] = lse - tg; rs[5] = m; rs[6] = ssum;
      if (sm) { sm[0] = m; sm[1] = ssum; sm[2] = tg; }
    }
  }
  if (has_arg) {
    float bv = REC_NEG_INF;
    int bi = 0x7fffffff;
    if (arg_cand) {
      float lastv = 3.402823466e+38f, cv = REC_NEG_INF;
      int lasti = -1, ci = 0x7fffffff;
      for (int k = 0; k < REC_ARG_CAND; ++k) {
        float sv; int si;
        select_next(lastv, lasti, sv, si);
        if (lane == k) { cv = sv; ci = si; }
        lastv = sv; lasti = si;
      }
      const float ex = exact_arg_score(R, ci == 0x7fffffff ? vocab_lo : ci, REC_ARG_CAND, vocab_lo, Vloc, R.h + (int64_t)row * D, D, lane);
      bv = (lane < REC_ARG_CAND && ci != 0x7fffffff) ? ex : REC_NEG_INF;
      bi = (lane < REC_ARG_CAND) ? ci : 0x7fffffff;
    } else {
#pragma unroll
      for (int r = 0; r < R5; ++r) {
        const float v = a[r].w;
        const int i = __float_as_int(a4[r]);
        if (lane + 32 * r < n_split && better(v, i, bv, bi)) { bv = v; bi = i; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      rs[2] = bv; rs[3] = __int_as_float(bi); astar[row] = bi;
      if (sm) { sm[3] = bv; sm[4] = __int_as_float(bi); }
    }
  }
  if (topk > 0 && !arg_cand) {
    float lastv = 3.402823466e+38f, cv = REC_NEG_INF;
    int lasti = -1, ci = 0x7fffffff;
    const int kt = R.approx ? rec_kt(topk) : topk;
    for (int k = 0; k < kt; ++k) {
      float bv; int bi;
      select_next(lastv, lasti, bv, bi);
      if (R.approx) {
        if (lane == k) { cv = bv; ci = bi; }
      } else if (lane == 0) {
        row_ids[(int64_t)row * REC_MAX_TOPK + k] = bi; row_topv[(int64_t)row * REC_MAX_TOPK + k] = bv;
        if (sm) { sm[PART_TOPK_OFF + k] = bv; sm[PART_TOPK_OFF + REC_MAX_TOPK + k] = __int_as_float(bi); }
      }
      lastv = bv; lasti = bi;
    }
    if (R.approx) {
      const bool valid = lane < kt && ci != 0x7fffffff;
      const float ex = exact_row_dot(R.w[0], R.b[0], valid ? (int64_t)ci - vocab_lo : 0, kt, R.h + (int64_t)row * D, D, lane);
      write_ranked(ex, ci, valid, topk, lane, row, row_ids, row_topv, sm);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Plain logits (API compatibility with `model(s, lengths)`): logits[b, v] = h[b].W[v] + bias[v].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_logits_kernel(const float *__restrict__ W, const float *__restrict__ bias,
                                                          const float *__restrict__ h, int B, int D, int Vloc,
                                                          float *__restrict__ out, int64_t ld) {
  __shared__ __align__(16) float As[TM][KP];
  __shared__ __align__(16) float Bs[TN][KP];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int v0 = blockIdx.x * TN, b0 = blockIdx.y * TM;
  float acc[4][4];
  logits_tile(As, Bs, h, b0, B, W, v0, Vloc, D, tid, ty, tx, acc);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int row = b0 + ty + 16 * a;
    if (row >= B) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int col = v0 + tx + 16 * c;
      if (col < Vloc) out[(int64_t)row * ld + col] = acc[a][c] + __ldg(bias + col);
    }
  }
}

// out[b, j] = h[b] . W_{first+j}[id_b - vocab_lo] + bias (0 when the row is on another shard).
__global__ void __launch_bounds__(256) row_dots_kernel(HeadPtrs hp, const float *__restrict__ h,
                                                       const int64_t *__restrict__ ids64,
                                                       const int32_t *__restrict__ ids32, int B, int D, int Vloc,
                                                       int vocab_lo, int first, int n, float *__restrict__ out,
                                                       int out_stride) {
  const int lane = threadIdx.x & 31;
  const int wi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wi >= B * n) return;
  const int b = wi / n, j = wi - b * n;
  int64_t id = ids64 ? ids64[b] : (int64_t)ids32[b];
  int64_t loc = id - vocab_lo;
  float acc = 0.f;
  if (loc >= 0 && loc < Vloc) {
    const float *wr = hp.w[first + j] + loc * D;
    const float *hr = h + (int64_t)b * D;
    for (int k = lane; k < D; k += 32) acc = fmaf(hr[k], __ldg(wr + k), acc);
    acc = warp_sum(acc);
    acc += __ldg(hp.b[first + j] + loc);
  }
  if (lane == 0) out[(int64_t)b * out_stride + j] = acc;
}

// ------------------------------------------------------------------------------------------------
// Everything between the greedy-action statistics and the backward passes, for one row per warp:
//   merge of the argmax records -> a*;  Q(s,a) on the main net, Q_boot(s',a*) on the bootstrap net;
//   SMORL rewards;  TD target, dq, per-row Q loss;  dh contribution of the Q heads.
// Same arithmetic, in the same order, as head_merge_kernel(has_arg) + row_dots_kernel x2 + td_kernel +
// q_dh_kernel, which stay in use for the phase-split (vocabulary-sharded) step.  One launch instead of
// five on the critical path of the single-GPU step.
// ------------------------------------------------------------------------------------------------
struct QRowArgs {
  const float *part; int part_stride, n_split;
  HeadPtrs main_heads, boot_heads;
  const float *h_main, *h_boot;      // [B, D] final states main(s), boot(s')
  const int64_t *a, *s, *div_lens;
  const float *r_acc; const uint8_t *is_end;
  const int32_t *row_ids;            // merged top-k ids of the supervised logits (SMORL rewards)
  int B, D, L, N, V, Vloc, vocab_lo, n_q;
  float alpha_eff;
  float *row_stats; int32_t *astar;
  float *q_sa, *q_boot, *dq, *q_loss_rows, *rewards, *dh_slice;
  RescoreSrc R;                      // greedy-action candidates: fp32 re-score on main(s') when R.approx
};

__global__ void __launch_bounds__(256) q_rows_fused_kernel(QRowArgs A, rec_train_hparams hp) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= A.B) return;
  const int D = A.D, n_q = A.n_q;
  // (1) greedy action: merge of the per-split argmax records (approximate scores: fp32 re-score of the best few)
  float bv = REC_NEG_INF;
  int bi = 0x7fffffff;
  if (A.R.approx && A.R.apub > 0) {
    merge_arg_rescored(A.part, A.part_stride, A.n_split, A.B, b, A.R, A.vocab_lo, A.Vloc, D, lane, bv, bi);
  } else {
    for (int sp = lane; sp < A.n_split; sp += 32) {
      const float *o = A.part + ((int64_t)sp * A.B + b) * A.part_stride;
      float v = o[3];
      int i = __float_as_int(o[4]);
      if (better(v, i, bv, bi)) { bv = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
  }
  if (lane == 0) {
    float *rs = A.row_stats + (int64_t)b * ROW_STRIDE;
    rs[2] = bv; rs[3] = __int_as_float(bi); A.astar[b] = bi;
  }
  // (2) Q(s,a) and Q_boot(s',a*)
  const int64_t loc_a = A.a[b] - A.vocab_lo, loc_s = (int64_t)bi - A.vocab_lo;
  float qsa[3] = {0.f, 0.f, 0.f}, qbt[3] = {0.f, 0.f, 0.f};
  for (int j = 0; j < n_q; ++j) {
    if (loc_a >= 0 && loc_a < A.Vloc) {
      const float *wr = A.main_heads.w[1 + j] + loc_a * D, *hr = A.h_main + (int64_t)b * D;
      float acc = 0.f;
      for (int k = lane; k < D; k += 32) acc = fmaf(hr[k], __ldg(wr + k), acc);
      qsa[j] = warp_sum(acc) + __ldg(A.main_heads.b[1 + j] + loc_a);
    }
    if (loc_s >= 0 && loc_s < A.Vloc) {
      const float *wr = A.boot_heads.w[1 + j] + loc_s * D, *hr = A.h_boot + (int64_t)b * D;
      float acc = 0.f;
      for (int k = lane; k < D; k += 32) acc = fmaf(hr[k], __ldg(wr + k), acc);
      qbt[j] = warp_sum(acc) + __ldg(A.boot_heads.b[1 + j] + loc_s);
    }
  }
  // (3) rewards
  float r[3] = {A.r_acc[b], 0.f, 0.f};
  if (n_q == 3) {
    const int32_t *ids = A.row_ids + (int64_t)b * REC_MAX_TOPK;
    int last = last_action_of(A.s, A.div_lens, b, A.L, A.N, hp.pad_pos_end);
    r[1] = diversity_reward_warp(hp.div_emb, hp.div_dim, last, ids, hp.topk_div, hp.out_to_in, A.N, lane, A.V);
    float nov = 0.f;
    for (int j = 0; j < hp.topk_nov; ++j) nov += ((unsigned)ids[j] < (unsigned)A.V && hp.unpopular[ids[j]]) ? hp.nov_reward : 0.f;
    r[2] = nov / (float)hp.topk_nov;
  }
  // (4) TD target, dq, per-row loss (every lane computes the same scalars)
  const bool end = A.is_end[b] != 0;
  float loss = 0.f, dq[3] = {0.f, 0.f, 0.f};
  for (int j = 0; j < n_q; ++j) {
    float boot = end ? 0.f : qbt[j];
    float y = r[j] + hp.gamma * boot;
    float diff = y - qsa[j];
    float w = (n_q == 3) ? hp.q_weights[j] : 1.f;
    loss += diff * diff * w;
    dq[j] = A.alpha_eff * w * 2.f * (-diff) / (float)A.B;
  }
  if (lane == 0) {
    for (int j = 0; j < n_q; ++j) {
      A.q_sa[b * 3 + j] = qsa[j]; A.q_boot[b * 3 + j] = qbt[j];
      A.dq[b * 3 + j] = dq[j]; A.rewards[b * 3 + j] = r[j];
    }
    A.q_loss_rows[b] = loss;
  }
  // (5) dh contribution of the Q heads (before Adam touches their weights)
  for (int k = lane; k < D; k += 32) {
    float acc = 0.f;
    if (loc_a >= 0 && loc_a < A.Vloc)
      for (int j = 0; j < n_q; ++j) acc = fmaf(dq[j], __ldg(A.main_heads.w[1 + j] + loc_a * D + k), acc);
    A.dh_slice[(int64_t)b * D + k] = acc;
  }
}

// Same computation with the memory round trips collapsed (D = 32 DPL <= 128, <= 160 records per row): the four
// independent dependency chains (records -> a* -> bootstrap rows | a_b -> main rows | lengths -> s -> E_div row |
// top-k ids) are requested side by side instead of one after the other, and the main-head rows loaded for Q(s,a)
// are reused for the dh slice.  Arithmetic orders are those of q_rows_fused_kernel (bit-identical results).
template <int DPL>
__global__ void __launch_bounds__(128) q_rows_fused_small_kernel(QRowArgs A, rec_train_hparams hp) {
  constexpr int R = 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= A.B) return;
  const int D = A.D, n_q = A.n_q;
  // round trip 1: records, a_b, lengths, reward inputs
  const bool cand = A.R.approx && A.R.apub > 0;  // records carry two approximate candidates each (tensor-core pass)
  float rv[R], rv2[R];
  int ri[R], ri2[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int sp = lane + 32 * r;
    rv[r] = REC_NEG_INF; ri[r] = 0x7fffffff; rv2[r] = REC_NEG_INF; ri2[r] = 0x7fffffff;
    if (sp < A.n_split) {
      const float *o = A.part + ((int64_t)sp * A.B + b) * A.part_stride;
      rv[r] = o[3]; ri[r] = __float_as_int(o[4]);
      if (cand) { rv2[r] = o[PART_TOPK_OFF + 1]; ri2[r] = __float_as_int(o[PART_TOPK_OFF + REC_MAX_TOPK + 1]); }
    }
  }
  const int64_t loc_a = A.a[b] - A.vocab_lo;
  const float r_acc = A.r_acc[b];
  const bool end = A.is_end[b] != 0;
  int64_t dlen = 1;
  if (n_q == 3 && hp.pad_pos_end) dlen = A.div_lens[b];
  const float *hm = A.h_main + (int64_t)b * D, *hb = A.h_boot + (int64_t)b * D;
  float hmr[DPL], hbr[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) { hmr[i] = hm[lane + 32 * i]; hbr[i] = hb[lane + 32 * i]; }
  // round trip 2 (a_b known): main-head rows; (lengths known): last item of s
  const bool a_here = loc_a >= 0 && loc_a < A.Vloc;
  float wm[3][DPL], bm_[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < DPL; ++i) wm[j][i] = (j < n_q && a_here) ? __ldg(A.main_heads.w[1 + j] + loc_a * D + lane + 32 * i) : 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) if (j < n_q && a_here) bm_[j] = __ldg(A.main_heads.b[1 + j] + loc_a);
  int last = 0;
  if (n_q == 3) {
    int64_t it;
    if (hp.pad_pos_end) {
      int64_t l = dlen < 1 ? 1 : (dlen > A.L ? A.L : dlen);
      it = A.s[(int64_t)b * A.L + (l - 1)];
    } else {
      it = A.s[(int64_t)b * A.L + (A.L - 1)];
    }
    last = (int)(it < 0 ? 0 : (it > A.N ? A.N : it));
  }
  // (1) greedy action
  float bv = REC_NEG_INF;
  int bi = 0x7fffffff;
  if (cand) {
    // best REC_ARG_CAND candidates by approximate score -> exact fp32 score on main(s') -> (score desc, id asc)
    float lastv = 3.402823466e+38f, cv = REC_NEG_INF;
    int lasti = -1, ci = 0x7fffffff;
    for (int k = 0; k < REC_ARG_CAND; ++k) {
      float sv = REC_NEG_INF;
      int si = 0x7fffffff;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (ri[r] != 0x7fffffff && better(lastv, lasti, rv[r], ri[r]) && better(rv[r], ri[r], sv, si)) { sv = rv[r]; si = ri[r]; }
        if (ri2[r] != 0x7fffffff && better(lastv, lasti, rv2[r], ri2[r]) && better(rv2[r], ri2[r], sv, si)) { sv = rv2[r]; si = ri2[r]; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, sv, o);
        int oi = __shfl_xor_sync(0xffffffffu, si, o);
        if (better(ov, oi, sv, si)) { sv = ov; si = oi; }
      }
      if (lane == k) { cv = sv; ci = si; }
      lastv = sv; lasti = si;
    }
    const float ex = exact_arg_score(A.R, ci == 0x7fffffff ? A.vocab_lo : ci, REC_ARG_CAND, A.vocab_lo, A.Vloc, A.R.h + (int64_t)b * D, D, lane);
    bv = (lane < REC_ARG_CAND && ci != 0x7fffffff) ? ex : REC_NEG_INF;
    bi = (lane < REC_ARG_CAND) ? ci : 0x7fffffff;
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (lane + 32 * r < A.n_split && better(rv[r], ri[r], bv, bi)) { bv = rv[r]; bi = ri[r]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
  }
  // round trip 3 (a* known): bootstrap rows
  const int64_t loc_s = (int64_t)bi - A.vocab_lo;
  const bool s_here = loc_s >= 0 && loc_s < A.Vloc;
  float wb[3][DPL], bb_[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int i = 0; i < DPL; ++i) wb[j][i] = (j < n_q && s_here) ? __ldg(A.boot_heads.w[1 + j] + loc_s * D + lane + 32 * i) : 0.f;
#pragma unroll
  for (int j = 0; j < 3; ++j) if (j < n_q && s_here) bb_[j] = __ldg(A.boot_heads.b[1 + j] + loc_s);
  // (3) rewards (their loads overlap the bootstrap rows in flight)
  float r[3] = {r_acc, 0.f, 0.f};
  if (n_q == 3) {
    const int32_t *ids = A.row_ids + (int64_t)b * REC_MAX_TOPK;
    r[1] = diversity_reward_warp(hp.div_emb, hp.div_dim, last, ids, hp.topk_div, hp.out_to_in, A.N, lane, A.V);
    float nov = 0.f;
    for (int j = 0; j < hp.topk_nov; ++j) nov += ((unsigned)ids[j] < (unsigned)A.V && hp.unpopular[ids[j]]) ? hp.nov_reward : 0.f;
    r[2] = nov / (float)hp.topk_nov;
  }
  // (2) Q(s,a) and Q_boot(s',a*)
  float qsa[3] = {0.f, 0.f, 0.f}, qbt[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j >= n_q) continue;
    if (a_here) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < DPL; ++i) acc = fmaf(hmr[i], wm[j][i], acc);
      qsa[j] = warp_sum(acc) + bm_[j];
    }
    if (s_here) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < DPL; ++i) acc = fmaf(hbr[i], wb[j][i], acc);
      qbt[j] = warp_sum(acc) + bb_[j];
    }
  }
  // (4) TD target, dq, per-row loss (every lane computes the same scalars)
  float loss = 0.f, dq[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j >= n_q) continue;
    float boot = end ? 0.f : qbt[j];
    float y = r[j] + hp.gamma * boot;
    float diff = y - qsa[j];
    float w = (n_q == 3) ? hp.q_weights[j] : 1.f;
    loss += diff * diff * w;
    dq[j] = A.alpha_eff * w * 2.f * (-diff) / (float)A.B;
  }
  if (lane == 0) {
    float *rs = A.row_stats + (int64_t)b * ROW_STRIDE;
    rs[2] = bv; rs[3] = __int_as_float(bi); A.astar[b] = bi;
    for (int j = 0; j < n_q; ++j) {
      A.q_sa[b * 3 + j] = qsa[j]; A.q_boot[b * 3 + j] = qbt[j];
      A.dq[b * 3 + j] = dq[j]; A.rewards[b * 3 + j] = r[j];
    }
    A.q_loss_rows[b] = loss;
  }
  // (5) dh contribution of the Q heads (before Adam touches their weights): the rows are still in registers
#pragma unroll
  for (int i = 0; i < DPL; ++i) {
    float acc = 0.f;
    if (a_here) {
#pragma unroll
      for (int j = 0; j < 3; ++j) if (j < n_q) acc = fmaf(dq[j], wm[j][i], acc);
    }
    A.dh_slice[(int64_t)b * D + lane + 32 * i] = acc;
  }
}

// dh_q[b, :] = sum_j dq[b, j] * W_{1+j}[a_b]  (Q heads only touch row a_b) -- runs BEFORE Adam.
__global__ void __launch_bounds__(256) q_dh_kernel(HeadPtrs hp, const int64_t *__restrict__ a,
                                                   const float *__restrict__ dq, int B, int D, int Vloc, int vocab_lo,
                                                   int n_q, float *__restrict__ dh_slice) {
  const int b = blockIdx.x;
  int64_t loc = a[b] - vocab_lo;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    float acc = 0.f;
    if (loc >= 0 && loc < Vloc)
      for (int j = 0; j < n_q; ++j) acc = fmaf(dq[b * 3 + j], __ldg(hp.w[1 + j] + loc * D + k), acc);
    dh_slice[(int64_t)b * D + k] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// Backward + Adam.  grid = (n_cta, n_heads), block 256, dynamic smem.  blockIdx.y = head.
//   head 0     : dense path (recompute logits, dlogits, dW/db, dh partial)
//   heads >= 1 : sparse path (only rows a_b carry gradient dq[b,head-1] * h[b])
// Every CTA then applies Adam to the [64, D] weight tile + bias it owns.  Persistent over tiles.
// ------------------------------------------------------------------------------------------------
struct HeadTrainPtrs {
  float *w[REC_MAX_HEADS], *wm[REC_MAX_HEADS], *wv[REC_MAX_HEADS];
  float *b[REC_MAX_HEADS], *bm[REC_MAX_HEADS], *bv[REC_MAX_HEADS];
};

__device__ __forceinline__ void adam1(float &p, float &m, float &v, float g, float b1, float b2, float eps,
                                      float step_size, float bc2_sqrt) {
  adam_elem(p, m, v, g, b1, b2, eps, step_size, 1.f / bc2_sqrt);
}

__global__ void __launch_bounds__(256) head_bwd_adam_kernel(HeadTrainPtrs hp, const float *__restrict__ h,
                                                            const int64_t *__restrict__ target,
                                                            const float *__restrict__ row_stats,
                                                            const float *__restrict__ dq, int B, int D, int Vloc,
                                                            int vocab_lo, int n_tiles, float inv_B,
                                                            float *__restrict__ dh_part, float b1, float b2,
                                                            float eps, float step_size, float bc2_sqrt,
                                                            int head_begin, const float *__restrict__ sc,
                                                            const float *__restrict__ extra) {
  if (sc) { step_size = sc[0]; bc2_sqrt = 1.f / sc[1]; }
  extern __shared__ __align__(16) float dyn[];
  const int DP = D + 4;
  float *dW = dyn;                                   // [TN][DP]
  float *db = dW + TN * DP;                          // [TN]
  float(*As)[KP] = reinterpret_cast<float(*)[KP]>(db + TN);     // [TM][KP]
  float(*Bs)[KP] = As + TM;                                      // [TN][KP]
  float(*DL)[TN + 4] = reinterpret_cast<float(*)[TN + 4]>(Bs + TN);   // [TM][TN+4]  dl[m][n]
  float(*DLT)[TM + 4] = reinterpret_cast<float(*)[TM + 4]>(DL + TM);  // [TN][TM+4]  dl[n][m]
  float(*Hc)[TN + 4] = reinterpret_cast<float(*)[TN + 4]>(DLT + TN);  // [64][68] h chunk / W chunk
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int head = blockIdx.y + head_begin;
  float *__restrict__ W = hp.w[head];
  float *dh_slice = dh_part + (int64_t)blockIdx.x * B * D;
  bool first_tile = true;

  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int v0 = t * TN;
    for (int i = tid; i < TN * DP; i += 256) dW[i] = 0.f;
    if (tid < TN) db[tid] = 0.f;
    __syncthreads();
    if (head == 0) {
      for (int b0 = 0; b0 < B; b0 += TM) {
        float acc[4][4];
        logits_tile(As, Bs, h, b0, B, W, v0, Vloc, D, tid, ty, tx, acc);
        // dlogits tile
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          int m = ty + 16 * a, row = b0 + m;
          float lse = row < B ? row_stats[(int64_t)row * ROW_STRIDE] : 0.f;
          int64_t trow = row < B ? target[row] - vocab_lo : -1;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            int n = tx + 16 * c, col = v0 + n;
            float dl = 0.f;
            if (row < B && col < Vloc) {
              float l = acc[a][c] + __ldg(hp.b[0] + col);
              dl = (__expf(l - lse) - ((int64_t)col == trow ? 1.f : 0.f)) * inv_B;
              if (extra && (int64_t)col == trow) dl += extra[row];
            }
            DL[m][n] = dl;
            DLT[n][m] = dl;
          }
        }
        __syncthreads();
        if (tid < TN) {
          float sacc = 0.f;
          for (int m = 0; m < TM; ++m) sacc += DL[m][tid];
          db[tid] += sacc;
        }
        for (int d0 = 0; d0 < D; d0 += 64) {
          // ---- dW[n][d0+kd] += sum_m dl[m][n] * h[b0+m][d0+kd] ----
          for (int e = tid; e < 64 * 16; e += 256) {
            int r = e >> 4, c4 = e & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b0 + r < B && d0 + c4 * 4 < D) v = __ldg(reinterpret_cast<const float4 *>(h + (int64_t)(b0 + r) * D + d0) + c4);
            *reinterpret_cast<float4 *>(&Hc[r][c4 * 4]) = v;
          }
          __syncthreads();
          {
            float g[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) g[i][j] = 0.f;
#pragma unroll 8
            for (int m = 0; m < TM; ++m) {
              float4 a4 = *reinterpret_cast<const float4 *>(&DL[m][ty * 4]);
              float4 b4 = *reinterpret_cast<const float4 *>(&Hc[m][tx * 4]);
              float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) g[i][j] = fmaf(av[i], bv[j], g[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                int kd = d0 + tx * 4 + j;
                if (kd < D) dW[(ty * 4 + i) * DP + kd] += g[i][j];
              }
          }
          __syncthreads();
          // ---- dh[b0+m][d0+kd] (+)= sum_n dl[m][n] * W[v0+n][d0+kd] ----
          for (int e = tid; e < 64 * 16; e += 256) {
            int r = e >> 4, c4 = e & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v0 + r < Vloc && d0 + c4 * 4 < D) v = *(reinterpret_cast<const float4 *>(W + (int64_t)(v0 + r) * D + d0) + c4);
            *reinterpret_cast<float4 *>(&Hc[r][c4 * 4]) = v;
          }
          __syncthreads();
          {
            float g[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) g[i][j] = 0.f;
#pragma unroll 8
            for (int n = 0; n < TN; ++n) {
              float4 a4 = *reinterpret_cast<const float4 *>(&DLT[n][ty * 4]);
              float4 b4 = *reinterpret_cast<const float4 *>(&Hc[n][tx * 4]);
              float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) g[i][j] = fmaf(av[i], bv[j], g[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              int row = b0 + ty * 4 + i;
              if (row >= B) continue;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                int kd = d0 + tx * 4 + j;
                if (kd < D) {
                  float *dst = dh_slice + (int64_t)row * D + kd;
                  *dst = first_tile ? g[i][j] : (*dst + g[i][j]);
                }
              }
            }
          }
          __syncthreads();
        }
      }
    } else {
      // sparse rows: deterministic order over b
      for (int b = 0; b < B; ++b) {
        int64_t loc = target[b] - vocab_lo - v0;
        if (loc >= 0 && loc < TN && v0 + loc < Vloc) {
          float g = dq[b * 3 + head - 1];
          for (int k = tid; k < D; k += 256) dW[(int)loc * DP + k] += g * h[(int64_t)b * D + k];
          if (tid == 0) db[loc] += g;
        }
      }
      __syncthreads();
    }
    first_tile = false;
    // ---- Adam on the tile (coalesced over d) ----
    const int D4 = D >> 2;
    for (int e = tid; e < TN * D4; e += 256) {
      int n = e / D4, c4 = e - n * D4;
      if (v0 + n >= Vloc) continue;
      int64_t off = ((int64_t)(v0 + n) * D >> 2) + c4;
      float4 p = reinterpret_cast<float4 *>(W)[off];
      float4 m = reinterpret_cast<float4 *>(hp.wm[head])[off];
      float4 v = reinterpret_cast<float4 *>(hp.wv[head])[off];
      const float *g = dW + n * DP + c4 * 4;
      adam1(p.x, m.x, v.x, g[0], b1, b2, eps, step_size, bc2_sqrt);
      adam1(p.y, m.y, v.y, g[1], b1, b2, eps, step_size, bc2_sqrt);
      adam1(p.z, m.z, v.z, g[2], b1, b2, eps, step_size, bc2_sqrt);
      adam1(p.w, m.w, v.w, g[3], b1, b2, eps, step_size, bc2_sqrt);
      reinterpret_cast<float4 *>(W)[off] = p;
      reinterpret_cast<float4 *>(hp.wm[head])[off] = m;
      reinterpret_cast<float4 *>(hp.wv[head])[off] = v;
    }
    if (tid < TN && v0 + tid < Vloc) {
      float p = hp.b[head][v0 + tid], m = hp.bm[head][v0 + tid], v = hp.bv[head][v0 + tid];
      adam1(p, m, v, db[tid], b1, b2, eps, step_size, bc2_sqrt);
      hp.b[head][v0 + tid] = p; hp.bm[head][v0 + tid] = m; hp.bv[head][v0 + tid] = v;
    }
    __syncthreads();
  }
  // a CTA of the dense head that owned no tile must still define its dh slice
  if (head == 0 && first_tile)
    for (int64_t i = tid; i < (int64_t)B * D; i += 256) dh_slice[i] = 0.f;
}

// dh[i] = sum over slices in a FIXED order: 4 slice groups per output run in parallel (group g takes slices
// g, g+4, ...), then the 4 partial sums are added in group order.  block = 256 threads = 64 outputs x 4 groups.
__global__ void __launch_bounds__(256) dh_reduce_kernel(const float *__restrict__ dh_part, int n_slices, int64_t n,
                                                        float *__restrict__ dh) {
  __shared__ float sh[4][64];
  const int o = threadIdx.x & 63, g = threadIdx.x >> 6;
  const int64_t i = (int64_t)blockIdx.x * 64 + o;
  float acc = 0.f;
  if (i < n)
    for (int s = g; s < n_slices; s += 4) acc += dh_part[(int64_t)s * n + i];
  sh[g][o] = acc;
  __syncthreads();
  if (g == 0 && i < n) dh[i] = ((sh[0][o] + sh[1][o]) + sh[2][o]) + sh[3][o];
}

// ------------------------------------------------------------------------------------------------
static HeadPtrs head_ptrs(const rec_engine *e, int net_id) {
  HeadPtrs hp;
  for (int i = 0; i < REC_MAX_HEADS; ++i) { hp.w[i] = e->nets[net_id].p.head_w[i]; hp.b[i] = e->nets[net_id].p.head_b[i]; }
  return hp;
}

int launch_head_stats(rec_engine *e, const HeadStatsArgs &a, int *n_split_out) {
  const int n_tiles = cdiv(e->Vloc, TN), nb = cdiv(a.B, TM);
  int n_split = cdiv(2 * e->sm_count, nb);
  if (n_split > n_tiles) n_split = n_tiles;
  if (n_split > e->n_split_max) n_split = e->n_split_max;
  if (n_split < 1) n_split = 1;
  // make every split non-empty
  int per = cdiv(n_tiles, n_split);
  n_split = cdiv(n_tiles, per);
  dim3 grid(n_split, nb);
  HeadPtrs hps = head_ptrs(e, a.net_id);
  if (a.n_arg > 0 && a.arg_shift != 0) {  // the kernel scores heads 1 + j: present the requested heads there
    const HeadPtrs all = hps;
    for (int j = 0; j < a.n_arg; ++j) { hps.w[1 + j] = all.w[1 + a.arg_shift + j]; hps.b[1 + j] = all.b[1 + a.arg_shift + j]; }
  }
  head_stats_kernel<<<grid, 256, 0, e->stream>>>(hps, a.h, a.B, e->D, e->Vloc, e->cfg.vocab_lo, n_tiles,
                                                a.do_stats, a.stats_head, a.target, a.topk, a.n_arg, a.w[0], a.w[1],
                                                a.w[2], e->part, e->part_stride);
  REC_LAUNCH_CHECK(e);
  *n_split_out = n_split;
  e->st_approx = false;  // fp32 FFMA scores: ranked as they are
  e->st_kpub = a.topk;
  e->st_apub = 0;
  return REC_OK;
}

static RescoreSrc rescore_src(const rec_engine *e, const HeadStatsArgs *src, int topk) {
  RescoreSrc R = {};
  R.kpub = topk;
  if (!src) return R;  // exact per-shard summaries (cross-GPU merge)
  const rec_net_params &p = e->nets[src->net_id].p;
  R.h = src->h;
  R.n_arg = src->n_arg;
  if (src->n_arg > 0) {
    for (int j = 0; j < src->n_arg && j < 3; ++j) { R.w[j] = p.head_w[1 + src->arg_shift + j]; R.b[j] = p.head_b[1 + src->arg_shift + j]; R.wq[j] = src->w[j]; }
  } else {
    R.w[0] = p.head_w[src->stats_head]; R.b[0] = p.head_b[src->stats_head];
  }
  R.approx = e->st_approx ? 1 : 0;
  R.kpub = topk > 0 ? e->st_kpub : 0;
  R.apub = e->st_apub;
  return R;
}

int launch_head_merge(rec_engine *e, const float *part, int n_split, int B, int topk, bool has_stats, bool has_arg,
                      float *summary, const HeadStatsArgs *src) {
  const RescoreSrc R = rescore_src(e, src, topk);
  if (R.approx && has_arg && topk > 0) REC_FAIL(e, REC_EINVAL, "head merge: a record holds either top-k or greedy-action candidates");
  if (n_split <= 160 && topk <= 2 && (e->part_stride & 3) == 0 && R.kpub <= 4 && R.apub <= 4)
    head_merge_small_kernel<<<cdiv(B, 4), 128, 0, e->stream>>>(part, e->part_stride, n_split, B, topk, has_stats ? 1 : 0,
                                                              has_arg ? 1 : 0, e->row_stats, e->row_ids, e->row_topv,
                                                              e->astar, summary, R, e->D, e->cfg.vocab_lo, e->Vloc);
  else
    head_merge_kernel<<<cdiv(B, 8), 256, 0, e->stream>>>(part, e->part_stride, n_split, B, topk, has_stats ? 1 : 0,
                                                        has_arg ? 1 : 0, e->row_stats, e->row_ids, e->row_topv, e->astar,
                                                        summary, R, e->D, e->cfg.vocab_lo, e->Vloc);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

int launch_head_logits(rec_engine *e, int net_id, int head, const float *h, int B, float *logits, int64_t ld) {
  dim3 grid(cdiv(e->Vloc, TN), cdiv(B, TM));
  const rec_net_params &p = e->nets[net_id].p;
  head_logits_kernel<<<grid, 256, 0, e->stream>>>(p.head_w[head], p.head_b[head], h, B, e->D, e->Vloc, logits, ld);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

int launch_row_dots(rec_engine *e, int net_id, const float *h, const int64_t *ids, const int32_t *ids32, int B,
                    int first_head, int n, float *out) {  // out is [B, 3]
  row_dots_kernel<<<cdiv(B * n, 8), 256, 0, e->stream>>>(head_ptrs(e, net_id), h, ids, ids32, B, e->D, e->Vloc,
                                                        e->cfg.vocab_lo, first_head, n, out, 3);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

size_t head_bwd_smem_bytes(int D) {
  return sizeof(float) * ((size_t)TN * (D + 4) + TN + 2 * TM * KP + TM * (TN + 4) + TN * (TM + 4) + 64 * (TN + 4));
}

// Number of per-CTA dh slices the supervised-head kernel writes; the Q heads' slice follows them.
int head_bwd_dense_slices(const rec_engine *e, int B) {
  if (tc_bwd_supported(e, B)) return tc_bwd_slices(e);
  if (tck_heads_supported(e)) return tck_bwd_slices(e);
  const int n_tiles = cdiv(e->Vloc, TN);
  int n_cta = e->n_dh_part - 1;
  return n_cta > n_tiles ? n_tiles : n_cta;
}

// Fused per-row Q path of the single-GPU step (see q_rows_fused_kernel).
int launch_q_rows_fused(rec_engine *e, int main_net, const rec_batch *b, const rec_train_hparams *hp, int n_split,
                        float alpha_eff, float *q_loss_rows, const HeadStatsArgs *src) {
  const int B = b->B;
  QRowArgs A;
  A.R = rescore_src(e, src, 0);
  A.part = e->part; A.part_stride = e->part_stride; A.n_split = n_split;
  A.main_heads = head_ptrs(e, main_net); A.boot_heads = head_ptrs(e, 1 - main_net);
  A.h_main = e->h_state[0]; A.h_boot = e->h_state[2];
  A.a = b->a; A.s = b->s; A.div_lens = b->true_next_len;  // (q2) the reference indexes s with true_next_len
  A.r_acc = b->r; A.is_end = b->is_end; A.row_ids = e->row_ids;
  A.B = B; A.D = e->D; A.L = e->cfg.state_size; A.N = e->cfg.item_num; A.V = e->cfg.action_dim; A.Vloc = e->Vloc; A.vocab_lo = e->cfg.vocab_lo;
  A.n_q = e->cfg.n_heads - 1; A.alpha_eff = alpha_eff;
  A.row_stats = e->row_stats; A.astar = e->astar;
  A.q_sa = e->q_sa; A.q_boot = e->q_boot; A.dq = e->dq; A.q_loss_rows = q_loss_rows; A.rewards = e->rewards;
  A.dh_slice = e->dh_part + (int64_t)head_bwd_dense_slices(e, B) * B * e->D;
  if (n_split <= 160 && e->D == 64) q_rows_fused_small_kernel<2><<<cdiv(B, 4), 128, 0, e->stream>>>(A, *hp);
  else if (n_split <= 160 && e->D == 128) q_rows_fused_small_kernel<4><<<cdiv(B, 4), 128, 0, e->stream>>>(A, *hp);
  else q_rows_fused_kernel<<<cdiv(B, 8), 256, 0, e->stream>>>(A, *hp);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

// Q heads: dh contribution of rows a_b; must run before Adam touches the Q-head weights.
int launch_q_dh(rec_engine *e, int net_id, const rec_batch *b, int B) {
  const int n_q = e->cfg.n_heads - 1;
  if (n_q <= 0) return REC_OK;
  float *q_slice = e->dh_part + (int64_t)head_bwd_dense_slices(e, B) * B * e->D;
  q_dh_kernel<<<B, 128, 0, e->stream>>>(head_ptrs(e, net_id), b->a, e->dq, B, e->D, e->Vloc, e->cfg.vocab_lo, n_q, q_slice);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

// Supervised head: backward + Adam fused (tensor-core kernel when the shape allows it).
int launch_sup_head_bwd(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B, float step_size,
                        float bc2_sqrt, const rec_train_hparams *hp, float inv_B) {
  if (e->timing) cudaEventRecord(e->ev[0], e->stream);
  if (tc_bwd_supported(e, B)) {
    int n_slices = 0;
    int rc = launch_head_bwd_adam_tc(e, net_id, h, b, B, step_size, bc2_sqrt, hp, inv_B, &n_slices);
    if (rc) return rc;
  } else if (tck_heads_supported(e)) {
    int rc = launch_head_bwd_adam_tck(e, net_id, h, b, B, step_size, bc2_sqrt, hp, inv_B);
    if (rc) return rc;
  } else {
    const rec_net_params &p = e->nets[net_id].p;
    HeadTrainPtrs t;
    for (int i = 0; i < REC_MAX_HEADS; ++i) {
      t.w[i] = p.head_w[i]; t.wm[i] = p.head_w_m[i]; t.wv[i] = p.head_w_v[i];
      t.b[i] = p.head_b[i]; t.bm[i] = p.head_b_m[i]; t.bv[i] = p.head_b_v[i];
    }
    const int n_tiles = cdiv(e->Vloc, TN);
    size_t smem = head_bwd_smem_bytes(e->D);
    if (smem > 220 * 1024) REC_FAIL(e, REC_EINVAL, "head backward needs %zu B of shared memory (D=%d too large)", smem, e->D);
    static bool attr_set[REC_MAX_DEVICES] = {};  // per device: the opt-in is a per-device function attribute
    if (!attr_set[e->dev]) {
      REC_CUDA(e, cudaFuncSetAttribute(head_bwd_adam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      attr_set[e->dev] = true;
    }
    dim3 grid(head_bwd_dense_slices(e, B), 1);
    head_bwd_adam_kernel<<<grid, 256, smem, e->stream>>>(t, h, b->a, e->row_stats, e->dq, B, e->D, e->Vloc, e->cfg.vocab_lo,
                                                        n_tiles, inv_B, e->dh_part, hp->beta1, hp->beta2, hp->eps,
                                                        step_size, bc2_sqrt, 0, e->d_sc, e->bwd_extra);
    REC_LAUNCH_CHECK(e);
  }
  if (e->timing) cudaEventRecord(e->ev[1], e->stream);
  return REC_OK;
}

// Row-sparse gradients of the Q heads + dense Adam: pure HBM streaming (24 B/param).
int launch_q_heads_update(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B, float step_size,
                          float bc2_sqrt, const rec_train_hparams *hp, int wait_mark) {
  if (e->cfg.n_heads < 2) return REC_OK;
  // kernel-timing mode: the events of slot 3 bracket the adam_stream_kernel launch alone (embed.cu)
  return launch_q_heads_adam(e, net_id, h, b, B, step_size, bc2_sqrt, hp, wait_mark);
}

int launch_dh_reduce(rec_engine *e, int B) {
  const int64_t n = (int64_t)B * e->D;
  const int slices = head_bwd_dense_slices(e, B) + (e->cfg.n_heads > 1 ? 1 : 0);
  dh_reduce_kernel<<<(int)cdiv64(n, 64), 256, 0, e->stream>>>(e->dh_part, slices, n, e->dh);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

// The whole head backward in program order, with the Q-head sweep on a side stream next to the (latency-bound)
// supervised-head kernel.  The fused single-GPU Q step orders the pieces itself (api.cu).
int launch_head_backward_adam(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B, float step_size,
                              float bc2_sqrt, const rec_train_hparams *hp, float inv_B) {
  int rc;
  if ((rc = launch_q_dh(e, net_id, b, B))) return rc;
  {
    SideScope side(e, 0);
    if ((rc = launch_q_heads_update(e, net_id, h, b, B, step_size, bc2_sqrt, hp))) return rc;
  }
  if ((rc = launch_sup_head_bwd(e, net_id, h, b, B, step_size, bc2_sqrt, hp, inv_B))) return rc;
  return launch_dh_reduce(e, B);
}
