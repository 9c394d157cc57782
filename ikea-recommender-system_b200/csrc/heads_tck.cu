// Full-vocabulary heads for wide states (D = 128, 256, 384, 512, ...: BidirGRU4Rec-SQN, BASELINE cfg3 has D = 512) on
// the 5th-generation tensor cores.  With D > 64 neither the states (h[256 x 512] as bf16 hi/lo = 512 KB) nor a weight
// tile fit in shared memory next to each other, so every GEMM here is a K-LOOP pipeline over PACKED OPERAND IMAGES:
//
//   image of a matrix X[R, C] (C % 64 == 0): blocks [ceil(R/128)][C/64], block = bf16 hi [128 rows][64] (16 KB) | bf16
//   lo (16 KB), 128-byte swizzle -- exactly the bytes tcgen05.mma wants in shared memory, so a pipeline stage is filled
//   by plain TMA bulk copies.  The same bytes serve as a K-major operand (MN = rows, K = the 64 columns) or as an
//   MN-major operand (MN = the 64 columns, K = rows): logits, dW and dh all read the SAME images.
//
//   pack_img_kernel          fp32 master weights -> image (once per step and head; the Q heads of the greedy-action
//                            pass are pre-combined  sum_j w_j W_j  while they are packed), h -> image, h^T -> image
//   tck_kernel<HeadFwd<M>>   logits tile [128 sessions x 128 items] = h . W^T, K-loop over D, accumulators in TMEM
//                            (double-buffered over vocabulary tiles); epilogue M = STATS: online (max, sum-exp) +
//                            target logit | ARG: running argmax (greedy action) | DL: dlogits = (softmax - onehot)/B
//                            written as the image dl^T[V, B] (+ bias-gradient partials)
//   tck_kernel<HeadDh>       dh[128 sessions x 256] += dl . W   (both operands MN-major views), split-K over the catalogue
//   tck_kernel<HeadDwAdam>   dW tile [128 items x 256] = dl^T . h (K = batch) -> fused Adam on W, m, v (+ bias)
//
// One kernel skeleton (tck_kernel): 8 epilogue warps + 1 MMA-issuer warp + 1 TMA-loader warp that meet only at
// mbarriers (stage full / stage empty / accumulator full / accumulator empty).  The products are evaluated as
// hi*hi + hi*lo + lo*hi with fp32 accumulation ("bf16x3", ~1e-5 relative) exactly like the D = 64 kernels.
// Reference semantics: nn.Linear heads + CrossEntropyLoss + Adam of models/SQN/sqn_gru.py:78-112,183-254 and
// models/BidirGRU4Rec/model.py:51-99 (cfg3 = SQN heads on the concatenated bidirectional state).
#include <stdio.h>
#include <stdlib.h>
#include "tck.cuh"

namespace tck {

// ------------------------------------------------------------------------------------------------------------
// Forward: logits[128 sessions x 128 items] per unit, K-loop over D
// ------------------------------------------------------------------------------------------------------------
struct FwdParams {
  const uint8_t *himg, *wimg;  // [n_sb][KB], [n_tiles][KB]
  int KB, B, Vloc, vocab_lo, n_tiles;
  const float *bias;           // [Vloc] (combined for the greedy-action pass)
  const int64_t *target;       // [B] or null
  float *part;                 // STATS / ARG: records [n_split][B][part_stride]
  int part_stride;
  const float *row_stats;      // DL: merged statistics (lse at [row * 8])
  float inv_B;
  uint8_t *dlT;                // DL: image [n_tiles][dl_cb]
  int dl_cb;
  float *db_part;              // DL: [n_sb][n_tiles * 128]
  int topk;                    // HeadTopk: k of the running top-k (<= 20)
  float *cmax;                 // HeadTopk<.., CM>: chunk maxima [B][cmax_ld >= n_tiles * 4]
  int64_t cmax_ld;
  long long *trace;            // HeadCmaxPair: clock64 stamps [8 kinds][64 units] of one CTA (debugging; normally null)
  const uint8_t *bblk;         // HeadCmaxPair: bias operand blocks [n_tiles + 1][4 KB] (block n_tiles = the "ones" operand)
  float *cmax2;                // HeadCmaxPair: level-2 maxima [B][cmax2_ld >= 2 * ceil(n_tiles / 8)]
  int64_t cmax2_ld;
  int n_sb, per;               // HeadCmaxFlat: session blocks, units per CTA of the flattened (session block, tile) space
  const float *extra;          // DL: optional per-row gradient added at the target column (SARM)
};

enum { M_STATS = 0, M_ARG = 1, M_DL = 2 };

template <int MODE>
struct HeadFwd {
  using Params = FwdParams;
  static constexpr bool CLUSTERED = false;
  static constexpr int RESIDENT_BYTES = 0;
  static constexpr int EPI_WARPS = 8;
  static constexpr const char *NAME = MODE == M_STATS ? "tck:head_stats" : MODE == M_ARG ? "tck:head_greedy" : "tck:head_dlogits";
  static constexpr int STAGES = 3, STAGE_BYTES = 2 * BLK2, ACC_COLS = 128, TMEM_COLS = 256;
  static constexpr int EXTRA_BYTES = 8192;

  __device__ static __forceinline__ void units(const Params &p, int &lo, int &hi) {
    const int per = (p.n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    lo = blockIdx.x * per;
    hi = min(p.n_tiles, lo + per);
    if (hi < lo) hi = lo;
  }
  __device__ static __forceinline__ int k_steps(const Params &p, int) { return p.KB; }
  __device__ static __forceinline__ void load(const Params &p, int u, int ks, uint8_t *stage, uint64_t *bar) {
    tc::mbar_expect_tx(bar, 2 * BLK2);
    tc::bulk_g2s(stage, p.himg + ((int64_t)blockIdx.y * p.KB + ks) * BLK2, BLK2, bar);
    tc::bulk_g2s(stage + BLK2, p.wimg + ((int64_t)u * p.KB + ks) * BLK2, BLK2, bar);
  }
  __device__ static __forceinline__ void mma(const Params &, int, int, uint32_t st, uint32_t tacc, bool first) {
    const uint32_t id = tc::instr_desc(128, 128, 0, 0);
    const uint64_t ah = tc::desc_kmajor(st, 0), al = tc::desc_kmajor(st + BLK, 0);
    const uint64_t bh = tc::desc_kmajor(st + BLK2, 0), bl = tc::desc_kmajor(st + BLK2 + BLK, 0);
    bool acc = !first;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const uint64_t a = pass == 2 ? al : ah, b = pass == 1 ? bl : bh;
#pragma unroll
      for (int k = 0; k < 4; ++k) { tc::mma_bf16(tacc, a + (uint64_t)(k * 2), b + (uint64_t)(k * 2), id, acc); acc = true; }
    }
  }

  struct Epi {
    float *xs;  // extra shared memory
    int q, cq, lane, row, trow;
    float m_run, s_run, tgt, av, cst;
    int ai;
    bool rv;
    __device__ __forceinline__ Epi(const Params &p, uint8_t *extra, int tid) {
      xs = reinterpret_cast<float *>(extra);
      const int warp = tid >> 5;
      lane = tid & 31; q = warp & 3; cq = warp >> 2;
      row = blockIdx.y * 128 + q * 32 + lane;
      rv = row < p.B;
      m_run = REC_NEG_INF; s_run = 0.f; tgt = REC_NEG_INF; av = REC_NEG_INF; ai = 0x7fffffff;
      trow = (MODE != M_ARG && p.target && rv) ? (int)(p.target[row] - p.vocab_lo) : -1;
      cst = 0.f;
      if (MODE == M_DL) cst = fmaf(-(rv ? p.row_stats[(int64_t)row * ROW_STRIDE_] : 0.f), LOG2E, __log2f(p.inv_B));
    }
    __device__ __forceinline__ void tile(const Params &p, int u, int i, uint32_t tacc) {
      const int v0 = u * 128;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int c_lo = v0 + cq * 64 + half * 32;
        float l[32];
        tc::tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 64 + half * 32), l);
        if (c_lo + 32 <= p.Vloc) {
          const float4 *bg = reinterpret_cast<const float4 *>(p.bias + c_lo);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(bg + (j >> 2));
            l[j] += b4.x; l[j + 1] += b4.y; l[j + 2] += b4.z; l[j + 3] += b4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) l[j] = (c_lo + j < p.Vloc) ? l[j] + __ldg(p.bias + c_lo + j) : REC_NEG_INF;
        }
        if (MODE == M_STATS) {
          float tmax = fmaxf(l[0], l[1]);
#pragma unroll
          for (int j = 2; j < 32; ++j) tmax = fmaxf(tmax, l[j]);
          const float nm = fmaxf(m_run, tmax), nml = -fmaxf(nm, -1e30f) * LOG2E;
          float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 32; ++j) ps[j & 3] += tc::ex2_ftz(fmaf(l[j], LOG2E, nml));
          s_run = fmaf(s_run, tc::ex2_ftz((m_run - nm) * LOG2E), (ps[0] + ps[1]) + (ps[2] + ps[3]));
          m_run = nm;
          if (trow >= c_lo && trow < c_lo + 32) {
            const int tj = trow - c_lo;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j == tj) tgt = l[j];
          }
        } else if (MODE == M_ARG) {
          float cm[4] = {l[0], l[1], l[2], l[3]};
#pragma unroll
          for (int j = 4; j < 32; ++j) cm[j & 3] = fmaxf(cm[j & 3], l[j]);
          if (fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])) > av) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (l[j] > av) { av = l[j]; ai = p.vocab_lo + c_lo + j; }  // ascending ids: strict > keeps the lowest id
          }
        } else {
          // dlogits = exp(l - lse) / B  (- 1/B at the target); zero outside the matrix
          const int tj = trow - c_lo;
          const int nvalid = rv ? min(32, p.Vloc - c_lo) : 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) l[j] = tc::ex2_ftz(fmaf(l[j], LOG2E, cst));
          if (tj >= 0 && tj < 32) {
            const float sub = p.inv_B - ((p.extra && rv) ? __ldg(p.extra + row) : 0.f);
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j == tj) l[j] -= sub;
          }
          if (nvalid < 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j >= nvalid) l[j] = 0.f;
          }
          // dl^T image: row = item (tile u, row r), column = session
          const int sl = q * 32 + lane;
          uint8_t *blk = p.dlT + ((int64_t)u * p.dl_cb + 2 * blockIdx.y + (sl >> 6)) * BLK2;
          const int c = sl & 63;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            uint32_t hi, lo;
            tc::split_bf16x2(l[j], l[j + 1], hi, lo);
            const int r = cq * 64 + half * 32 + j;
            const uint32_t o0 = tc::sw128_off(r, c), o1 = tc::sw128_off(r + 1, c);
            *reinterpret_cast<uint16_t *>(blk + o0) = (uint16_t)(hi & 0xFFFFu);
            *reinterpret_cast<uint16_t *>(blk + o1) = (uint16_t)(hi >> 16);
            *reinterpret_cast<uint16_t *>(blk + BLK + o0) = (uint16_t)(lo & 0xFFFFu);
            *reinterpret_cast<uint16_t *>(blk + BLK + o1) = (uint16_t)(lo >> 16);
          }
          // bias gradient: column sums over the warp's 32 sessions (transpose-reduce: lane j ends with column j)
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
#pragma unroll
            for (int k = 0; k < o; ++k) {
              const bool up = (lane & o) != 0;
              const float send = up ? l[k] : l[k + o];
              const float keep = up ? l[k + o] : l[k];
              l[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          xs[(i & 1) * 512 + q * 128 + cq * 64 + half * 32 + lane] = l[0];
        }
      }
      if (MODE == M_DL) {
        epi_bar();
        const int t = threadIdx.x;
        if (t < 128) {
          const float *d = xs + (i & 1) * 512 + t;
          p.db_part[((int64_t)blockIdx.y * p.n_tiles + u) * 128 + t] = (d[0] + d[128]) + (d[256] + d[384]);
        }
      }
    }
    __device__ __forceinline__ void finish(const Params &p) {
      if (MODE == M_DL) return;
      // combine the two column halves of every row, publish ONE record per (split, row)
      float *x = xs + ((q * 32 + lane) * 2 + cq) * 5;
      x[0] = m_run; x[1] = s_run; x[2] = tgt; x[3] = av; x[4] = __int_as_float(ai);
      epi_bar();
      if (cq == 0 && rv) {
        const float *y = xs + ((q * 32 + lane) * 2) * 5;
        const float m = fmaxf(y[0], y[5]), tg = fmaxf(y[2], y[7]);
        float ssum = 0.f;
        if (y[1] > 0.f) ssum += y[1] * __expf(y[0] - m);
        if (y[6] > 0.f) ssum += y[6] * __expf(y[5] - m);
        float v0 = y[3], v1 = y[8];
        int i0 = __float_as_int(y[4]), i1 = __float_as_int(y[9]);
        if (better(v1, i1, v0, i0)) { float tv = v0; v0 = v1; v1 = tv; int ti = i0; i0 = i1; i1 = ti; }
        float *o = p.part + ((int64_t)blockIdx.x * p.B + row) * p.part_stride;
        o[0] = m; o[1] = ssum; o[2] = tg; o[3] = v0; o[4] = __int_as_float(i0);
        if (MODE == M_ARG) {  // two approximate candidates per record: the merge re-scores the best few in fp32
          o[TOPK_OFF] = v0; o[TOPK_OFF + REC_MAX_TOPK] = __int_as_float(i0);
          o[TOPK_OFF + 1] = v1; o[TOPK_OFF + REC_MAX_TOPK + 1] = __int_as_float(i1);
        }
      }
    }
  };
};

// ------------------------------------------------------------------------------------------------------------
// Evaluation-shaped forward: statistics + running top-k (k <= 20; score desc, id asc) of every row, 16 epilogue warps
// (four 32-column quarters per row and tile).  ARES: D = 64 -- the 128-session state block stays resident in shared
// memory and only the weight image streams (3 stages); D > 64: state and weight k-blocks stream together (2 stages).
// Against head_stats_tc_kernel (heads_tc.cu) at evaluation shapes: no fp32 staging ring and no converter warps (the
// image is packed once per batch: 85 us at 1 M items against ~4 ms of scoring), twice the epilogue warps per logit.
// ------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void topk_insert(float v, int id, float *lv, int *li, int stride, int topk, int &cnt, float &tau) {
  int p = cnt < topk ? cnt : topk - 1;
  while (p > 0 && lv[(p - 1) * stride] < v) {
    lv[p * stride] = lv[(p - 1) * stride];
    li[p * stride] = li[(p - 1) * stride];
    --p;
  }
  lv[p * stride] = v;
  li[p * stride] = id;
  if (cnt < topk) ++cnt;
  tau = cnt == topk ? lv[(topk - 1) * stride] : REC_NEG_INF;
}

// (Large evaluation batches over large catalogues use HeadCmaxFlat + chunk_select / chunk_score below instead: no lists,
// no divergence in the tensor-core pass.)
template <bool ARES>
struct HeadTopk {
  static constexpr bool CM = false;
  using Params = FwdParams;
  static constexpr bool CLUSTERED = false;
  static constexpr int EPI_WARPS = 16, NT = 512, CS = 4, KMAX = 20;
  static constexpr int RESIDENT_BYTES = ARES ? BLK2 : 0;
  static constexpr const char *NAME = CM ? (ARES ? "tck:head_cmax64" : "tck:head_cmax") : (ARES ? "tck:head_topk64" : "tck:head_topk");
  static constexpr int STAGES = ARES ? (CM ? 4 : 3) : (CM ? 3 : 2), STAGE_BYTES = ARES ? BLK2 : 2 * BLK2, ACC_COLS = 128, TMEM_COLS = 256;
  static constexpr int LIST_BYTES = CM ? 0 : KMAX * NT * 8, XCH_BYTES = 128 * CS * 5 * 4;
  static constexpr int EXTRA_BYTES = LIST_BYTES + XCH_BYTES;

  __device__ static __forceinline__ void units(const Params &p, int &lo, int &hi) {
    const int per = (p.n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    lo = blockIdx.x * per;
    hi = min(p.n_tiles, lo + per);
    if (hi < lo) hi = lo;
  }
  __device__ static __forceinline__ int k_steps(const Params &p, int) { return p.KB; }
  __device__ static __forceinline__ void load_resident(const Params &p, uint8_t *res, uint64_t *bar) {
    tc::mbar_expect_tx(bar, BLK2);
    tc::bulk_g2s(res, p.himg + (int64_t)blockIdx.y * p.KB * BLK2, BLK2, bar);
  }
  __device__ static __forceinline__ void load(const Params &p, int u, int ks, uint8_t *stage, uint64_t *bar) {
    if (ARES) {
      tc::mbar_expect_tx(bar, BLK2);
      tc::bulk_g2s(stage, p.wimg + ((int64_t)u * p.KB + ks) * BLK2, BLK2, bar);
    } else {
      tc::mbar_expect_tx(bar, 2 * BLK2);
      tc::bulk_g2s(stage, p.himg + ((int64_t)blockIdx.y * p.KB + ks) * BLK2, BLK2, bar);
      tc::bulk_g2s(stage + BLK2, p.wimg + ((int64_t)u * p.KB + ks) * BLK2, BLK2, bar);
    }
  }
  __device__ static __forceinline__ void mma_ab(uint32_t a0, uint32_t b0, uint32_t tacc, bool first) {
    const uint32_t id = tc::instr_desc(128, 128, 0, 0);
    const uint64_t ah = tc::desc_kmajor(a0, 0), al = tc::desc_kmajor(a0 + BLK, 0);
    const uint64_t bh = tc::desc_kmajor(b0, 0), bl = tc::desc_kmajor(b0 + BLK, 0);
    bool acc = !first;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const uint64_t a = pass == 2 ? al : ah, b = pass == 1 ? bl : bh;
#pragma unroll
      for (int k = 0; k < 4; ++k) { tc::mma_bf16(tacc, a + (uint64_t)(k * 2), b + (uint64_t)(k * 2), id, acc); acc = true; }
    }
  }
  __device__ static __forceinline__ void mma(const Params &, int, int, uint32_t st, uint32_t res, uint32_t tacc, bool first) {
    mma_ab(res, st, tacc, first);
  }
  __device__ static __forceinline__ void mma(const Params &, int, int, uint32_t st, uint32_t tacc, bool first) {
    mma_ab(st, st + BLK2, tacc, first);
  }

  struct Epi {
    float *lv, *xs;
    int *li;
    int q, cq, lane, row, trow, cnt, tid_;
    float m_run, s_run, tgt, tau;
    bool rv;
    __device__ __forceinline__ Epi(const Params &p, uint8_t *extra, int tid) {
      tid_ = tid;
      lv = reinterpret_cast<float *>(extra) + tid;
      li = reinterpret_cast<int *>(extra + (CM ? 0 : KMAX * NT * 4)) + tid;
      xs = reinterpret_cast<float *>(extra + LIST_BYTES);
      const int warp = tid >> 5;
      lane = tid & 31; q = warp & 3; cq = warp >> 2;
      row = blockIdx.y * 128 + q * 32 + lane;
      rv = row < p.B;
      m_run = REC_NEG_INF; s_run = 0.f; tgt = REC_NEG_INF; tau = REC_NEG_INF; cnt = 0;
      trow = (p.target && rv) ? (int)(p.target[row] - p.vocab_lo) : -1;
    }
    __device__ __forceinline__ void tile(const Params &p, int u, int, uint32_t tacc) {
      const int c_lo = u * 128 + cq * 32;
      float l[32];
      tc::tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 32), l);
      if (c_lo + 32 <= p.Vloc) {
        const float4 *bg = reinterpret_cast<const float4 *>(p.bias + c_lo);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(bg + (j >> 2));
          l[j] += b4.x; l[j + 1] += b4.y; l[j + 2] += b4.z; l[j + 3] += b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) l[j] = (c_lo + j < p.Vloc) ? l[j] + __ldg(p.bias + c_lo + j) : REC_NEG_INF;
      }
      float tmax = fmaxf(l[0], l[1]);
#pragma unroll
      for (int j = 2; j < 32; ++j) tmax = fmaxf(tmax, l[j]);
      {
        const float nm = fmaxf(m_run, tmax), nml = -fmaxf(nm, -1e30f) * LOG2E;
        float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 32; ++j) ps[j & 3] += tc::ex2_ftz(fmaf(l[j], LOG2E, nml));
        s_run = fmaf(s_run, tc::ex2_ftz((m_run - nm) * LOG2E), (ps[0] + ps[1]) + (ps[2] + ps[3]));
        m_run = nm;
      }
      if (trow >= c_lo && trow < c_lo + 32) {
        const int tj = trow - c_lo;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j == tj) tgt = l[j];
      }
      if (CM) {
        if (rv) p.cmax[(int64_t)row * p.cmax_ld + (u * CS + cq)] = tmax;
      } else if (tmax > tau) {  // rare once the list has warmed up: walk the chunk in ascending column order (ties: lowest id first)
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (l[j] > tau) topk_insert(l[j], p.vocab_lo + c_lo + j, lv, li, NT, p.topk, cnt, tau);
      }
    }
    __device__ __forceinline__ void finish(const Params &p) {
      float *x = xs + ((q * 32 + lane) * CS + cq) * 5;
      x[0] = m_run; x[1] = s_run; x[2] = tgt;
      if (!CM) for (int k = cnt; k < p.topk; ++k) { lv[k * NT] = REC_NEG_INF; li[k * NT] = 0x7fffffff; }
      epi_bar<NT>();
      if (cq != 0 || !rv) return;
      const float *y = xs + ((q * 32 + lane) * CS) * 5;
      float m = REC_NEG_INF, tg = REC_NEG_INF;
#pragma unroll
      for (int c = 0; c < CS; ++c) { m = fmaxf(m, y[c * 5]); tg = fmaxf(tg, y[c * 5 + 2]); }
      float ssum = 0.f;
#pragma unroll
      for (int c = 0; c < CS; ++c) if (y[c * 5 + 1] > 0.f) ssum += y[c * 5 + 1] * __expf(y[c * 5] - m);
      float *o = p.part + ((int64_t)blockIdx.x * p.B + row) * p.part_stride;
      o[0] = m; o[1] = ssum; o[2] = tg; o[3] = REC_NEG_INF; o[4] = __int_as_float(0x7fffffff);
      if (CM) return;
      // CS-way merge of the sorted private lists (ids of different quarters are disjoint); rec_kpub(topk, CS) >= topk
      // entries leave the CTA so that the fp32 re-score of the final merge has a margin of candidates
      const float *tvb = reinterpret_cast<const float *>(lv - tid_);
      const int *tib = reinterpret_cast<const int *>(li - tid_);
      const int kpub = rec_kpub(p.topk, CS);
      int pos[CS];
#pragma unroll
      for (int c = 0; c < CS; ++c) pos[c] = 0;
      for (int k = 0; k < kpub; ++k) {
        float cv = REC_NEG_INF;
        int ci = 0x7fffffff, cc = 0;
#pragma unroll
        for (int c = 0; c < CS; ++c) {
          if (pos[c] < p.topk) {
            const int t2 = (c * 4 + q) * 32 + lane;  // thread id of (q, cq = c, lane)
            const float v = tvb[pos[c] * NT + t2];
            const int i = tib[pos[c] * NT + t2];
            if (better(v, i, cv, ci)) { cv = v; ci = i; cc = c; }
          }
        }
#pragma unroll
        for (int c = 0; c < CS; ++c) if (c == cc) ++pos[c];
        o[TOPK_OFF + k] = cv;
        o[TOPK_OFF + REC_MAX_TOPK + k] = __int_as_float(ci);
      }
    }
  };
};

// The chunk-maxima pass over the FLATTENED unit space (session block, tile): every CTA gets the same number of units
// whatever the ratio of session blocks to SMs (a [splits x session blocks] grid leaves 28 of 148 SMs idle at 40 session
// blocks).  A CTA may cross session-block boundaries: it publishes one statistics record per block it touched, in slot
// (CTA index - first CTA of that block); untouched slots hold neutral records (fill_neutral_records_kernel).
__global__ void fill_neutral_records_kernel(float *__restrict__ part, int64_t n, int stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float *o = part + i * stride;
  o[0] = REC_NEG_INF; o[1] = 0.f; o[2] = REC_NEG_INF; o[3] = REC_NEG_INF; o[4] = __int_as_float(0x7fffffff);
}

struct HeadCmaxFlat {
  using Params = FwdParams;
  static constexpr bool CLUSTERED = false;
  static constexpr int EPI_WARPS = 16, NT = 512, CS = 4;
  static constexpr int RESIDENT_BYTES = 0;
  static constexpr const char *NAME = "tck:head_cmax_flat";
  static constexpr int STAGES = 3, STAGE_BYTES = 2 * BLK2, ACC_COLS = 128, TMEM_COLS = 256;
  static constexpr int EXTRA_BYTES = 128 * CS * 5 * 4;
  __device__ static __forceinline__ void units(const Params &p, int &lo, int &hi) {
    const int total = p.n_sb * p.n_tiles;
    lo = blockIdx.x * p.per;
    hi = min(total, lo + p.per);
    if (hi < lo) hi = lo;
  }
  __device__ static __forceinline__ int k_steps(const Params &p, int) { return p.KB; }
  __device__ static __forceinline__ void load(const Params &p, int u, int ks, uint8_t *stage, uint64_t *bar) {
    const int sb = u / p.n_tiles, t = u - sb * p.n_tiles;
    tc::mbar_expect_tx(bar, 2 * BLK2);
    tc::bulk_g2s(stage, p.himg + ((int64_t)sb * p.KB + ks) * BLK2, BLK2, bar);
    tc::bulk_g2s(stage + BLK2, p.wimg + ((int64_t)t * p.KB + ks) * BLK2, BLK2, bar);
  }
  __device__ static __forceinline__ void mma(const Params &, int, int, uint32_t st, uint32_t tacc, bool first) {
    HeadTopk<false>::mma_ab(st, st + BLK2, tacc, first);
  }
  struct Epi {
    float *xs;
    int q, cq, lane, row, trow, cur_sb;
    float m_run, s_run, tgt;
    bool rv;
    __device__ __forceinline__ Epi(const Params &, uint8_t *extra, int tid) {
      xs = reinterpret_cast<float *>(extra);
      const int warp = tid >> 5;
      lane = tid & 31; q = warp & 3; cq = warp >> 2;
      cur_sb = -1; row = 0; trow = -1; rv = false;
      m_run = REC_NEG_INF; s_run = 0.f; tgt = REC_NEG_INF;
    }
    __device__ __forceinline__ void flush(const Params &p) {
      float *x = xs + ((q * 32 + lane) * CS + cq) * 5;
      x[0] = m_run; x[1] = s_run; x[2] = tgt;
      epi_bar<NT>();
      if (cq == 0 && rv) {
        const float *y = xs + ((q * 32 + lane) * CS) * 5;
        float m = REC_NEG_INF, tg = REC_NEG_INF;
#pragma unroll
        for (int c = 0; c < CS; ++c) { m = fmaxf(m, y[c * 5]); tg = fmaxf(tg, y[c * 5 + 2]); }
        float ssum = 0.f;
#pragma unroll
        for (int c = 0; c < CS; ++c) if (y[c * 5 + 1] > 0.f) ssum += y[c * 5 + 1] * __expf(y[c * 5] - m);
        const int slot = (int)blockIdx.x - (cur_sb * p.n_tiles) / p.per;
        float *o = p.part + ((int64_t)slot * p.B + row) * p.part_stride;
        o[0] = m; o[1] = ssum; o[2] = tg;
      }
      epi_bar<NT>();  // xs is rewritten by the next flush
    }
    __device__ __forceinline__ void tile(const Params &p, int u, int, uint32_t tacc) {
      const int sb = u / p.n_tiles, t = u - sb * p.n_tiles;
      if (sb != cur_sb) {  // uniform across the CTA: every epilogue thread walks the same unit sequence
        if (cur_sb >= 0) flush(p);
        cur_sb = sb;
        row = sb * 128 + q * 32 + lane;
        rv = row < p.B;
        trow = (p.target && rv) ? (int)(p.target[row] - p.vocab_lo) : -1;
        m_run = REC_NEG_INF; s_run = 0.f; tgt = REC_NEG_INF;
      }
      const int c_lo = t * 128 + cq * 32;
      float l[32];
      tc::tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 32), l);
      if (c_lo + 32 <= p.Vloc) {
        const float4 *bg = reinterpret_cast<const float4 *>(p.bias + c_lo);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(bg + (j >> 2));
          l[j] += b4.x; l[j + 1] += b4.y; l[j + 2] += b4.z; l[j + 3] += b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) l[j] = (c_lo + j < p.Vloc) ? l[j] + __ldg(p.bias + c_lo + j) : REC_NEG_INF;
      }
      float tmax = fmaxf(l[0], l[1]);
#pragma unroll
      for (int j = 2; j < 32; ++j) tmax = fmaxf(tmax, l[j]);
      const float nm = fmaxf(m_run, tmax), nml = -fmaxf(nm, -1e30f) * LOG2E;
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 32; ++j) ps[j & 3] += tc::ex2_ftz(fmaf(l[j], LOG2E, nml));
      s_run = fmaf(s_run, tc::ex2_ftz((m_run - nm) * LOG2E), (ps[0] + ps[1]) + (ps[2] + ps[3]));
      m_run = nm;
      if (trow >= c_lo && trow < c_lo + 32) {
        const int tj = trow - c_lo;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j == tj) tgt = l[j];
      }
      if (rv) p.cmax[(int64_t)row * p.cmax_ld + (t * CS + cq)] = tmax;
    }
    __device__ __forceinline__ void finish(const Params &p) {
      if (cur_sb >= 0) flush(p);
    }
  };
};

// Bias operand blocks of HeadCmaxPair: block t = [128 items of tile t][K = 16] with (hi, lo, lo2) of the fp32 bias in
// k = 0..2 (the three bf16 terms add up to the fp32 value), zeros elsewhere; items beyond the catalogue get -1e30 (their
// weight rows are zero: the column vanishes from the maxima and the sum).  Block n_tiles is the "ones" operand.
__global__ void __launch_bounds__(128) pack_bias_blocks_kernel(const float *__restrict__ bias, int Vloc, int n_tiles,
                                                              uint8_t *__restrict__ out) {
  const int t = blockIdx.x, r = threadIdx.x;
  uint32_t *o = reinterpret_cast<uint32_t *>(out + (int64_t)t * 4096);
  // row r of the block: two 16-byte chunks (k = 0..7 at nosw_off(r, 0), k = 8..15 at nosw_off(r, 8))
  uint4 *c0 = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(o) + tc::nosw_off(r, 0));
  uint4 *c1 = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(o) + tc::nosw_off(r, 8));
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (t == n_tiles) {
    v.x = 0x3f803f80u; v.y = 0x00003f80u;  // bf16 1.0 in k = 0, 1, 2
  } else {
    const int item = t * 128 + r;
    const float b = item < Vloc ? bias[item] : -1e30f;
    const __nv_bfloat16 h0 = __float2bfloat16_rn(b);
    const float r1 = b - __bfloat162float(h0);
    const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(r1 - __bfloat162float(h1));
    v.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    v.y = (uint32_t)__bfloat16_as_ushort(h2);
  }
  *c0 = v;
  *c1 = make_uint4(0u, 0u, 0u, 0u);
}

// D = 64: the chunk-maxima pass with TWO session blocks per CTA.  Grid = (vocabulary splits, session-block pairs): a CTA
// keeps the state blocks of its pair RESIDENT in shared memory and streams every weight tile of its split ONCE for both
// blocks (HeadCmaxFlat re-loads a state block AND a weight tile per 128 x 128 unit: 64 KB from L2 per 768 tensor cycles,
// above the ~42 B/clk/SM the L2 delivers; and its CTAs walk the catalogue out of phase, so the 256 MB weight image is
// streamed from HBM once per session block).  All CTAs of a split sweep the same tiles at about the same time: the weight
// image is read from HBM once.  Per unit (tile) the issuer fills two accumulators (256 TMEM columns, double-buffered =
// all 512); 16 epilogue warps = (lane quadrant, session block of the pair, column half): 64 logits per thread and unit.
// Besides the 32-column chunk maxima every thread keeps the maximum of its 64 columns over 8 consecutive tiles: the
// level-2 maxima [B][2 * ceil(n_tiles / 8)] that chunk_select2_kernel reads first (a row's 31 k chunk maxima are then
// only touched inside the ~30 best groups).
// PV (REC_CMAX_EXP, experiments behind DESIGN.md 4.7): 0 = the product; 4 = every 4th exponential through tc::ex2_poly on
// the FMA pipe; 10 = no epilogue, 11 = no MMAs (which side of the pipeline bounds the kernel); 30 = a single issuer warp.
template <int PV>
struct HeadCmaxPairT {
  using Params = FwdParams;
  static constexpr bool CLUSTERED = false;
  static constexpr int EPI_WARPS = 16, NT = 512;
  static constexpr int ISSUERS = PV == 30 ? 1 : 2;
  static constexpr int BB = 4096;  // one un-swizzled [128][16] bf16 operand (tc::nosw_off)
  static constexpr int RESIDENT_BYTES = 2 * BLK2 + BB;           // state blocks of the pair + the "ones" operand
  static constexpr const char *NAME = "tck:head_cmax_pair";
  static constexpr int STAGES = 4, STAGE_BYTES = BLK2 + BB, ACC_COLS = 256, TMEM_COLS = 512;  // weight tile + its bias operand
  static constexpr int EXTRA_BYTES = 256 * 3 * 4;
  static constexpr int GROUP = 8;  // tiles per level-2 maximum
  __device__ static __forceinline__ void units(const Params &p, int &lo, int &hi) {
    lo = blockIdx.x * p.per;  // p.per % GROUP == 0: a level-2 group belongs to one CTA
    hi = min(p.n_tiles, lo + p.per);
    if (hi < lo) hi = lo;
  }
  __device__ static __forceinline__ int k_steps(const Params &, int) { return 1; }
  __device__ static void trace(const Params &p, int kind, int i) {
    if (p.trace && blockIdx.x == 1 && blockIdx.y == 1 && i >= 64 && i < 128) p.trace[kind * 64 + (i - 64)] = clock64();
  }
  __device__ static __forceinline__ bool has_second(const Params &p) { return 2 * (int)blockIdx.y + 1 < p.n_sb; }
  __device__ static __forceinline__ void load_resident(const Params &p, uint8_t *res, uint64_t *bar) {
    const uint32_t bytes = has_second(p) ? 2 * BLK2 : BLK2;  // consecutive blocks of the state image
    tc::mbar_expect_tx(bar, bytes + BB);
    tc::bulk_g2s(res, p.himg + (int64_t)(2 * blockIdx.y) * BLK2, bytes, bar);
    tc::bulk_g2s(res + 2 * BLK2, p.bblk + (int64_t)p.n_tiles * BB, BB, bar);
  }
  __device__ static __forceinline__ void load(const Params &p, int u, int, uint8_t *stage, uint64_t *bar) {
    tc::mbar_expect_tx(bar, BLK2 + BB);
    tc::bulk_g2s(stage, p.wimg + (int64_t)u * BLK2, BLK2, bar);
    tc::bulk_g2s(stage + BLK2, p.bblk + (int64_t)u * BB, BB, bar);
  }
  __device__ static __forceinline__ void mma(const Params &p, int, int, uint32_t st, uint32_t res, uint32_t tacc, bool) {
    if (PV == 11) return;  // timing experiment: no MMAs
    // + bias[item] for every session: ones[128][16] . (b_hi, b_lo, b_lo2, 0 ...)[128][16]^T, one K = 16 MMA per block
    const uint32_t id = tc::instr_desc(128, 128, 0, 0);
    const uint64_t ones = tc::smem_desc_nosw(res + 2 * BLK2, 128, 256), bb = tc::smem_desc_nosw(st + BLK2, 128, 256);
    HeadTopk<false>::mma_ab(res, st, tacc, true);
    tc::mma_bf16(tacc, ones, bb, id, true);
    if (has_second(p)) {
      HeadTopk<false>::mma_ab(res + BLK2, st, tacc + 128, true);
      tc::mma_bf16(tacc + 128, ones, bb, id, true);
    }
  }
  struct Epi {
    float *xs;
    int q, sbi, half, lane, row, trow;
    float m_run, s_run, tgt, g2;
    bool rv, active;
    __device__ __forceinline__ Epi(const Params &p, uint8_t *extra, int tid) {
      xs = reinterpret_cast<float *>(extra);
      const int warp = tid >> 5;
      lane = tid & 31; q = warp & 3; sbi = (warp >> 2) & 1; half = warp >> 3;
      const int sb = 2 * blockIdx.y + sbi;
      active = sb < p.n_sb;
      row = sb * 128 + q * 32 + lane;
      rv = active && row < p.B;
      trow = (p.target && rv) ? (int)(p.target[row] - p.vocab_lo) : -1;
      m_run = REC_NEG_INF; s_run = 0.f; tgt = REC_NEG_INF; g2 = REC_NEG_INF;
    }
    __device__ __forceinline__ float chunk(const Params &, int c_lo, uint32_t taddr) {
      float l[32];
      tc::tmem_ld32(taddr, l);
      // (the bias arrived through the tensor cores; columns beyond the catalogue carry -1e30)
      float tmax = fmaxf(l[0], l[1]);
#pragma unroll
      for (int j = 2; j < 32; ++j) tmax = fmaxf(tmax, l[j]);
      const float nm = fmaxf(m_run, tmax), nml = -fmaxf(nm, -1e30f) * LOG2E;
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = fmaf(l[j], LOG2E, nml);
        const bool poly = PV == 4 && (j & 3) == 3;
        ps[j & 3] += poly ? tc::ex2_poly(x) : tc::ex2_ftz(x);
      }
      s_run = fmaf(s_run, tc::ex2_ftz((m_run - nm) * LOG2E), (ps[0] + ps[1]) + (ps[2] + ps[3]));
      m_run = nm;
      if (trow >= c_lo && trow < c_lo + 32) {
        const int tj = trow - c_lo;
#pragma unroll
        for (int j = 0; j < 32; ++j) if (j == tj) tgt = l[j];
      }
      return tmax;
    }
    __device__ __forceinline__ void tile(const Params &p, int u, int, uint32_t tacc) {
      if (!active) return;  // odd number of session blocks: the second half of the last pair idles (warp-uniform)
      if (PV == 10) return;  // timing experiment: no epilogue
      const int c_lo = u * 128 + half * 64;
      const uint32_t ta = tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)(sbi * 128 + half * 64);
      const float c0 = chunk(p, c_lo, ta);
      const float c1 = chunk(p, c_lo + 32, ta + 32);
      g2 = fmaxf(g2, fmaxf(c0, c1));
      if (rv) {
        *reinterpret_cast<float2 *>(p.cmax + (int64_t)row * p.cmax_ld + (u * 4 + half * 2)) = make_float2(c0, c1);
        if ((u & (GROUP - 1)) == GROUP - 1 || u == p.n_tiles - 1) {
          p.cmax2[(int64_t)row * p.cmax2_ld + ((u / GROUP) * 2 + half)] = g2;
          g2 = REC_NEG_INF;
        }
      }
    }
    __device__ __forceinline__ void finish(const Params &p) {
      // one statistics record per (split, row): the two column halves meet in shared memory
      float *x = xs + ((sbi * 128 + q * 32 + lane)) * 3;
      if (half == 1) { x[0] = m_run; x[1] = s_run; x[2] = tgt; }
      epi_bar<NT>();
      if (half == 0 && rv) {
        const float m1 = x[0], s1 = x[1], t1 = x[2];
        const float m = fmaxf(m_run, m1);
        float ssum = 0.f;
        if (s_run > 0.f) ssum += s_run * __expf(m_run - m);
        if (s1 > 0.f) ssum += s1 * __expf(m1 - m);
        float *o = p.part + ((int64_t)blockIdx.x * p.B + row) * p.part_stride;
        o[0] = m; o[1] = ssum; o[2] = fmaxf(tgt, t1);
      }
    }
  };
};

using HeadCmaxPair = HeadCmaxPairT<0>;

// ---- exact top-k from chunk maxima ------------------------------------------------------------------------------------
constexpr int KC = 24;      // chunks kept per row: k <= 20 plus a margin for the ~1e-5 relative error of the bf16x3 maxima
constexpr int CCAP = 512;   // candidate list of the threshold pass (expected ~50 entries)

// warp = row, two kernels (selection: few registers, many warps; scoring: register-heavy, 16 loads in flight per lane).
//  (1) lane-strided pass over the row's chunk maxima: lane maximum; t0 = KC-th largest of the 32 lane maxima -- at least
//      KC chunks reach t0, so the KC best chunks all do;
//  (2) second pass (L2 hits): chunks >= t0 are appended to a shared-memory list (ballot + prefix); KC rounds of a warp
//      argmax pick the KC best by (maximum desc, chunk asc); a list overflow (a row of massive ties) falls back to KC
//      full passes;
//  (3) the 32 columns of every kept chunk are scored EXACTLY in fp32 (lane = column: h . W[v] in ascending k with one
//      accumulator, + bias -- nn.Linear up to summation order), and the top-k of the KC x 32 exact scores by
//      (score desc, id asc) is torch.topk on fp32 logits (eval_protocol.py:75).
// Writes row_ids / row_topv like head_merge_kernel, and the candidate slots of `summary` when this shard's result
// travels on to a cross-shard merge.
// Selection kernel (few registers: 64 warps per SM hide the latency of streaming the maxima).
__global__ void __launch_bounds__(256) chunk_select_kernel(const float *__restrict__ cmax, int64_t ld, int n_chunks, int B,
                                                          int *__restrict__ chosen) {
  __shared__ float lv_all[8 * CCAP];
  __shared__ int lc_all[8 * CCAP];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + wid;
  if (row >= B) return;
  float *lv = lv_all + wid * CCAP;
  int *lc = lc_all + wid * CCAP;
  const float *cm = cmax + (int64_t)row * ld;
  // (1) lane maxima -> t0  (8 independent 128-byte loads per warp and iteration)
  float lm = REC_NEG_INF;
  for (int c = lane; c < n_chunks; c += 32 * 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = c + 32 * j < n_chunks ? __ldg(cm + c + 32 * j) : REC_NEG_INF;
#pragma unroll
    for (int j = 0; j < 8; ++j) lm = fmaxf(lm, v[j]);
  }
  int rank = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float o = __shfl_sync(0xffffffffu, lm, j);
    rank += (o > lm || (o == lm && j < lane)) ? 1 : 0;
  }
  const int src = __ffs(__ballot_sync(0xffffffffu, rank == KC - 1)) - 1;
  const float t0 = __shfl_sync(0xffffffffu, lm, src);
  // (2) candidates >= t0
  int n_list = 0;
  bool overflow = false;
  for (int c0 = 0; c0 < n_chunks && !overflow; c0 += 32 * 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = c0 + 32 * j + lane < n_chunks ? __ldg(cm + c0 + 32 * j + lane) : REC_NEG_INF;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + 32 * j + lane;
      const bool take = c < n_chunks && v[j] >= t0;
      const unsigned m = __ballot_sync(0xffffffffu, take);
      if (m) {
        const int pos = n_list + __popc(m & ((1u << lane) - 1u));
        if (take && pos < CCAP) { lv[pos] = v[j]; lc[pos] = c; }
        n_list += __popc(m);
        if (n_list > CCAP) overflow = true;
      }
    }
  }
  __syncwarp();
  float lastv = 3.402823466e+38f;
  int lastc = -1;
  for (int r = 0; r < KC; ++r) {
    float bv = REC_NEG_INF;
    int bc = 0x7fffffff;
    if (!overflow) {
      for (int e = lane; e < n_list; e += 32) {
        const float v = lv[e];
        const int c = lc[e];
        if (better(lastv, lastc, v, c) && better(v, c, bv, bc)) { bv = v; bc = c; }
      }
    } else {
      for (int c = lane; c < n_chunks; c += 32) {
        const float v = __ldg(cm + c);
        if (better(lastv, lastc, v, c) && better(v, c, bv, bc)) { bv = v; bc = c; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (better(ov, oc, bv, bc)) { bv = ov; bc = oc; }
    }
    if (lane == 0) chosen[(int64_t)row * KC + r] = bc;
    lastv = bv; lastc = bc;
  }
}

template <int D4>  // D / 4 when the state row fits in registers (D = 64), else 0 (row in shared memory)
__global__ void __launch_bounds__(256) chunk_score_kernel(const int *__restrict__ chosen,
                                                         const float *__restrict__ W, const float *__restrict__ bias,
                                                         const float *__restrict__ h, int B, int D, int Vloc, int vocab_lo,
                                                         int topk, int32_t *__restrict__ row_ids, float *__restrict__ row_topv,
                                                         float *__restrict__ summary, int part_stride) {
  extern __shared__ float smem_f[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + wid;
  if (row >= B) return;
  float *hs = smem_f + wid * D;               // state row (D4 == 0)
  const int mine = lane < KC ? chosen[(int64_t)row * KC + lane] : 0x7fffffff;  // lane r holds the r-th best chunk
  // (3) exact scores of the 32 columns of every kept chunk (lane = column)
  float4 hr[D4 > 0 ? D4 : 1];
  if (D4 > 0) {
#pragma unroll
    for (int k4 = 0; k4 < D4; ++k4) hr[k4] = __ldg(reinterpret_cast<const float4 *>(h + (int64_t)row * D) + k4);
  } else {
    for (int k = lane; k < D; k += 32) hs[k] = h[(int64_t)row * D + k];
    __syncwarp();
  }
  float sc[KC];
  int ids[KC];
#pragma unroll
  for (int r = 0; r < KC; ++r) {
    const int c = __shfl_sync(0xffffffffu, mine, r);
    const int v = c == 0x7fffffff ? Vloc : c * 32 + lane;
    float acc = REC_NEG_INF;
    ids[r] = 0x7fffffff;
    if (v < Vloc) {
      const float4 *wr = reinterpret_cast<const float4 *>(W + (int64_t)v * D);
      acc = 0.f;
      if (D4 > 0) {
        float4 w4[D4 > 0 ? D4 : 1];
#pragma unroll
        for (int k4 = 0; k4 < D4; ++k4) w4[k4] = __ldg(wr + k4);
#pragma unroll
        for (int k4 = 0; k4 < D4; ++k4) {
          acc = fmaf(hr[k4].x, w4[k4].x, acc); acc = fmaf(hr[k4].y, w4[k4].y, acc);
          acc = fmaf(hr[k4].z, w4[k4].z, acc); acc = fmaf(hr[k4].w, w4[k4].w, acc);
        }
      } else {
        for (int k4 = 0; k4 < (D >> 2); ++k4) {
          const float4 w4 = __ldg(wr + k4);
          const float4 h4 = *reinterpret_cast<const float4 *>(hs + 4 * k4);
          acc = fmaf(h4.x, w4.x, acc); acc = fmaf(h4.y, w4.y, acc); acc = fmaf(h4.z, w4.z, acc); acc = fmaf(h4.w, w4.w, acc);
        }
      }
      acc += __ldg(bias + v);
      ids[r] = vocab_lo + v;
    }
    sc[r] = acc;
  }
  float pv = 3.402823466e+38f;
  int pi = -1;
  float *sm = summary ? summary + (int64_t)row * part_stride : nullptr;
  for (int k = 0; k < topk; ++k) {
    float bv = REC_NEG_INF;
    int bi = 0x7fffffff;
#pragma unroll
    for (int r = 0; r < KC; ++r)
      if (ids[r] != 0x7fffffff && better(pv, pi, sc[r], ids[r]) && better(sc[r], ids[r], bv, bi)) { bv = sc[r]; bi = ids[r]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      row_ids[(int64_t)row * REC_MAX_TOPK + k] = bi; row_topv[(int64_t)row * REC_MAX_TOPK + k] = bv;
      if (sm) { sm[TOPK_OFF + k] = bv; sm[TOPK_OFF + REC_MAX_TOPK + k] = __int_as_float(bi); }
    }
    pv = bv; pi = bi;
  }
}

// Selection over the two-level maxima written by HeadCmaxPair (warp = row):
//  (1) lane-strided pass over the row's LEVEL-2 maxima (group = 64 columns x 8 tiles = 16 chunks); t0 = KC-th largest of
//      the 32 lane maxima: at least KC groups reach t0, every one of them holds a chunk >= t0, so the KC best chunks all
//      reach t0 and lie in groups >= t0;
//  (2) second pass over the level-2 row (L1 / L2 hits): for every group >= t0 sixteen lanes fetch its chunk maxima and
//      the chunks >= t0 go to the shared-memory list; KC rounds of a warp argmax pick the KC best by (maximum desc,
//      chunk asc).  A list overflow (a row of massive ties) falls back to KC full passes over the chunk maxima.
__global__ void __launch_bounds__(256) chunk_select2_kernel(const float *__restrict__ cmax, int64_t ld, int n_chunks,
                                                           const float *__restrict__ cmax2, int64_t ld2, int n_tiles, int B,
                                                           int *__restrict__ chosen) {
  constexpr int GROUP = HeadCmaxPair::GROUP;
  __shared__ float lv_all[8 * CCAP];
  __shared__ int lc_all[8 * CCAP];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + wid;
  if (row >= B) return;
  float *lv = lv_all + wid * CCAP;
  int *lc = lc_all + wid * CCAP;
  const float *cm = cmax + (int64_t)row * ld;
  const float *cm2 = cmax2 + (int64_t)row * ld2;
  const int n2 = ((n_tiles + GROUP - 1) / GROUP) * 2;
  float lm = REC_NEG_INF;
  for (int g = lane; g < n2; g += 32 * 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = g + 32 * j < n2 ? __ldg(cm2 + g + 32 * j) : REC_NEG_INF;
#pragma unroll
    for (int j = 0; j < 8; ++j) lm = fmaxf(lm, v[j]);
  }
  int rank = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float o = __shfl_sync(0xffffffffu, lm, j);
    rank += (o > lm || (o == lm && j < lane)) ? 1 : 0;
  }
  const int src = __ffs(__ballot_sync(0xffffffffu, rank == KC - 1)) - 1;
  const float t0 = __shfl_sync(0xffffffffu, lm, src);
  int n_list = 0;
  bool overflow = false;
  for (int g0 = 0; g0 < n2 && !overflow; g0 += 32) {
    const float v2 = g0 + lane < n2 ? __ldg(cm2 + g0 + lane) : REC_NEG_INF;
    unsigned gm = __ballot_sync(0xffffffffu, v2 >= t0);
    while (gm && !overflow) {
      // two groups per iteration: lanes 0..15 the lowest set bit, lanes 16..31 the next one
      const int b0 = __ffs(gm) - 1;
      gm &= gm - 1;
      const int b1 = gm ? __ffs(gm) - 1 : -1;
      if (gm) gm &= gm - 1;
      const int bsel = lane < 16 ? b0 : b1;
      const int g = g0 + bsel, l16 = lane & 15;
      const int t = (g >> 1) * GROUP + (l16 >> 1);
      const int c = t * 4 + (g & 1) * 2 + (l16 & 1);
      float v = REC_NEG_INF;
      if (bsel >= 0 && t < n_tiles) v = __ldg(cm + c);
      const bool take = v >= t0 && bsel >= 0 && t < n_tiles;
      const unsigned m = __ballot_sync(0xffffffffu, take);
      if (m) {
        const int pos = n_list + __popc(m & ((1u << lane) - 1u));
        if (take && pos < CCAP) { lv[pos] = v; lc[pos] = c; }
        n_list += __popc(m);
        if (n_list > CCAP) overflow = true;
      }
    }
  }
  __syncwarp();
  float lastv = 3.402823466e+38f;
  int lastc = -1;
  for (int r = 0; r < KC; ++r) {
    float bv = REC_NEG_INF;
    int bc = 0x7fffffff;
    if (!overflow) {
      for (int e = lane; e < n_list; e += 32) {
        const float v = lv[e];
        const int c = lc[e];
        if (better(lastv, lastc, v, c) && better(v, c, bv, bc)) { bv = v; bc = c; }
      }
    } else {
      for (int c = lane; c < n_chunks; c += 32) {
        const float v = __ldg(cm + c);
        if (better(lastv, lastc, v, c) && better(v, c, bv, bc)) { bv = v; bc = c; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (better(ov, oc, bv, bc)) { bv = ov; bc = oc; }
    }
    if (lane == 0) chosen[(int64_t)row * KC + r] = bc;
    lastv = bv; lastc = bc;
  }
}

// Exact scores of the kept chunks with COALESCED loads (D = 64): the 32 weight rows of a chunk are 8 KB of contiguous
// memory; load instruction i of a warp covers rows 2i and 2i + 1 (lane = 16 * (row & 1) + float4 index), every lane
// multiplies by ITS float4 of the state row and the sixteen partial sums of a row meet in a fixed xor butterfly -- the
// same order for every row, so identical weight rows score identically wherever they sit.  (The lane-per-column version
// above touches 32 different 128-byte lines per load instruction: 8 x the L1 tag traffic, 380 us per 5000 rows.)
// Lane j ends up with the score of column 2 * (j & 15) + (j >> 4) of each chunk.
__global__ void __launch_bounds__(128) chunk_score64_kernel(const int *__restrict__ chosen,
                                                           const float *__restrict__ W, const float *__restrict__ bias,
                                                           const float *__restrict__ h, int B, int Vloc, int vocab_lo,
                                                           int topk, int32_t *__restrict__ row_ids, float *__restrict__ row_topv,
                                                           float *__restrict__ summary, int part_stride) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int row = blockIdx.x * 4 + wid;
  if (row >= B) return;
  const int mine = lane < KC ? chosen[(int64_t)row * KC + lane] : 0x7fffffff;  // lane r holds the r-th best chunk
  const float4 hq = __ldg(reinterpret_cast<const float4 *>(h + (int64_t)row * 64) + (lane & 15));
  const int col = 2 * (lane & 15) + (lane >> 4);
  float sc[KC];
  int ids[KC];
#pragma unroll
  for (int r = 0; r < KC; ++r) {
    const int c = __shfl_sync(0xffffffffu, mine, r);
    float my = REC_NEG_INF;
    ids[r] = 0x7fffffff;
    if (c != 0x7fffffff) {  // warp-uniform
      const int v0 = c * 32;
      const float4 *wc = reinterpret_cast<const float4 *>(W + (int64_t)v0 * 64) + lane;
      float4 w4[16];
      if (v0 + 32 <= Vloc) {
#pragma unroll
        for (int i = 0; i < 16; ++i) w4[i] = __ldg(wc + i * 32);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          w4[i] = (v0 + 2 * i + (lane >> 4) < Vloc) ? __ldg(wc + i * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float pr = w4[i].x * hq.x;
        pr = fmaf(w4[i].y, hq.y, pr); pr = fmaf(w4[i].z, hq.z, pr); pr = fmaf(w4[i].w, hq.w, pr);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) pr += __shfl_xor_sync(0xffffffffu, pr, o);
        if ((lane & 15) == i) my = pr;
      }
      const int v = v0 + col;
      if (v < Vloc) { my += __ldg(bias + v); ids[r] = vocab_lo + v; }
      else my = REC_NEG_INF;
    }
    sc[r] = my;
  }
  // top-k of the KC x 32 exact scores by (score desc, id asc).  Every lane keeps the best of ITS candidates; a round is a
  // warp argmax over the 32 lane bests, and only the winning lane rescans its registers (ids are unique: a column is
  // scored once).  (Rescanning all KC candidates in every lane and round was half of this kernel's instructions.)
  float lbv = REC_NEG_INF;
  int lbi = 0x7fffffff;
#pragma unroll
  for (int r = 0; r < KC; ++r)
    if (ids[r] != 0x7fffffff && better(sc[r], ids[r], lbv, lbi)) { lbv = sc[r]; lbi = ids[r]; }
  float *sm = summary ? summary + (int64_t)row * part_stride : nullptr;
  for (int k = 0; k < topk; ++k) {
    float bv = lbv;
    int bi = lbi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      row_ids[(int64_t)row * REC_MAX_TOPK + k] = bi; row_topv[(int64_t)row * REC_MAX_TOPK + k] = bv;
      if (sm) { sm[TOPK_OFF + k] = bv; sm[TOPK_OFF + REC_MAX_TOPK + k] = __int_as_float(bi); }
    }
    if (bi == lbi && bi != 0x7fffffff) {  // this lane held the winner: retire it, find the lane's next best
      lbv = REC_NEG_INF; lbi = 0x7fffffff;
#pragma unroll
      for (int r = 0; r < KC; ++r) {
        if (ids[r] == bi) ids[r] = 0x7fffffff;
        if (ids[r] != 0x7fffffff && better(sc[r], ids[r], lbv, lbi)) { lbv = sc[r]; lbi = ids[r]; }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// dh[128 sessions x (64 NCB)] = sum over this CTA's vocabulary range of dl . W   (A, B: MN-major views)
// ------------------------------------------------------------------------------------------------------------
struct DhParams {
  const uint8_t *dlT, *wimg;
  int KB, NCB, dl_cb, n_tiles, B, D;
  float *dh_part;  // [n_split][B][D]
};

struct HeadDh {
  using Params = DhParams;
  static constexpr bool CLUSTERED = false;
  static constexpr int RESIDENT_BYTES = 0;
  static constexpr int EPI_WARPS = 8;
  static constexpr const char *NAME = "tck:head_dh";
  static constexpr int STAGES = 2, STAGE_BYTES = 12 * HALF, ACC_COLS = 256, TMEM_COLS = 256;
  static constexpr int EXTRA_BYTES = 0;
  __device__ static __forceinline__ void range(const Params &p, int &t_lo, int &t_hi) {
    const int per = (p.n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    t_lo = blockIdx.x * per;
    t_hi = min(p.n_tiles, t_lo + per);
  }
  __device__ static __forceinline__ void units(const Params &, int &lo, int &hi) { lo = 0; hi = 1; }
  __device__ static __forceinline__ int k_steps(const Params &p, int) {
    int t_lo, t_hi;
    range(p, t_lo, t_hi);
    return max(0, t_hi - t_lo) * 2;
  }
  __device__ static __forceinline__ void load(const Params &p, int, int ks, uint8_t *stage, uint64_t *bar) {
    int t_lo, t_hi;
    range(p, t_lo, t_hi);
    const int t = t_lo + (ks >> 1), rh = ks & 1;
    tc::mbar_expect_tx(bar, (uint32_t)(4 + 2 * p.NCB) * HALF);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const uint8_t *blk = p.dlT + ((int64_t)t * p.dl_cb + 2 * blockIdx.z + c) * BLK2 + rh * HALF;
      tc::bulk_g2s(stage + c * HALF, blk, HALF, bar);
      tc::bulk_g2s(stage + 2 * HALF + c * HALF, blk + BLK, HALF, bar);
    }
    for (int c = 0; c < p.NCB; ++c) {
      const uint8_t *blk = p.wimg + ((int64_t)t * p.KB + blockIdx.y * p.NCB + c) * BLK2 + rh * HALF;
      tc::bulk_g2s(stage + 4 * HALF + c * HALF, blk, HALF, bar);
      tc::bulk_g2s(stage + 8 * HALF + c * HALF, blk + BLK, HALF, bar);
    }
  }
  __device__ static __forceinline__ void mma(const Params &p, int, int, uint32_t st, uint32_t tacc, bool first) {
    const uint32_t id = tc::instr_desc(128, 64 * p.NCB, 1, 1);
    const uint64_t ah = tc::desc_mnmajor(st, 0, HALF), al = tc::desc_mnmajor(st + 2 * HALF, 0, HALF);
    const uint64_t bh = tc::desc_mnmajor(st + 4 * HALF, 0, HALF), bl = tc::desc_mnmajor(st + 8 * HALF, 0, HALF);
    bool acc = !first;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const uint64_t a = pass == 2 ? al : ah, b = pass == 1 ? bl : bh;
#pragma unroll
      for (int k = 0; k < 4; ++k) { tc::mma_bf16(tacc, a + (uint64_t)(k * 128), b + (uint64_t)(k * 128), id, acc); acc = true; }
    }
  }
  struct Epi {
    int q, cq, lane;
    __device__ __forceinline__ Epi(const Params &, uint8_t *, int tid) {
      const int warp = tid >> 5;
      lane = tid & 31; q = warp & 3; cq = warp >> 2;
    }
    __device__ __forceinline__ void tile(const Params &p, int, int, uint32_t tacc) {
      const int row = blockIdx.z * 128 + q * 32 + lane;
      const int N = 64 * p.NCB, nch = N / 64;  // 32-column chunks per column half
      float *dst = p.dh_part + ((int64_t)blockIdx.x * p.B + row) * p.D + blockIdx.y * N;
      for (int ch = 0; ch < nch; ++ch) {
        const int c0 = cq * (N / 2) + ch * 32;
        float g[32];
        tc::tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, g);
        if (row < p.B) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4 *>(dst + c0 + j) = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
        }
      }
    }
    __device__ __forceinline__ void finish(const Params &) {}
  };
};

// ------------------------------------------------------------------------------------------------------------
// dW^T tile [128 state columns x 256 items] = h^T . dl (K = batch) -> Adam on W, m, v (+ bias for the first column tile).
// The tile is computed TRANSPOSED (TMEM lane = state column d, TMEM column = item row v): a warp's global accesses to
// W[v, d0 .. d0 + 32) are then 128 contiguous bytes per instruction with no shared-memory transpose, and the four
// lane-quarter warps cover 512 contiguous bytes of every weight row.
// ------------------------------------------------------------------------------------------------------------
struct DwParams {
  const uint8_t *dlT, *hT;  // [n_tiles][KBS], [D/128][KBS]
  int KBS, n_vt, n_dt, n_tiles, Vloc, D, n_sb;
  float *w, *wm, *wv, *b, *bm, *bv;
  const float *db_part;     // [n_sb][n_tiles * 128]
  float b1, b2, eps, step_size, inv_bc2_sqrt;
  const float *sc;
};

struct HeadDwAdam {
  using Params = DwParams;
  static constexpr bool CLUSTERED = false;
  static constexpr int RESIDENT_BYTES = 0;
  static constexpr int EPI_WARPS = 8;
  static constexpr const char *NAME = "tck:head_dw_adam";
  static constexpr int STAGES = 2, STAGE_BYTES = 3 * BLK2, ACC_COLS = 256, TMEM_COLS = 512;
  static constexpr int EXTRA_BYTES = 0;
  __device__ static __forceinline__ void units(const Params &p, int &lo, int &hi) {
    const int total = p.n_vt * p.n_dt;
    const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
    lo = blockIdx.x * per;
    hi = min(total, lo + per);
    if (hi < lo) hi = lo;
  }
  __device__ static __forceinline__ int k_steps(const Params &p, int) { return p.KBS; }
  __device__ static __forceinline__ void load(const Params &p, int u, int ks, uint8_t *stage, uint64_t *bar) {
    const int vt = u / p.n_dt, dt = u - vt * p.n_dt;
    tc::mbar_expect_tx(bar, 3 * BLK2);
    tc::bulk_g2s(stage, p.hT + ((int64_t)dt * p.KBS + ks) * BLK2, BLK2, bar);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int t = min(2 * vt + i, p.n_tiles - 1);  // an odd tile count: the duplicate's rows lie beyond Vloc
      const uint8_t *blk = p.dlT + ((int64_t)t * p.KBS + ks) * BLK2;
      tc::bulk_g2s(stage + BLK2 + i * BLK, blk, BLK, bar);
      tc::bulk_g2s(stage + BLK2 + 2 * BLK + i * BLK, blk + BLK, BLK, bar);
    }
  }
  __device__ static __forceinline__ void mma(const Params &, int, int, uint32_t st, uint32_t tacc, bool first) {
    const uint32_t id = tc::instr_desc(128, 256, 0, 0);
    const uint64_t ah = tc::desc_kmajor(st, 0), al = tc::desc_kmajor(st + BLK, 0);
    const uint64_t bh = tc::desc_kmajor(st + BLK2, 0), bl = tc::desc_kmajor(st + BLK2 + 2 * BLK, 0);
    bool acc = !first;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const uint64_t a = pass == 2 ? al : ah, b = pass == 1 ? bl : bh;
#pragma unroll
      for (int k = 0; k < 4; ++k) { tc::mma_bf16(tacc, a + (uint64_t)(k * 2), b + (uint64_t)(k * 2), id, acc); acc = true; }
    }
  }
  struct Epi {
    int q, cq, lane;
    float step_size, inv_bc2_sqrt;
    __device__ __forceinline__ Epi(const Params &p, uint8_t *, int tid) {
      const int warp = tid >> 5;
      lane = tid & 31; q = warp & 3; cq = warp >> 2;
      step_size = p.sc ? p.sc[0] : p.step_size;
      inv_bc2_sqrt = p.sc ? p.sc[1] : p.inv_bc2_sqrt;
    }
    __device__ __forceinline__ void tile(const Params &p, int u, int, uint32_t tacc) {
      const int vt = u / p.n_dt, dt = u - vt * p.n_dt;
      const int d = dt * 128 + q * 32 + lane;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        const int v0 = vt * 256 + cq * 128 + ch * 32;
        if (v0 >= p.Vloc) break;  // warp-uniform
        float g[32];
        tc::tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 128 + ch * 32), g);
        const int nrows = min(32, p.Vloc - v0);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float P[16], M[16], U[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int64_t off = (int64_t)(v0 + min(hf * 16 + j, nrows - 1)) * p.D + d;
            P[j] = p.w[off]; M[j] = p.wm[off]; U[j] = p.wv[off];
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            adam_elem(P[j], M[j], U[j], g[hf * 16 + j], p.b1, p.b2, p.eps, step_size, inv_bc2_sqrt);
            if (hf * 16 + j < nrows) {
              const int64_t off = (int64_t)(v0 + hf * 16 + j) * p.D + d;
              p.w[off] = P[j]; p.wm[off] = M[j]; p.wv[off] = U[j];
            }
          }
        }
      }
      if (dt == 0) {  // bias of this unit's 256 rows: gradient = sum of the per-session-block partials
        const int v = vt * 256 + cq * 128 + q * 32 + lane;
        if (v < p.Vloc) {
          float g = 0.f;
          for (int sb = 0; sb < p.n_sb; ++sb) g += p.db_part[((int64_t)sb * p.n_tiles + (v >> 7)) * 128 + (v & 127)];
          float bp = p.b[v], m = p.bm[v], vv = p.bv[v];
          adam_elem(bp, m, vv, g, p.b1, p.b2, p.eps, step_size, inv_bc2_sqrt);
          p.b[v] = bp; p.bm[v] = m; p.bv[v] = vv;
        }
      }
    }
    __device__ __forceinline__ void finish(const Params &) {}
  };
};

}  // namespace tck

// ------------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------------
bool tck_heads_supported(const rec_engine *e) { return e->use_tc && e->D >= 128 && e->D % 128 == 0 && e->D <= 1024; }
// statistics + running top-k over operand images: any D % 64 == 0 (D = 64: evaluation-shaped k only -- the training
// steps' k <= 2 stay with head_stats_tc_kernel, which keeps them in registers)
bool tck_topk_supported(const rec_engine *e, const HeadStatsArgs &a) {
  static const int off = getenv("REC_NO_TCK_TOPK") ? 1 : 0;
  if (off || !e->use_tc || a.n_arg > 0 || a.topk < 1 || a.topk > tck::HeadTopk<true>::KMAX) return false;
  if (e->D == 64) return a.topk > 8;
  return e->D % 128 == 0 && e->D <= 1024;
}

static int tck_alloc(rec_engine *e, void **ptr, size_t bytes) {
  if (*ptr) return REC_OK;
  cudaError_t st = cudaMalloc(ptr, bytes);
  if (st != cudaSuccess) REC_FAIL(e, REC_ENOMEM, "cudaMalloc(%zu B) for the tensor-core head images failed: %s", bytes, cudaGetErrorString(st));
  REC_CUDA(e, cudaMemsetAsync(*ptr, 0, bytes, e->stream));
  return REC_OK;
}

// Operand images are sized by max_batch / Vloc / D at first use and live as long as the engine.
// what: 1 = forward (weight image 0, state image 0), 2 = greedy-action pass (image 1), 4 = backward (dl^T, h^T, db)
static int tck_ensure(rec_engine *e, int what) {
  const int KB = e->D / 64, n_tiles = cdiv(e->Vloc, 128), n_sb = cdiv(e->cfg.max_batch, 128);
  int rc;
  const size_t wbytes = (size_t)n_tiles * KB * tck::BLK2, hbytes = (size_t)n_sb * KB * tck::BLK2;
  if (what & 1) {
    if ((rc = tck_alloc(e, (void **)&e->k_wimg[0], wbytes))) return rc;
    if ((rc = tck_alloc(e, (void **)&e->k_himg[0], hbytes))) return rc;
  }
  if (what & 2) {
    if ((rc = tck_alloc(e, (void **)&e->k_wimg[1], wbytes))) return rc;
    if ((rc = tck_alloc(e, (void **)&e->k_himg[1], hbytes))) return rc;
    if ((rc = tck_alloc(e, (void **)&e->k_bias, sizeof(float) * (size_t)n_tiles * 128))) return rc;
  }
  if (what & 4) {
    if ((rc = tck_alloc(e, (void **)&e->k_hT, (size_t)(e->D / 128) * (2 * n_sb) * tck::BLK2))) return rc;
    if ((rc = tck_alloc(e, (void **)&e->k_dlT, (size_t)n_tiles * (2 * n_sb) * tck::BLK2))) return rc;
    if ((rc = tck_alloc(e, (void **)&e->k_db, sizeof(float) * (size_t)n_sb * n_tiles * 128))) return rc;
  }
  return REC_OK;
}

void tck_free(rec_engine *e) {
  void *ptrs[] = {e->k_wimg[0], e->k_wimg[1], e->k_himg[0], e->k_himg[1], e->k_hT, e->k_dlT, e->k_db, e->k_bias, e->k_cmax, e->k_cmax2, e->k_chosen, e->k_bblk};
  for (void *p : ptrs) if (p) cudaFree(p);
}

static int tck_pack(rec_engine *e, const tck::PackSrc &s, int R, int C, uint8_t *img) {
  const int64_t n_chunks = (int64_t)cdiv(R, 128) * 128 * (C / 8);
  int64_t blocks = cdiv64(n_chunks, 256);
  const int64_t cap = (int64_t)e->sm_count * 4;
  if (blocks > cap) blocks = cap;
  tck::pack_img_kernel<<<(int)blocks, 256, 0, e->stream>>>(s, R, C, C, img);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

// Weight image of one head (head >= 0 -> k_wimg[0]) or of the pre-combined greedy-action heads sum_j w_j Q_j
// (head < 0 -> k_wimg[1], + combined bias).
static int tck_pack_head_image(rec_engine *e, int net_id, int head, int n_arg, const float *w, int arg_shift = 0) {
  const rec_net_params &np = e->nets[net_id].p;
  tck::PackSrc ws = {};
  if (head < 0) {
    ws.n = n_arg;
    for (int j = 0; j < n_arg; ++j) { ws.p[j] = np.head_w[1 + arg_shift + j]; ws.w[j] = n_arg > 1 ? w[j] : 1.f; }
    if (n_arg > 1) {
      tck::PackSrc bs = ws;
      for (int j = 0; j < n_arg; ++j) bs.p[j] = np.head_b[1 + arg_shift + j];
      tck::bias_combine_kernel<<<cdiv(e->Vloc, 256), 256, 0, e->stream>>>(bs, e->Vloc, e->k_bias);
      REC_LAUNCH_CHECK(e);
    }
    return tck_pack(e, ws, e->Vloc, e->D, e->k_wimg[1]);
  }
  ws.n = 1; ws.p[0] = np.head_w[head]; ws.w[0] = 1.f;
  int rc = tck_pack(e, ws, e->Vloc, e->D, e->k_wimg[0]);
  if (rc) return rc;
  e->k_sup_net = net_id; e->k_sup_head = head;
  return REC_OK;
}

// The weight images only depend on the parameters: the fused train steps produce them on a side stream while the
// GRU forward runs (HBM-bound packing next to the latency-bound recurrence).
int tck_prepack_heads(rec_engine *e, int net_id, int n_arg, const float *w) {
  if (!tck_heads_supported(e)) return REC_OK;
  int rc = tck_ensure(e, n_arg > 0 ? 3 : 1);
  if (rc) return rc;
  if ((rc = tck_pack_head_image(e, net_id, 0, 0, w))) return rc;
  e->k_fresh[0] = true;
  if (n_arg > 0) {
    if ((rc = tck_pack_head_image(e, net_id, -1, n_arg, w))) return rc;
    e->k_fresh[1] = true;
  }
  return REC_OK;
}

// Evaluation-shaped batches over a large catalogue: statistics + chunk maxima on the tensor cores, then the exact top-k
// from the best chunks (see HeadTopk<.., CM>).  Writes the per-(split, row) statistics records (merged by the caller with
// launch_head_merge(topk = 0)) and row_ids / row_topv (+ the candidate slots of `summary`).
bool tck_chunk_topk_supported(const rec_engine *e, const HeadStatsArgs &a) {
  static const int off = getenv("REC_NO_CHUNK_TOPK") ? 1 : 0;
  static const int min_b = getenv("REC_CHUNK_MIN_B") ? atoi(getenv("REC_CHUNK_MIN_B")) : 1024;
  return !off && tck_topk_supported(e, a) && a.B >= min_b && e->Vloc >= 32768;
}

int launch_head_topk_chunks(rec_engine *e, const HeadStatsArgs &a, int *n_split_out, float *summary) {
  int rc = tck_ensure(e, 1);
  if (rc) return rc;
  const rec_net_params &np = e->nets[a.net_id].p;
  const int KB = e->D / 64, n_tiles = cdiv(e->Vloc, 128), n_sb = cdiv(a.B, 128), n_chunks = n_tiles * 4;
  const int64_t mb = e->cfg.max_batch;
  const int64_t cld = (n_chunks + 31) / 32 * 32;  // row pitch of the chunk maxima: whole 128-byte lines
  if ((rc = tck_alloc(e, (void **)&e->k_cmax, sizeof(float) * (size_t)cld * mb))) return rc;
  // inside an evaluation sweep (rec_eval_hold_params) the weight image and the bias operand blocks are packed once
  const bool reuse = e->k_hold && e->k_img_epoch == e->param_epoch && e->k_img_net == a.net_id && e->k_img_head == a.stats_head &&
                     e->k_sup_net == a.net_id && e->k_sup_head == a.stats_head;
  if (!reuse && (rc = tck_pack_head_image(e, a.net_id, a.stats_head, 0, a.w))) return rc;
  tck::PackSrc hs = {};
  hs.n = 1; hs.p[0] = a.h; hs.w[0] = 1.f;
  if ((rc = tck_pack(e, hs, a.B, e->D, e->k_himg[0]))) return rc;
  tck::FwdParams p = {};
  p.himg = e->k_himg[0]; p.wimg = e->k_wimg[0]; p.KB = KB; p.B = a.B; p.Vloc = e->Vloc; p.vocab_lo = e->cfg.vocab_lo; p.n_tiles = n_tiles;
  p.bias = np.head_b[a.stats_head]; p.target = a.target; p.part = e->part; p.part_stride = e->part_stride;
  p.topk = a.topk; p.cmax = e->k_cmax; p.cmax_ld = cld; p.n_sb = n_sb;
  const int64_t rec_cap = (int64_t)e->sm_count * 4 * 128 + 4 * (int64_t)e->cfg.max_batch;
  const bool pair = KB == 1 && !getenv("REC_NO_CMAX_PAIR");
  int n_split;
  if ((rc = tck_alloc(e, (void **)&e->k_chosen, sizeof(int) * (size_t)mb * tck::KC))) return rc;
  if (pair) {
    // grid = (vocabulary splits, session-block pairs); a split is a multiple of 8 tiles (level-2 groups are CTA-local)
    constexpr int GROUP = tck::HeadCmaxPair::GROUP;
    const int n_pairs = cdiv(n_sb, 2);
    int splits = e->sm_count / n_pairs;
    if (splits < 1) splits = 1;
    while ((int64_t)splits * a.B > rec_cap && splits > 1) --splits;
    const int per = cdiv(cdiv(n_tiles, splits), GROUP) * GROUP;
    n_split = cdiv(n_tiles, per);
    const int64_t cld2 = (2 * cdiv(n_tiles, GROUP) + 31) / 32 * 32;
    if ((rc = tck_alloc(e, (void **)&e->k_cmax2, sizeof(float) * (size_t)cld2 * mb))) return rc;
    if ((rc = tck_alloc(e, (void **)&e->k_bblk, (size_t)(n_tiles + 1) * 4096))) return rc;
    if (!reuse) {
      tck::pack_bias_blocks_kernel<<<n_tiles + 1, 128, 0, e->stream>>>(np.head_b[a.stats_head], e->Vloc, n_tiles, e->k_bblk);
      REC_LAUNCH_CHECK(e);
    }
    p.per = per; p.cmax2 = e->k_cmax2; p.cmax2_ld = cld2; p.bblk = e->k_bblk;
    if ((int64_t)n_split * a.B > rec_cap)
      REC_FAIL(e, REC_EINVAL, "chunk-maxima pass: %d records per row exceed the statistics workspace", n_split);
    if (e->timing) cudaEventRecord(e->ev[2], e->stream);  // slot 1 (evaluation head kernel): this ONE launch
    static const int pv = getenv("REC_CMAX_EXP") ? atoi(getenv("REC_CMAX_EXP")) : 0;
    static const char *trace_path = getenv("REC_CMAX_TRACE");  // debugging: clock64 phase stamps of one CTA -> file
    static long long *d_trace = nullptr;
    if (trace_path && !d_trace) cudaMalloc(&d_trace, sizeof(long long) * 8 * 64);
    if (trace_path) { cudaMemsetAsync(d_trace, 0, sizeof(long long) * 8 * 64, e->stream); p.trace = d_trace; }
    const dim3 grid(n_split, n_pairs);
    switch (pv) {
      case 4: rc = tck::launch_tck<tck::HeadCmaxPairT<4>>(e, grid, p); break;
      case 10: rc = tck::launch_tck<tck::HeadCmaxPairT<10>>(e, grid, p); break;
      case 11: rc = tck::launch_tck<tck::HeadCmaxPairT<11>>(e, grid, p); break;
      case 30: rc = tck::launch_tck<tck::HeadCmaxPairT<30>>(e, grid, p); break;
      default: rc = tck::launch_tck<tck::HeadCmaxPair>(e, grid, p); break;
    }
    if (rc) return rc;
    if (e->timing) cudaEventRecord(e->ev[3], e->stream);
    if (trace_path) {
      long long host[8 * 64];
      cudaStreamSynchronize(e->stream);
      cudaMemcpy(host, d_trace, sizeof(host), cudaMemcpyDeviceToHost);
      if (FILE *f = fopen(trace_path, "wb")) { fwrite(host, sizeof(host), 1, f); fclose(f); }
    }
    tck::chunk_select2_kernel<<<cdiv(a.B, 8), 256, 0, e->stream>>>(e->k_cmax, cld, n_chunks, e->k_cmax2, cld2, n_tiles, a.B, e->k_chosen);
    REC_LAUNCH_CHECK(e);
    tck::chunk_score64_kernel<<<cdiv(a.B, 4), 128, 0, e->stream>>>(e->k_chosen, np.head_w[a.stats_head], np.head_b[a.stats_head], a.h, a.B,
                                                                  e->Vloc, e->cfg.vocab_lo, a.topk, e->row_ids, e->row_topv, summary,
                                                                  e->part_stride);
  } else {
    // flattened (session block, tile) unit space: one persistent CTA per SM, equal shares
    const int total = n_sb * n_tiles;
    int n_cta = e->sm_count < total ? e->sm_count : total;
    const int per = cdiv(total, n_cta);
    n_cta = cdiv(total, per);
    n_split = cdiv(n_tiles, per) + 1;  // records per row: the CTAs that can touch one session block
    if (n_split > n_cta) n_split = n_cta;
    p.per = per;
    if ((int64_t)n_split * a.B > rec_cap)
      REC_FAIL(e, REC_EINVAL, "chunk-maxima pass: %d records per row exceed the statistics workspace", n_split);
    tck::fill_neutral_records_kernel<<<(int)cdiv64((int64_t)n_split * a.B, 256), 256, 0, e->stream>>>(e->part, (int64_t)n_split * a.B, e->part_stride);
    REC_LAUNCH_CHECK(e);
    if (e->timing) cudaEventRecord(e->ev[2], e->stream);  // slot 1 (evaluation head kernel): this ONE launch
    if ((rc = tck::launch_tck<tck::HeadCmaxFlat>(e, dim3(n_cta), p))) return rc;
    if (e->timing) cudaEventRecord(e->ev[3], e->stream);
    tck::chunk_select_kernel<<<cdiv(a.B, 8), 256, 0, e->stream>>>(e->k_cmax, cld, n_chunks, a.B, e->k_chosen);
    REC_LAUNCH_CHECK(e);
    const size_t smem = 8 * (size_t)e->D * sizeof(float);
    tck::chunk_score_kernel<0><<<cdiv(a.B, 8), 256, smem, e->stream>>>(e->k_chosen, np.head_w[a.stats_head], np.head_b[a.stats_head], a.h, a.B,
                                                                      e->D, e->Vloc, e->cfg.vocab_lo, a.topk, e->row_ids, e->row_topv,
                                                                      summary, e->part_stride);
  }
  REC_LAUNCH_CHECK(e);
  e->k_img_epoch = e->param_epoch; e->k_img_net = a.net_id; e->k_img_head = a.stats_head;
  *n_split_out = n_split;
  e->st_approx = false;  // the ids / scores in row_ids / row_topv are exact
  e->st_kpub = 0;
  e->st_apub = 0;
  return REC_OK;
}

// Same contract as launch_head_stats (heads.cu) for statistics (no top-k) and the greedy-action pass.
int launch_head_stats_tck(rec_engine *e, const HeadStatsArgs &a, int *n_split_out) {
  int rc = tck_ensure(e, a.n_arg > 0 ? 2 : 1);
  if (rc) return rc;
  if (a.topk > tck::HeadTopk<true>::KMAX || (a.topk > 0 && a.n_arg > 0))
    REC_FAIL(e, REC_EINVAL, "tensor-core statistics over operand images: top-k <= %d, not combined with a greedy-action pass", tck::HeadTopk<true>::KMAX);
  const rec_net_params &np = e->nets[a.net_id].p;
  const int KB = e->D / 64, n_tiles = cdiv(e->Vloc, 128), n_sb = cdiv(a.B, 128);
  const bool arg = a.n_arg > 0;
  if (arg && a.n_arg > 3) REC_FAIL(e, REC_EINVAL, "greedy-action pass supports at most 3 Q heads (got %d)", a.n_arg);
  // weight image of the scored head(s) -- unless tck_prepack_heads() already produced it on a side stream
  uint8_t *wimg = e->k_wimg[arg ? 1 : 0];
  const float *bias = arg ? (a.n_arg > 1 ? e->k_bias : np.head_b[1 + a.arg_shift]) : np.head_b[a.stats_head];
  const bool fresh = arg ? (e->k_fresh[1] && a.arg_shift == 0) : (e->k_fresh[0] && e->k_sup_net == a.net_id && e->k_sup_head == a.stats_head);
  e->k_fresh[arg ? 1 : 0] = false;
  if (!fresh && (rc = tck_pack_head_image(e, a.net_id, arg ? -1 : a.stats_head, a.n_arg, a.w, a.arg_shift))) return rc;
  // state image
  tck::PackSrc hs = {};
  hs.n = 1; hs.p[0] = a.h; hs.w[0] = 1.f;
  uint8_t *himg = e->k_himg[arg ? 1 : 0];
  if ((rc = tck_pack(e, hs, a.B, e->D, himg))) return rc;

  int n_split = e->sm_count / n_sb;
  if (n_split > n_tiles) n_split = n_tiles;
  if (n_split < 1) n_split = 1;
  const int per = cdiv(n_tiles, n_split);
  n_split = cdiv(n_tiles, per);
  tck::FwdParams p = {};
  p.himg = himg; p.wimg = wimg; p.KB = KB; p.B = a.B; p.Vloc = e->Vloc; p.vocab_lo = e->cfg.vocab_lo; p.n_tiles = n_tiles;
  p.bias = bias; p.target = arg ? nullptr : a.target; p.part = e->part; p.part_stride = e->part_stride;
  dim3 grid(n_split, n_sb);
  // kernel-timing mode: the caller's start event (slot 4 statistics / slot 5 greedy action) is re-recorded here so that
  // it brackets exactly this ONE launch, not the packing kernels above
  if (e->timing) cudaEventRecord(e->ev[arg ? 10 : 8], e->stream);
  p.topk = a.topk;
  if (a.topk > 0) rc = KB == 1 ? tck::launch_tck<tck::HeadTopk<true>>(e, grid, p) : tck::launch_tck<tck::HeadTopk<false>>(e, grid, p);
  else rc = arg ? tck::launch_tck<tck::HeadFwd<tck::M_ARG>>(e, grid, p) : tck::launch_tck<tck::HeadFwd<tck::M_STATS>>(e, grid, p);
  if (rc) return rc;
  *n_split_out = n_split;
  e->st_approx = true;
  e->st_kpub = a.topk > 0 ? rec_kpub(a.topk, tck::HeadTopk<true>::CS) : 0;
  e->st_apub = arg ? 2 : 0;
  return REC_OK;
}

int tck_bwd_slices(const rec_engine *e) {
  const int n_tiles = cdiv(e->Vloc, 128), KB = e->D / 64;
  const int NCB = KB % 4 == 0 ? 4 : 2, n_dchunks = KB / NCB;
  const int n_sb_max = cdiv(e->cfg.max_batch, 128);
  int n_split = e->sm_count / (n_dchunks * (n_sb_max < 2 ? n_sb_max : 2));
  if (n_split > n_tiles) n_split = n_tiles;
  if (n_split > e->n_dh_part - 1) n_split = e->n_dh_part - 1;
  if (n_split < 1) n_split = 1;
  const int per = cdiv(n_tiles, n_split);
  return cdiv(n_tiles, per);
}

// Supervised head backward + Adam for D >= 128: dlogits image -> (dh partial slices || dW + Adam).
// Writes tck_bwd_slices(e) slices of dh_part (every slice fully).
int launch_head_bwd_adam_tck(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B, float step_size,
                             float bc2_sqrt, const rec_train_hparams *hp, float inv_B) {
  int rc = tck_ensure(e, 1 | 4);
  if (rc) return rc;
  const rec_net_params &np = e->nets[net_id].p;
  const int KB = e->D / 64, n_tiles = cdiv(e->Vloc, 128), n_sb = cdiv(B, 128);
  uint8_t *wimg = e->k_wimg[0];
  if (e->k_sup_net != net_id || e->k_sup_head != 0) {  // the statistics pass of this step packed another head
    tck::PackSrc ws = {};
    ws.n = 1; ws.p[0] = np.head_w[0]; ws.w[0] = 1.f;
    if ((rc = tck_pack(e, ws, e->Vloc, e->D, wimg))) return rc;
    tck::PackSrc hs = {};
    hs.n = 1; hs.p[0] = h; hs.w[0] = 1.f;
    if ((rc = tck_pack(e, hs, B, e->D, e->k_himg[0]))) return rc;
  }
  e->k_sup_net = -1;
  const int KBS = 2 * n_sb;
  tck::pack_img_T_kernel<<<cdiv(e->D * KBS * 8, 256), 256, 0, e->stream>>>(h, B, e->D, KBS, e->k_hT);
  REC_LAUNCH_CHECK(e);
  // (1) dlogits image + bias-gradient partials
  {
    int n_split = e->sm_count / n_sb;
    if (n_split > n_tiles) n_split = n_tiles;
    if (n_split < 1) n_split = 1;
    const int per = cdiv(n_tiles, n_split);
    n_split = cdiv(n_tiles, per);
    tck::FwdParams p = {};
    p.himg = e->k_himg[0]; p.wimg = wimg; p.KB = KB; p.B = B; p.Vloc = e->Vloc; p.vocab_lo = e->cfg.vocab_lo; p.n_tiles = n_tiles;
    p.bias = np.head_b[0]; p.target = b->a; p.row_stats = e->row_stats; p.inv_B = inv_B;
    p.dlT = e->k_dlT; p.dl_cb = KBS; p.db_part = e->k_db; p.extra = e->bwd_extra;
    if ((rc = tck::launch_tck<tck::HeadFwd<tck::M_DL>>(e, dim3(n_split, n_sb), p))) return rc;
  }
  // (2) dh partial slices (reads the weight IMAGE: independent of the Adam update below)
  {
    const int NCB = KB % 4 == 0 ? 4 : 2, n_dchunks = KB / NCB;
    int n_split = e->sm_count / (n_dchunks * n_sb);
    if (n_split > n_tiles) n_split = n_tiles;
    const int n_slices = tck_bwd_slices(e);
    if (n_split > n_slices) n_split = n_slices;
    if (n_split < 1) n_split = 1;
    const int per = cdiv(n_tiles, n_split);
    n_split = cdiv(n_tiles, per);
    tck::DhParams p = {};
    p.dlT = e->k_dlT; p.wimg = wimg; p.KB = KB; p.NCB = NCB; p.dl_cb = KBS; p.n_tiles = n_tiles; p.B = B; p.D = e->D;
    p.dh_part = e->dh_part;
    if (n_split < n_slices)  // slices the split does not reach must still read as zero
      REC_CUDA(e, cudaMemsetAsync(e->dh_part + (int64_t)n_split * B * e->D, 0, sizeof(float) * (size_t)(n_slices - n_split) * B * e->D, e->stream));
    if ((rc = tck::launch_tck<tck::HeadDh>(e, dim3(n_split, n_dchunks, n_sb), p))) return rc;
  }
  side_mark(e, 3);  // the dh slices are final: the GRU backward need not wait for the (HBM-bound) weight update below
  // (3) dW + Adam
  {
    tck::DwParams p = {};
    p.dlT = e->k_dlT; p.hT = e->k_hT; p.KBS = KBS; p.n_vt = cdiv(n_tiles, 2); p.n_dt = e->D / 128; p.n_tiles = n_tiles; p.Vloc = e->Vloc;
    p.D = e->D; p.n_sb = n_sb;
    p.w = np.head_w[0]; p.wm = np.head_w_m[0]; p.wv = np.head_w_v[0]; p.b = np.head_b[0]; p.bm = np.head_b_m[0]; p.bv = np.head_b_v[0];
    p.db_part = e->k_db; p.b1 = hp->beta1; p.b2 = hp->beta2; p.eps = hp->eps; p.step_size = step_size; p.inv_bc2_sqrt = 1.f / bc2_sqrt;
    p.sc = e->d_sc;
    const int total = p.n_vt * p.n_dt;
    // with branch overlap on, this kernel runs NEXT TO the GRU backward (32 latency-bound CTAs in clusters of 8 that need
    // a whole SM's shared memory each): leave SMs free for them instead of queueing behind one full persistent wave
    // (measured at cfg3: reserve 0 / 32 / 40 / 56 SMs -> 4.23 / 4.24 / 7.47 (a cluster starved) / 3.98 ms per step)
    static const int reserve = getenv("REC_DW_RESERVE") ? atoi(getenv("REC_DW_RESERVE")) : 56;
    int n_cta = side_enabled(e) ? e->sm_count - reserve : e->sm_count;
    if (n_cta > total) n_cta = total;
    if (n_cta < 1) n_cta = 1;
    if (e->timing) cudaEventRecord(e->ev[0], e->stream);  // slot 0 brackets this ONE launch (the caller records the end)
    if ((rc = tck::launch_tck<tck::HeadDwAdam>(e, dim3(n_cta), p))) return rc;
  }
  return REC_OK;
}
