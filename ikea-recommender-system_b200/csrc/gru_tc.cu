// GRU trunk on the tensor cores for wide layers (E, H multiples of 128: BASELINE cfg3 has E = H = 256, L = 50,
// bidirectional).  The CUDA-core kernels of gru.cu stream W_ih / W_hh from L2 for every group of 4 sessions and time
// step (cfg3: 5.4 ms for the three forward passes, 3.6 ms BPTT, 0.9 ms weight gradients); here every product is a
// tcgen05 GEMM over packed bf16 hi/lo operand images (tck.cuh; bf16x3 = fp32-class accuracy):
//
//   forward   X image (embedding rows gathered + packed once per input)                       gather_pack_kernel
//             gi[B L, 3H] = X . W_ih^T + b_ih  for ALL time steps: one GEMM per (pass, direction)      Gemm (K x K)
//             per time step ONE launch for all passes / directions / session blocks:
//               gh = h_{t-1} . W_hh^T (A = the bf16 image of h_{t-1}, B = W_hh with the r|z|n rows of 64 hidden units
//               grouped, N = 192) -> epilogue: gates, h_t (fp32 state + next step's bf16 image), saved activations     GruStep
//   backward  per time step ONE launch: dh_{t-1} = direct + dgh_t . W_hh (A = step image of dgh_t, B = MN-major view of
//             the W_hh image) -> epilogue: gate gradients of step t-1 -> step image + the [B L, 3H] images           GruBptt
//             dx = dgi . W_ih, dW_ih = dgi^T . X, dW_hh = dgh^T . Hprev (split-K partials), bias = column sums        Gemm
// The time loop is a chain of small launches captured in the step's CUDA graph: no grid-wide barrier, no cluster.
// Semantics: torch.nn.GRU + pack_padded_sequence exactly as gru.cu (models/BidirGRU4Rec/model.py:51-99,
// models/SQN/sqn_gru.py:69-104): per-row length mask, reverse direction walks len-1..0, layer 0 only.
#include <stdlib.h>
#include "tck.cuh"

namespace gtc {
using tck::BLK;
using tck::BLK2;
using tck::HALF;
using tck::EPI_THREADS;

#define GTC_STAMP(p, u, k) do { if ((p).trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) (p).trace[(u) * 8 + (k)] = clock64(); } while (0)
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ int eff_len(const int64_t *lens, int b, int L, int packed) {
  if (!packed) return L;
  const int64_t l = lens[b];
  return (int)(l < 1 ? 1 : (l > L ? L : l));
}

// ---- packing --------------------------------------------------------------------------------------------------
// X image: row p = b * L + t holds emb[s[b, t]] (rows beyond B * L are zeros).
__global__ void __launch_bounds__(256) gather_pack_kernel(const float *__restrict__ emb, const int64_t *__restrict__ s, int P,
                                                         int E, int N, uint8_t *__restrict__ img) {
  const int c8n = E >> 3, KB = E >> 6;
  const int64_t n_chunks = (int64_t)((P + 127) / 128) * 128 * c8n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_chunks; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / c8n;
    const int c8 = (int)(i - row * c8n);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (row < P) {
      int64_t it = s[row];
      it = it < 0 ? 0 : (it > N ? N : it);
      const float4 *p0 = reinterpret_cast<const float4 *>(emb + it * E + c8 * 8);
      a = __ldg(p0); b = __ldg(p0 + 1);
    }
    uint8_t *blk = img + ((row >> 7) * KB + (c8 >> 3)) * (int64_t)BLK2;
    tc::store_split8(blk, blk + BLK, (int)(row & 127), c8 & 7, a, b);
  }
}

// GRU weights of up to 2 nets x 2 directions -> images.  which 0: W_ih [3H, E] natural; 1: W_hh^T [H, 3H] (rows =
// state column, K = gate row: the B operand of the BPTT step, any 32-row slice of it is contiguous per k-block);
// 2: W_hh with the rows regrouped per 32 hidden units: unit j = rows {r_j (32) | z_j (32) | n_j (32)}, stored as
// [H/32 units][H/64 k-blocks][hi 96 x 128 B | lo 96 x 128 B] (the B operand of the forward step).
struct WPackArgs {
  const float *wih[4], *whh[4];
  uint8_t *wih_img[4], *whh_img[4], *whh_perm[4];
  int n_slots, E, H;
};
__global__ void __launch_bounds__(256) gru_pack_weights_kernel(WPackArgs a) {
  const int slot = blockIdx.z, which = blockIdx.y;
  const int G = 3 * a.H;
  if (which == 1) {  // transposed: image row c (state column), image column j (gate row)
    const int c8n = G >> 3, KB = G >> 6, n_chunks = a.H * c8n;
    const float *src = a.whh[slot];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_chunks; i += gridDim.x * blockDim.x) {
      const int c8 = i / a.H, row = i - c8 * a.H;  // consecutive threads: consecutive c -> coalesced reads of W_hh rows
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = src[(int64_t)(c8 * 8 + k) * a.H + row];
      uint8_t *blk = a.whh_img[slot] + ((int64_t)(row >> 7) * KB + (c8 >> 3)) * BLK2;
      tc::store_split8(blk, blk + BLK, row & 127, c8 & 7, make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]));
    }
    return;
  }
  const int C = which == 0 ? a.E : a.H, c8n = C >> 3, KB = C >> 6;
  const float *src = which == 0 ? a.wih[slot] : a.whh[slot];
  const int n_chunks = G * c8n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_chunks; i += gridDim.x * blockDim.x) {
    const int row = i / c8n, c8 = i - row * c8n;
    const float4 *p0 = reinterpret_cast<const float4 *>(src + (int64_t)row * C + c8 * 8);
    const float4 x = p0[0], y = p0[1];
    if (which == 0) {
      uint8_t *blk = a.wih_img[slot] + ((int64_t)(row >> 7) * KB + (c8 >> 3)) * BLK2;
      tc::store_split8(blk, blk + BLK, row & 127, c8 & 7, x, y);
    } else {
      const int gt = row / a.H, u = row - gt * a.H, j = u >> 5, r = gt * 32 + (u & 31);
      uint8_t *blk = a.whh_perm[slot] + ((int64_t)j * KB + (c8 >> 3)) * (2 * 96 * 128);
      tc::store_split8(blk, blk + 96 * 128, r, c8 & 7, x, y);
    }
  }
}

// ---- generic GEMM over images: C[z][M, N] (+ bias) = A . B^T, fp32 output, optional split-K slices --------------
struct GemmParams {
  const uint8_t *A[6], *B[6];
  float *C[6];
  const float *bias[6];
  int a_mn, b_mn;          // 0: K-major view (MN = image rows), 1: MN-major view (MN = image columns)
  int a_cbs, b_cbs;        // column blocks per row block (image pitch)
  int m_tiles, n_tiles;    // tiles of 128 rows / NT columns
  int NT;                  // 128 or 256
  int k_total, n_split;    // K / 64; split-K
  int M;                   // C rows >= M are not written
  int c_trans;             // 1: the tile is computed transposed -- accumulator row m, column n is C[n * ldc + m] (M <-> N
                           //    swapped so that a warp's 32 lanes, = 32 consecutive m, store 128 contiguous bytes)
  int64_t ldc, c_split_stride;
};

struct Gemm {
  using Params = GemmParams;
  static constexpr bool CLUSTERED = false;
  static constexpr int RESIDENT_BYTES = 0;
  static constexpr int EPI_WARPS = 8;
  static constexpr const char *NAME = "tck:gemm";
  static constexpr int STAGES = 2, STAGE_BYTES = 3 * BLK2, ACC_COLS = 256, TMEM_COLS = 512;
  static constexpr int EXTRA_BYTES = 0;
  __device__ static __forceinline__ void units(const Params &p, int &lo, int &hi) {
    const int total = p.m_tiles * p.n_tiles * p.n_split;
    const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
    lo = blockIdx.x * per;
    hi = min(total, lo + per);
    if (hi < lo) hi = lo;
  }
  __device__ static __forceinline__ void decode(const Params &p, int u, int &mt, int &nt, int &k_lo, int &k_hi, int &sp) {
    sp = u % p.n_split;
    const int r = u / p.n_split;
    nt = r % p.n_tiles;
    mt = r / p.n_tiles;
    const int per = (p.k_total + p.n_split - 1) / p.n_split;
    k_lo = sp * per;
    k_hi = min(p.k_total, k_lo + per);
    if (k_hi < k_lo) k_hi = k_lo;
  }
  __device__ static __forceinline__ int k_steps(const Params &p, int u) {
    int mt, nt, k_lo, k_hi, sp;
    decode(p, u, mt, nt, k_lo, k_hi, sp);
    return k_hi - k_lo;
  }
  __device__ static __forceinline__ void load(const Params &p, int u, int ks, uint8_t *stage, uint64_t *bar) {
    int mt, nt, k_lo, k_hi, sp;
    decode(p, u, mt, nt, k_lo, k_hi, sp);
    const int kk = k_lo + ks, rb = kk >> 1, rh = kk & 1, z = blockIdx.z;
    tc::mbar_expect_tx(bar, (uint32_t)(BLK2 + p.NT * 256));
    if (!p.a_mn) {
      tc::bulk_g2s(stage, p.A[z] + ((int64_t)mt * p.a_cbs + kk) * BLK2, BLK2, bar);
    } else {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint8_t *blk = p.A[z] + ((int64_t)rb * p.a_cbs + 2 * mt + c) * BLK2 + rh * HALF;
        tc::bulk_g2s(stage + c * HALF, blk, HALF, bar);
        tc::bulk_g2s(stage + 2 * HALF + c * HALF, blk + BLK, HALF, bar);
      }
    }
    uint8_t *bst = stage + BLK2;
    if (!p.b_mn) {
      const int nrb = p.NT >> 7;
      for (int i = 0; i < nrb; ++i) {
        const uint8_t *blk = p.B[z] + ((int64_t)(nt * nrb + i) * p.b_cbs + kk) * BLK2;
        tc::bulk_g2s(bst + i * BLK, blk, BLK, bar);
        tc::bulk_g2s(bst + nrb * BLK + i * BLK, blk + BLK, BLK, bar);
      }
    } else {
      const int ncb = p.NT >> 6;
      for (int c = 0; c < ncb; ++c) {
        const uint8_t *blk = p.B[z] + ((int64_t)rb * p.b_cbs + nt * ncb + c) * BLK2 + rh * HALF;
        tc::bulk_g2s(bst + c * HALF, blk, HALF, bar);
        tc::bulk_g2s(bst + ncb * HALF + c * HALF, blk + BLK, HALF, bar);
      }
    }
  }
  __device__ static __forceinline__ void mma(const Params &p, int, int, uint32_t st, uint32_t tacc, bool first) {
    const uint32_t id = tc::instr_desc(128, p.NT, p.a_mn, p.b_mn);
    const uint64_t ah = p.a_mn ? tc::desc_mnmajor(st, 0, HALF) : tc::desc_kmajor(st, 0);
    const uint64_t al = p.a_mn ? tc::desc_mnmajor(st + 2 * HALF, 0, HALF) : tc::desc_kmajor(st + BLK, 0);
    const uint32_t bs = st + BLK2;
    const uint64_t bh = p.b_mn ? tc::desc_mnmajor(bs, 0, HALF) : tc::desc_kmajor(bs, 0);
    const uint64_t bl = p.b_mn ? tc::desc_mnmajor(bs + (p.NT >> 6) * HALF, 0, HALF) : tc::desc_kmajor(bs + (p.NT >> 7) * BLK, 0);
    const uint64_t sa = p.a_mn ? 128 : 2, sb = p.b_mn ? 128 : 2;
    bool acc = !first;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
      const uint64_t a = pass == 2 ? al : ah, b = pass == 1 ? bl : bh;
#pragma unroll
      for (int k = 0; k < 4; ++k) { tc::mma_bf16(tacc, a + (uint64_t)k * sa, b + (uint64_t)k * sb, id, acc); acc = true; }
    }
  }
  struct Epi {
    int q, cq, lane;
    __device__ __forceinline__ Epi(const Params &, uint8_t *, int tid) {
      const int warp = tid >> 5;
      lane = tid & 31; q = warp & 3; cq = warp >> 2;
    }
    __device__ __forceinline__ void tile(const Params &p, int u, int, uint32_t tacc) {
      int mt, nt, k_lo, k_hi, sp;
      decode(p, u, mt, nt, k_lo, k_hi, sp);
      const int z = blockIdx.z;
      const int row = mt * 128 + q * 32 + lane;
      const int half = p.NT >> 1;
      if (p.c_trans) {
        // accumulator row = C column: lanes store consecutive floats of one C row
        float *dst = p.C[z] + (int64_t)sp * p.c_split_stride + row;
        const float bv = p.bias[z] ? __ldg(p.bias[z] + row) : 0.f;
        for (int ch = 0; ch < half / 32; ++ch) {
          const int c0 = nt * p.NT + cq * half + ch * 32;
          if (c0 >= p.M) break;  // warp-uniform
          float g[32];
          if (k_hi > k_lo) {
            tc::tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * half + ch * 32), g);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) g[j] = 0.f;
          }
          const int nv = min(32, p.M - c0);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nv) dst[(int64_t)(c0 + j) * p.ldc] = g[j] + bv;
        }
        return;
      }
      float *dst = p.C[z] + (int64_t)sp * p.c_split_stride + (int64_t)row * p.ldc + nt * p.NT;
      const float *bias = p.bias[z] ? p.bias[z] + nt * p.NT : nullptr;
      const bool vec = (p.ldc & 3) == 0 && (p.c_split_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C[z]) & 15) == 0;
      for (int ch = 0; ch < half / 32; ++ch) {
        const int c0 = cq * half + ch * 32;
        float g[32];
        if (k_hi > k_lo) {
          tc::tmem_ld32(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, g);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) g[j] = 0.f;
        }
        if (row < p.M) {
          if (bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) g[j] += __ldg(bias + c0 + j);
          }
          if (vec) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(dst + c0 + j) = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[c0 + j] = g[j];
          }
        }
      }
    }
    __device__ __forceinline__ void finish(const Params &) {}
  };
};

// ---- one forward time step -----------------------------------------------------------------------------------------
struct StepPass {
  const uint8_t *whh_perm;
  const float *gi, *b_hh;
  const int64_t *lens;
  float *h_state;
  int save;
};
struct StepParams {
  StepPass pass[6];  // z = pass * dirs + dir
  uint8_t *himg[2];  // ping/pong images of h_t: step t reads [t & 1], writes [(t + 1) & 1]; each [z][n_sb][KBh]
  float *gates_save, *hprev_save;
  uint8_t *hprev_img;  // [dir][Prb][KBh]
  int KBh, B, L, H, dirs, packed, n_sb, Prb;
  long long *trace;  // debug (REC_TRACE_SEL=3): clock64 stamps of CTA (0,0,0): [step][8] = loader in/out, issuer in/out, epilogue in/out
};

struct GruStep {
  using Params = StepParams;
  // one launch = all L time steps: unit u = step u.  The H/32 CTAs of a cluster own 32 hidden units each of the same
  // (pass, direction, session block); their slice of W_hh (r|z|n rows of the 32 units, all k-blocks) stays RESIDENT in
  // shared memory, only the image of h_{t-1} streams in (all four k-blocks in flight at once), and h_t travels to the
  // other CTAs through the ping/pong image in global memory + the cluster-wide step barrier.
  static constexpr bool CLUSTERED = true;
  static constexpr int EPI_WARPS = 8;
  static constexpr const char *NAME = "tck:gru_steps";
  static constexpr int WB = 2 * 96 * 128;  // one (unit, k-block) of the regrouped W_hh image: hi 12 KB | lo 12 KB
  static constexpr int STAGES = 4, STAGE_BYTES = BLK2, ACC_COLS = 128, TMEM_COLS = 256;
  static constexpr int RESIDENT_BYTES = 4 * WB;  // H <= 256
  static constexpr int EXTRA_BYTES = 0;
  __device__ static __forceinline__ void units(const Params &p, int &lo, int &hi) { lo = 0; hi = p.L; }
  __device__ static __forceinline__ int k_steps(const Params &p, int) { return p.KBh; }
  __device__ static __forceinline__ void load_resident(const Params &p, uint8_t *res, uint64_t *bar) {
    tc::mbar_expect_tx(bar, (uint32_t)p.KBh * WB);
    tc::bulk_g2s(res, p.pass[blockIdx.z].whh_perm + (int64_t)blockIdx.x * p.KBh * WB, (uint32_t)p.KBh * WB, bar);
  }
  __device__ static __forceinline__ void load(const Params &p, int u, int ks, uint8_t *stage, uint64_t *bar) {
    if (ks == 0) GTC_STAMP(p, u, 0);
    tc::mbar_expect_tx(bar, BLK2);
    tc::bulk_g2s(stage, p.himg[u & 1] + (((int64_t)blockIdx.z * p.n_sb + blockIdx.y) * p.KBh + ks) * BLK2, BLK2, bar);
    if (ks == p.KBh - 1) GTC_STAMP(p, u, 1);
  }
  __device__ static __forceinline__ void mma(const Params &p, int u, int ks, uint32_t st, uint32_t res, uint32_t tacc, bool first) {
    if (ks == 0) GTC_STAMP(p, u, 2);
    if (ks == p.KBh - 1) GTC_STAMP(p, u, 3);
    const uint32_t id = tc::instr_desc(128, 96, 0, 0);
    const uint64_t ah = tc::desc_kmajor(st, 0), al = tc::desc_kmajor(st + BLK, 0);
    const uint64_t bh = tc::desc_kmajor(res + ks * WB, 0), bl = tc::desc_kmajor(res + ks * WB + 96 * 128, 0);
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
    __device__ __forceinline__ Epi(const Params &, uint8_t *, int tid) {
      const int warp = tid >> 5;
      lane = tid & 31; q = warp & 3; cq = warp >> 2;
    }
    __device__ __forceinline__ void tile(const Params &p, int step, int, uint32_t tacc) {
      if (threadIdx.x == 0) GTC_STAMP(p, step, 4);
      const int z = blockIdx.z, j = blockIdx.x, sb = blockIdx.y;
      const StepPass &P = p.pass[z];
      const int dir = z % p.dirs, H = p.H, G = 3 * H;
      const int rl = q * 32 + lane, row = sb * 128 + rl;
      const bool valid = row < p.B;
      const int len = valid ? eff_len(P.lens, row, p.L, p.packed) : 0;
      const bool active = valid && step < len;
      const int tok = active ? (dir ? len - 1 - step : step) : 0;
      const int u0 = j * 32 + cq * 16;
      const int64_t pos = (int64_t)row * p.L + tok;
      const float *gi = P.gi + (pos * p.dirs + dir) * G + u0;
      float *hs = P.h_state + (int64_t)row * (p.dirs * H) + dir * H + u0;
      float *gs = p.gates_save + (pos * p.dirs + dir) * 4 * H + u0;
      float *hp = p.hprev_save + (pos * p.dirs + dir) * H + u0;
      uint8_t *oblk = p.himg[(step + 1) & 1] + (((int64_t)z * p.n_sb + sb) * p.KBh + (j >> 1)) * BLK2;
      uint8_t *pblk = p.hprev_img + (((int64_t)dir * p.Prb + (pos >> 7)) * p.KBh + (j >> 1)) * BLK2;
      const int ch0 = (j & 1) * 4 + cq * 2;  // first 16-byte chunk of this thread's 16 hidden units inside their 64-column block
      const uint32_t tbase = tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 16);
      if (valid && step + 1 < len) {  // the next step's input projections (HBM-resident) -> L2 while this step computes
        const int tok1 = dir ? len - 2 - step : step + 1;
        const float *g1 = P.gi + (((int64_t)row * p.L + tok1) * p.dirs + dir) * G + u0;
        prefetch_l2(g1); prefetch_l2(g1 + H); prefetch_l2(g1 + 2 * H);
      }
#pragma unroll 1
      for (int c8 = 0; c8 < 2; ++c8) {  // 8 hidden units per round
        float ar[8], az[8], an[8];
        float hold[8], hnew[8];
        {
          // 8 columns of each gate.  The three loads and the wait share ONE asm statement: the compiler must not touch
          // the destination registers before tcgen05.wait::ld (it cannot see the asynchrony).
          uint32_t r8[8], z8[8], n8[8];
          __syncwarp();
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%24];\n\t"
              "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%25];\n\t"
              "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%26];\n\t"
              "tcgen05.wait::ld.sync.aligned;"
              : "=r"(r8[0]), "=r"(r8[1]), "=r"(r8[2]), "=r"(r8[3]), "=r"(r8[4]), "=r"(r8[5]), "=r"(r8[6]), "=r"(r8[7]),
                "=r"(z8[0]), "=r"(z8[1]), "=r"(z8[2]), "=r"(z8[3]), "=r"(z8[4]), "=r"(z8[5]), "=r"(z8[6]), "=r"(z8[7]),
                "=r"(n8[0]), "=r"(n8[1]), "=r"(n8[2]), "=r"(n8[3]), "=r"(n8[4]), "=r"(n8[5]), "=r"(n8[6]), "=r"(n8[7])
              : "r"(tbase + (uint32_t)(c8 * 8)), "r"(tbase + (uint32_t)(32 + c8 * 8)), "r"(tbase + (uint32_t)(64 + c8 * 8))
              : "memory");
#pragma unroll
          for (int k = 0; k < 8; ++k) { ar[k] = __uint_as_float(r8[k]); az[k] = __uint_as_float(z8[k]); an[k] = __uint_as_float(n8[k]); }
        }
        const int uo = c8 * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) hold[k] = 0.f;
        if (valid && step > 0) {
          const float4 a = *reinterpret_cast<const float4 *>(hs + uo), b = *reinterpret_cast<const float4 *>(hs + uo + 4);
          hold[0] = a.x; hold[1] = a.y; hold[2] = a.z; hold[3] = a.w; hold[4] = b.x; hold[5] = b.y; hold[6] = b.z; hold[7] = b.w;
        }
        if (active) {
          float gr[8], gz[8], gn[8], rg[8], zg[8], ng[8], phn[8];
#pragma unroll
          for (int h4 = 0; h4 < 2; ++h4) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(gi + uo + 4 * h4));
            const float4 y = __ldg(reinterpret_cast<const float4 *>(gi + H + uo + 4 * h4));
            const float4 w = __ldg(reinterpret_cast<const float4 *>(gi + 2 * H + uo + 4 * h4));
            gr[4 * h4] = x.x; gr[4 * h4 + 1] = x.y; gr[4 * h4 + 2] = x.z; gr[4 * h4 + 3] = x.w;
            gz[4 * h4] = y.x; gz[4 * h4 + 1] = y.y; gz[4 * h4 + 2] = y.z; gz[4 * h4 + 3] = y.w;
            gn[4 * h4] = w.x; gn[4 * h4 + 1] = w.y; gn[4 * h4 + 2] = w.z; gn[4 * h4 + 3] = w.w;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int u = u0 + uo + k;
            rg[k] = sigmoidf_(gr[k] + (ar[k] + __ldg(P.b_hh + u)));
            zg[k] = sigmoidf_(gz[k] + (az[k] + __ldg(P.b_hh + H + u)));
            phn[k] = an[k] + __ldg(P.b_hh + 2 * H + u);
            ng[k] = tanhf(gn[k] + rg[k] * phn[k]);
            hnew[k] = (1.f - zg[k]) * ng[k] + zg[k] * hold[k];
          }
          *reinterpret_cast<float4 *>(hs + uo) = make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
          *reinterpret_cast<float4 *>(hs + uo + 4) = make_float4(hnew[4], hnew[5], hnew[6], hnew[7]);
          if (P.save) {
            *reinterpret_cast<float4 *>(gs + uo) = make_float4(rg[0], rg[1], rg[2], rg[3]);
            *reinterpret_cast<float4 *>(gs + uo + 4) = make_float4(rg[4], rg[5], rg[6], rg[7]);
            *reinterpret_cast<float4 *>(gs + H + uo) = make_float4(zg[0], zg[1], zg[2], zg[3]);
            *reinterpret_cast<float4 *>(gs + H + uo + 4) = make_float4(zg[4], zg[5], zg[6], zg[7]);
            *reinterpret_cast<float4 *>(gs + 2 * H + uo) = make_float4(ng[0], ng[1], ng[2], ng[3]);
            *reinterpret_cast<float4 *>(gs + 2 * H + uo + 4) = make_float4(ng[4], ng[5], ng[6], ng[7]);
            *reinterpret_cast<float4 *>(gs + 3 * H + uo) = make_float4(phn[0], phn[1], phn[2], phn[3]);
            *reinterpret_cast<float4 *>(gs + 3 * H + uo + 4) = make_float4(phn[4], phn[5], phn[6], phn[7]);
            *reinterpret_cast<float4 *>(hp + uo) = make_float4(hold[0], hold[1], hold[2], hold[3]);
            *reinterpret_cast<float4 *>(hp + uo + 4) = make_float4(hold[4], hold[5], hold[6], hold[7]);
            tc::store_split8(pblk, pblk + BLK, (int)(pos & 127), ch0 + c8, make_float4(hold[0], hold[1], hold[2], hold[3]),
                             make_float4(hold[4], hold[5], hold[6], hold[7]));
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) hnew[k] = hold[k];
          if (valid && step == 0) {  // (cannot happen: len >= 1) keep the state defined
            *reinterpret_cast<float4 *>(hs + uo) = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4 *>(hs + uo + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        tc::store_split8(oblk, oblk + BLK, rl, ch0 + c8, make_float4(hnew[0], hnew[1], hnew[2], hnew[3]),
                         make_float4(hnew[4], hnew[5], hnew[6], hnew[7]));
      }
      if (threadIdx.x == 0) GTC_STAMP(p, step, 5);
    }
    __device__ __forceinline__ void finish(const Params &) {}
  };
};

// ---- one BPTT time step --------------------------------------------------------------------------------------------
struct BpttParams {
  const uint8_t *whh_img[2];  // per direction: image of W_hh^T [H/128][3H/64]
  uint8_t *dstep[2];          // ping/pong step images [dir][n_sb][KG]: step t reads [(t + 1) & 1], writes [t & 1]
  float *dhw;                 // [B, dirs * H] running dL/dh ("direct" part between the steps)
  const float *gates_save, *hprev_save;
  const int64_t *lens;
  uint8_t *dgi_img, *dgh_img;  // [dir][Prb][KG]
  int Prb, KG, KBh, B, L, H, dirs, packed, n_sb;
};

struct GruBptt {
  using Params = BpttParams;
  // one launch = all L steps in reverse: unit u = step L - 1 - u.  The H/32 CTAs of a cluster own 32 state columns each;
  // their rows of W_hh^T (all 3H/64 k-blocks) stay RESIDENT in shared memory, the step image of the gate gradients
  // streams in through 4 stages.
  static constexpr bool CLUSTERED = true;
  static constexpr int EPI_WARPS = 8;
  static constexpr const char *NAME = "tck:gru_bptt";
  static constexpr int STAGES = 4, STAGE_BYTES = BLK2, ACC_COLS = 32, TMEM_COLS = 64;
  static constexpr int RESIDENT_BYTES = 12 * 8192;  // 3H/64 k-blocks x (32 rows hi 4 KB | lo 4 KB), H <= 256
  static constexpr int EXTRA_BYTES = 0;
  __device__ static __forceinline__ void units(const Params &p, int &lo, int &hi) { lo = 0; hi = p.L; }
  __device__ static __forceinline__ int k_steps(const Params &p, int) { return p.KG; }
  __device__ static __forceinline__ void load_resident(const Params &p, uint8_t *res, uint64_t *bar) {
    tc::mbar_expect_tx(bar, (uint32_t)p.KG * 8192);
    const int r0 = blockIdx.x * 32;  // first state column of this CTA = row of the W_hh^T image
    const uint8_t *src = p.whh_img[blockIdx.z] + (int64_t)(r0 >> 7) * p.KG * BLK2 + (r0 & 127) * 128;
    for (int kb = 0; kb < p.KG; ++kb) {
      tc::bulk_g2s(res + kb * 8192, src + (int64_t)kb * BLK2, 4096, bar);
      tc::bulk_g2s(res + kb * 8192 + 4096, src + (int64_t)kb * BLK2 + BLK, 4096, bar);
    }
  }
  __device__ static __forceinline__ void load(const Params &p, int u, int ks, uint8_t *stage, uint64_t *bar) {
    const int t = p.L - 1 - u;
    tc::mbar_expect_tx(bar, BLK2);
    tc::bulk_g2s(stage, p.dstep[(t + 1) & 1] + (((int64_t)blockIdx.z * p.n_sb + blockIdx.y) * p.KG + ks) * BLK2, BLK2, bar);
  }
  __device__ static __forceinline__ void mma(const Params &, int, int ks, uint32_t st, uint32_t res, uint32_t tacc, bool first) {
    const uint32_t id = tc::instr_desc(128, 32, 0, 0);
    const uint64_t ah = tc::desc_kmajor(st, 0), al = tc::desc_kmajor(st + BLK, 0);
    const uint64_t bh = tc::desc_kmajor(res + ks * 8192, 0), bl = tc::desc_kmajor(res + ks * 8192 + 4096, 0);
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
    __device__ __forceinline__ Epi(const Params &, uint8_t *, int tid) {
      const int warp = tid >> 5;
      lane = tid & 31; q = warp & 3; cq = warp >> 2;
    }
    __device__ __forceinline__ void tile(const Params &p, int u, int, uint32_t tacc) {
      const int step = p.L - 1 - u;
      const int dir = blockIdx.z, cb = blockIdx.x, sb = blockIdx.y;
      const int H = p.H;
      const int rl = q * 32 + lane, row = sb * 128 + rl;
      const bool valid = row < p.B;
      const int len = valid ? eff_len(p.lens, row, p.L, p.packed) : 0;
      const bool active = valid && step < len;
      const int tok = active ? (dir ? len - 1 - step : step) : 0;
      const int c0 = cb * 32 + cq * 16;
      const int64_t pos = (int64_t)row * p.L + tok;
      float acc[16];
      tc::tmem_ld16(tacc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 16), acc);
      float *dh = p.dhw + (int64_t)row * (p.dirs * H) + dir * H + c0;
      const float *g = p.gates_save + (pos * p.dirs + dir) * 4 * H + c0;
      const float *hpv = p.hprev_save + (pos * p.dirs + dir) * H + c0;
      uint8_t *sblk = p.dstep[step & 1] + (((int64_t)dir * p.n_sb + sb) * p.KG) * BLK2;          // + kb * BLK2
      uint8_t *ib = p.dgi_img + (((int64_t)dir * p.Prb + (pos >> 7)) * p.KG) * BLK2;
      uint8_t *hb = p.dgh_img + (((int64_t)dir * p.Prb + (pos >> 7)) * p.KG) * BLK2;
      const int pr = (int)(pos & 127);
      if (valid && step >= 1 && step - 1 < len) {  // the saved activations of the next (earlier) step -> L2
        const int tok1 = dir ? len - step : step - 1;
        const int64_t pos1 = (int64_t)row * p.L + tok1;
        const float *g1 = p.gates_save + (pos1 * p.dirs + dir) * 4 * H + c0;
        prefetch_l2(g1); prefetch_l2(g1 + H); prefetch_l2(g1 + 2 * H); prefetch_l2(g1 + 3 * H);
        prefetch_l2(p.hprev_save + (pos1 * p.dirs + dir) * H + c0);
      }
#pragma unroll 1
      for (int c8 = 0; c8 < 2; ++c8) {
        const int uo = c8 * 8;
        float dhv[8], a_r[8], a_z[8], a_n[8], h_n[8], direct[8];
        if (valid) {
          const float4 x = *reinterpret_cast<const float4 *>(dh + uo), y = *reinterpret_cast<const float4 *>(dh + uo + 4);
          dhv[0] = x.x; dhv[1] = x.y; dhv[2] = x.z; dhv[3] = x.w; dhv[4] = y.x; dhv[5] = y.y; dhv[6] = y.z; dhv[7] = y.w;
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) dhv[k] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) dhv[k] += acc[uo + k];
        if (active) {
          float rg[8], zg[8], ng[8], phn[8], hp[8];
#pragma unroll
          for (int h4 = 0; h4 < 2; ++h4) {
            const float4 a = *reinterpret_cast<const float4 *>(g + uo + 4 * h4);
            const float4 b = *reinterpret_cast<const float4 *>(g + H + uo + 4 * h4);
            const float4 c = *reinterpret_cast<const float4 *>(g + 2 * H + uo + 4 * h4);
            const float4 d = *reinterpret_cast<const float4 *>(g + 3 * H + uo + 4 * h4);
            const float4 e = *reinterpret_cast<const float4 *>(hpv + uo + 4 * h4);
            rg[4 * h4] = a.x; rg[4 * h4 + 1] = a.y; rg[4 * h4 + 2] = a.z; rg[4 * h4 + 3] = a.w;
            zg[4 * h4] = b.x; zg[4 * h4 + 1] = b.y; zg[4 * h4 + 2] = b.z; zg[4 * h4 + 3] = b.w;
            ng[4 * h4] = c.x; ng[4 * h4 + 1] = c.y; ng[4 * h4 + 2] = c.z; ng[4 * h4 + 3] = c.w;
            phn[4 * h4] = d.x; phn[4 * h4 + 1] = d.y; phn[4 * h4 + 2] = d.z; phn[4 * h4 + 3] = d.w;
            hp[4 * h4] = e.x; hp[4 * h4 + 1] = e.y; hp[4 * h4 + 2] = e.z; hp[4 * h4 + 3] = e.w;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float dn = dhv[k] * (1.f - zg[k]);
            const float dz = dhv[k] * (hp[k] - ng[k]);
            a_n[k] = dn * (1.f - ng[k] * ng[k]);
            const float dr = a_n[k] * phn[k];
            a_r[k] = dr * rg[k] * (1.f - rg[k]);
            a_z[k] = dz * zg[k] * (1.f - zg[k]);
            h_n[k] = a_n[k] * rg[k];
            direct[k] = dhv[k] * zg[k];
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) { a_r[k] = 0.f; a_z[k] = 0.f; a_n[k] = 0.f; h_n[k] = 0.f; direct[k] = dhv[k]; }
        }
        if (valid) {
          *reinterpret_cast<float4 *>(dh + uo) = make_float4(direct[0], direct[1], direct[2], direct[3]);
          *reinterpret_cast<float4 *>(dh + uo + 4) = make_float4(direct[4], direct[5], direct[6], direct[7]);
        }
        const float4 r0 = make_float4(a_r[0], a_r[1], a_r[2], a_r[3]), r1 = make_float4(a_r[4], a_r[5], a_r[6], a_r[7]);
        const float4 z0 = make_float4(a_z[0], a_z[1], a_z[2], a_z[3]), z1 = make_float4(a_z[4], a_z[5], a_z[6], a_z[7]);
        const float4 n0 = make_float4(a_n[0], a_n[1], a_n[2], a_n[3]), n1 = make_float4(a_n[4], a_n[5], a_n[6], a_n[7]);
        const float4 m0 = make_float4(h_n[0], h_n[1], h_n[2], h_n[3]), m1 = make_float4(h_n[4], h_n[5], h_n[6], h_n[7]);
        const int ch = (cb & 1) * 4 + cq * 2 + c8;  // 16-byte chunk inside the 64-column block (cb >> 1) of each gate
        // step image of dgh (r | z | n-through-r blocks): every row is written (zeros for rows that are not active)
        {
          uint8_t *b0 = sblk + (int64_t)(0 * p.KBh + (cb >> 1)) * BLK2, *b1 = sblk + (int64_t)(1 * p.KBh + (cb >> 1)) * BLK2, *b2 = sblk + (int64_t)(2 * p.KBh + (cb >> 1)) * BLK2;
          tc::store_split8(b0, b0 + BLK, rl, ch, r0, r1);
          tc::store_split8(b1, b1 + BLK, rl, ch, z0, z1);
          tc::store_split8(b2, b2 + BLK, rl, ch, m0, m1);
        }
        if (active) {
          uint8_t *i0 = ib + (int64_t)(0 * p.KBh + (cb >> 1)) * BLK2, *i1 = ib + (int64_t)(1 * p.KBh + (cb >> 1)) * BLK2, *i2 = ib + (int64_t)(2 * p.KBh + (cb >> 1)) * BLK2;
          uint8_t *h0 = hb + (int64_t)(0 * p.KBh + (cb >> 1)) * BLK2, *h1 = hb + (int64_t)(1 * p.KBh + (cb >> 1)) * BLK2, *h2 = hb + (int64_t)(2 * p.KBh + (cb >> 1)) * BLK2;
          tc::store_split8(i0, i0 + BLK, pr, ch, r0, r1);
          tc::store_split8(i1, i1 + BLK, pr, ch, z0, z1);
          tc::store_split8(i2, i2 + BLK, pr, ch, n0, n1);
          tc::store_split8(h0, h0 + BLK, pr, ch, r0, r1);
          tc::store_split8(h1, h1 + BLK, pr, ch, z0, z1);
          tc::store_split8(h2, h2 + BLK, pr, ch, m0, m1);
        }
      }
    }
    __device__ __forceinline__ void finish(const Params &) {}
  };
};

// Bias gradients: column sums of the dgi / dgh images over this split's position range -> bias column (KS - 1) of the
// split-K partial layout part[split][dir][which][3H][KS] that gru_adam_kernel reduces.
// grid (splits, 3H/64 column blocks, dirs * 2); a warp reads whole 128-byte image rows (lane = two columns).
__global__ void __launch_bounds__(256) gru_bias_colsum_kernel(const uint8_t *__restrict__ dgi_img, const uint8_t *__restrict__ dgh_img,
                                                             int Prb, int KG, int G, int dirs, int n_split, int KS,
                                                             float *__restrict__ part) {
  __shared__ float red[8][64];
  const int split = blockIdx.x, cb = blockIdx.y, which = blockIdx.z & 1, dir = blockIdx.z >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_total = Prb * 2, per = (k_total + n_split - 1) / n_split;
  const int k_lo = split * per, k_hi = min(k_total, k_lo + per);
  const uint8_t *img = (which ? dgh_img : dgi_img) + (int64_t)dir * Prb * KG * BLK2;
  float a0 = 0.f, a1 = 0.f;
  for (int r = k_lo * 64 + warp; r < k_hi * 64; r += 8) {  // global position row
    const uint8_t *blk = img + ((int64_t)(r >> 7) * KG + cb) * BLK2;
    const uint32_t off = tc::sw128_off(r & 127, 2 * lane);
    const uint32_t hi = *reinterpret_cast<const uint32_t *>(blk + off), lo = *reinterpret_cast<const uint32_t *>(blk + BLK + off);
    a0 += __uint_as_float(hi << 16) + __uint_as_float(lo << 16);
    a1 += __uint_as_float(hi & 0xFFFF0000u) + __uint_as_float(lo & 0xFFFF0000u);
  }
  red[warp][2 * lane] = a0; red[warp][2 * lane + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) acc += red[w][threadIdx.x];
    part[((((int64_t)split * dirs + dir) * 2 + which) * G + cb * 64 + threadIdx.x) * KS + KS - 1] = acc;
  }
}

}  // namespace gtc

// ------------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------------
bool gru_tc_supported(const rec_engine *e) {
  const rec_config &c = e->cfg;
  return e->use_tc && c.embedding_dim % 128 == 0 && c.hidden_dim % 128 == 0 && c.embedding_dim <= 512 && c.hidden_dim <= 256;  // H / 32 <= 8 CTAs per cluster, resident W_hh slices sized for H <= 256
}

static int gtc_alloc(rec_engine *e, void **ptr, size_t bytes) {
  if (*ptr) return REC_OK;
  cudaError_t st = cudaMalloc(ptr, bytes);
  if (st != cudaSuccess) REC_FAIL(e, REC_ENOMEM, "cudaMalloc(%zu B) for the tensor-core GRU buffers failed: %s", bytes, cudaGetErrorString(st));
  // zero once: positions a step never writes (padding rows, tokens beyond a session's length) must hold finite values
  // -- they meet zero gradients in the weight-gradient GEMMs, and 0 x NaN would poison the sums
  REC_CUDA(e, cudaMemsetAsync(*ptr, 0, bytes, e->stream));
  return REC_OK;
}

void gtc_free(rec_engine *e) {
  void *ptrs[] = {e->g_ximg[0], e->g_ximg[1], e->g_gi[0], e->g_gi[1], e->g_gi[2], e->g_himg[0], e->g_himg[1], e->g_dstep[0],
                  e->g_dstep[1], e->g_dgi_img, e->g_dgh_img, e->g_hprev_img, e->g_dhw, e->g_wimg, e->g_ximg2};
  for (void *p : ptrs) if (p) cudaFree(p);
}

struct GtcDims {
  int E, H, G, L, dirs, KBe, KBh, KG, n_sb, Prb;
  int64_t P;
  size_t wih_bytes, whh_bytes, slot_bytes;
};
static GtcDims gtc_dims(const rec_engine *e, int B) {
  GtcDims d;
  const rec_config &c = e->cfg;
  d.E = c.embedding_dim; d.H = c.hidden_dim; d.G = 3 * d.H; d.L = c.state_size; d.dirs = e->dirs;
  d.KBe = d.E / 64; d.KBh = d.H / 64; d.KG = d.G / 64;
  d.n_sb = cdiv(B, 128);
  d.P = (int64_t)B * d.L;
  d.Prb = (int)cdiv64(d.P, 128);
  d.wih_bytes = (size_t)(d.G / 128) * d.KBe * tck::BLK2;
  d.whh_bytes = (size_t)(d.G / 128) * d.KBh * tck::BLK2;
  d.slot_bytes = d.wih_bytes + 2 * d.whh_bytes;  // W_ih | W_hh natural | W_hh regrouped (same size)
  return d;
}
// weight images of (net, dir): [W_ih | W_hh | W_hh regrouped]
static uint8_t *gtc_wslot(const rec_engine *e, const GtcDims &d, int net, int dir) {
  return e->g_wimg + (size_t)(net * 2 + dir) * d.slot_bytes;
}

static int gtc_pack_weights(rec_engine *e, const GtcDims &d, const int *nets, int n_nets) {
  gtc::WPackArgs a = {};
  a.E = d.E; a.H = d.H;
  int n = 0;
  for (int i = 0; i < n_nets; ++i)
    for (int dir = 0; dir < d.dirs; ++dir) {
      const rec_net_params &p = e->nets[nets[i]].p;
      uint8_t *slot = gtc_wslot(e, d, nets[i], dir);
      a.wih[n] = p.w_ih[dir]; a.whh[n] = p.w_hh[dir];
      a.wih_img[n] = slot; a.whh_img[n] = slot + d.wih_bytes; a.whh_perm[n] = slot + d.wih_bytes + d.whh_bytes;
      ++n;
    }
  a.n_slots = n;
  const int chunks = d.G * ((d.E > d.H ? d.E : d.H) / 8);
  gtc::gru_pack_weights_kernel<<<dim3(cdiv(chunks, 256), 3, n), 256, 0, e->stream>>>(a);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

static int gtc_ensure_fwd(rec_engine *e, int n_pass) {
  const GtcDims d = gtc_dims(e, e->cfg.max_batch);
  int rc;
  if ((rc = gtc_alloc(e, (void **)&e->g_wimg, (size_t)REC_MAX_NETS * 2 * d.slot_bytes))) return rc;
  for (int i = 0; i < (n_pass > 1 ? 2 : 1); ++i)
    if ((rc = gtc_alloc(e, (void **)&e->g_ximg[i], (size_t)d.Prb * d.KBe * tck::BLK2))) return rc;
  for (int i = 0; i < n_pass; ++i)
    if ((rc = gtc_alloc(e, (void **)&e->g_gi[i], sizeof(float) * (size_t)d.Prb * 128 * d.dirs * d.G))) return rc;
  for (int i = 0; i < 2; ++i)
    if ((rc = gtc_alloc(e, (void **)&e->g_himg[i], (size_t)3 * d.dirs * d.n_sb * d.KBh * tck::BLK2))) return rc;
  if ((rc = gtc_alloc(e, (void **)&e->g_hprev_img, (size_t)d.dirs * d.Prb * d.KBh * tck::BLK2))) return rc;
  return REC_OK;
}

// Up to three independent passes (main(s), main(s'), boot(s')) advance together, one launch per time step.
int launch_gru_forward_tc(rec_engine *e, int n_pass, const int *net_ids, const int64_t *const *s,
                          const int64_t *const *lengths, float *const *h_out, const bool *save, int B) {
  int rc = gtc_ensure_fwd(e, n_pass);
  if (rc) return rc;
  const rec_config &c = e->cfg;
  const GtcDims d = gtc_dims(e, B);
  // weight images of the nets involved (the parameters change every step)
  int nets[2], n_nets = 0;
  for (int i = 0; i < n_pass; ++i) {
    bool seen = false;
    for (int j = 0; j < n_nets; ++j) seen = seen || nets[j] == net_ids[i];
    if (!seen) nets[n_nets++] = net_ids[i];
  }
  if ((rc = gtc_pack_weights(e, d, nets, n_nets))) return rc;
  // X images: one per distinct input sequence
  int ximg_of[3];
  const int64_t *xs[2] = {nullptr, nullptr};
  int n_x = 0;
  for (int i = 0; i < n_pass; ++i) {
    int k = -1;
    for (int j = 0; j < n_x; ++j) if (xs[j] == s[i]) k = j;
    if (k < 0) {
      if (n_x == 2) REC_FAIL(e, REC_EINVAL, "GRU forward: at most two distinct input sequences per launch");
      k = n_x++;
      xs[k] = s[i];
      const int64_t n_chunks = (int64_t)d.Prb * 128 * (d.E / 8);
      int64_t blocks = cdiv64(n_chunks, 256);
      if (blocks > (int64_t)e->sm_count * 16) blocks = (int64_t)e->sm_count * 16;
      // the table of the FIRST net that reads this sequence: main(s') and boot(s') share s' but not the table
      gtc::gather_pack_kernel<<<(int)blocks, 256, 0, e->stream>>>(e->nets[net_ids[i]].p.emb, s[i], (int)d.P, d.E, c.item_num, e->g_ximg[k]);
      REC_LAUNCH_CHECK(e);
    }
    ximg_of[i] = k;
  }
  // main(s') and boot(s') read DIFFERENT embedding tables: a pass whose table differs from the one its sequence was
  // gathered with needs its own image
  for (int i = 0; i < n_pass; ++i)
    for (int j = 0; j < i; ++j)
      if (s[i] == s[j] && e->nets[net_ids[i]].p.emb != e->nets[net_ids[j]].p.emb && ximg_of[i] == ximg_of[j]) {
        if ((rc = gtc_alloc(e, (void **)&e->g_ximg2, (size_t)gtc_dims(e, c.max_batch).Prb * d.KBe * tck::BLK2))) return rc;
        const int64_t n_chunks = (int64_t)d.Prb * 128 * (d.E / 8);
        int64_t blocks = cdiv64(n_chunks, 256);
        if (blocks > (int64_t)e->sm_count * 16) blocks = (int64_t)e->sm_count * 16;
        gtc::gather_pack_kernel<<<(int)blocks, 256, 0, e->stream>>>(e->nets[net_ids[i]].p.emb, s[i], (int)d.P, d.E, c.item_num, e->g_ximg2);
        REC_LAUNCH_CHECK(e);
        ximg_of[i] = 2;
      }
  auto ximg = [&](int k) { return k == 2 ? e->g_ximg2 : e->g_ximg[k]; };
  // gi = X . W_ih^T + b_ih for every position, pass and direction
  {
    gtc::GemmParams g = {};
    const int Z = n_pass * d.dirs;
    for (int i = 0; i < n_pass; ++i)
      for (int dir = 0; dir < d.dirs; ++dir) {
        const int z = i * d.dirs + dir;
        g.A[z] = gtc_wslot(e, d, net_ids[i], dir);
        g.B[z] = ximg(ximg_of[i]);
        g.C[z] = e->g_gi[i] + (int64_t)dir * d.G;
        g.bias[z] = e->nets[net_ids[i]].p.b_ih[dir];
      }
    // computed transposed (accumulator rows = gate index): coalesced 128-byte stores into gi[p, dir, :]
    g.a_mn = 0; g.b_mn = 0; g.a_cbs = d.KBe; g.b_cbs = d.KBe; g.c_trans = 1;
    g.NT = d.Prb % 2 == 0 ? 256 : 128;
    g.m_tiles = d.G / 128; g.n_tiles = (d.Prb * 128) / g.NT; g.k_total = d.KBe; g.n_split = 1;
    g.M = d.Prb * 128;  // the gi buffer is padded to whole row blocks
    g.ldc = (int64_t)d.dirs * d.G; g.c_split_stride = 0;
    const int total = g.m_tiles * g.n_tiles;
    int n_cta = e->sm_count / Z;
    if (n_cta > total) n_cta = total;
    if (n_cta < 1) n_cta = 1;
    if ((rc = tck::launch_tck<gtc::Gemm>(e, dim3(n_cta, 1, Z), g))) return rc;
  }
  // time steps
  REC_CUDA(e, cudaMemsetAsync(e->g_himg[0], 0, (size_t)n_pass * d.dirs * d.n_sb * d.KBh * tck::BLK2, e->stream));
  gtc::StepParams sp = {};
  for (int i = 0; i < n_pass; ++i)
    for (int dir = 0; dir < d.dirs; ++dir) {
      gtc::StepPass &P = sp.pass[i * d.dirs + dir];
      P.whh_perm = gtc_wslot(e, d, net_ids[i], dir) + d.wih_bytes + d.whh_bytes;
      P.gi = e->g_gi[i]; P.b_hh = e->nets[net_ids[i]].p.b_hh[dir]; P.lens = lengths[i]; P.h_state = h_out[i];
      P.save = save[i] ? 1 : 0;
    }
  sp.gates_save = e->gates_save; sp.hprev_save = e->hprev_save; sp.hprev_img = e->g_hprev_img;
  sp.KBh = d.KBh; sp.B = B; sp.L = d.L; sp.H = d.H; sp.dirs = d.dirs; sp.packed = c.use_packed_seq; sp.n_sb = d.n_sb; sp.Prb = d.Prb;
  sp.himg[0] = e->g_himg[0]; sp.himg[1] = e->g_himg[1];
  {
    static const int sel = getenv("REC_TRACE_SEL") ? atoi(getenv("REC_TRACE_SEL")) : 0;
    sp.trace = sel == 3 ? e->trace : nullptr;
  }
  return tck::launch_tck<gtc::GruStep>(e, dim3(d.H / 32, d.n_sb, n_pass * d.dirs), sp);
}

// stages as launch_gru_backward: 1 = BPTT (+ dx), 2 = weight gradients (split-K partials in e->wgrad_part)
int launch_gru_backward_tc(rec_engine *e, int net_id, const int64_t *s, const int64_t *lengths, int B, const float *dh,
                           int stages) {
  (void)s;
  const rec_config &c = e->cfg;
  const GtcDims dm = gtc_dims(e, c.max_batch), d = gtc_dims(e, B);
  int rc;
  const size_t step_bytes = (size_t)dm.dirs * dm.n_sb * dm.KG * tck::BLK2, big_bytes = (size_t)dm.dirs * dm.Prb * dm.KG * tck::BLK2;
  for (int i = 0; i < 2; ++i) if ((rc = gtc_alloc(e, (void **)&e->g_dstep[i], step_bytes))) return rc;
  if ((rc = gtc_alloc(e, (void **)&e->g_dgi_img, big_bytes))) return rc;
  if ((rc = gtc_alloc(e, (void **)&e->g_dgh_img, big_bytes))) return rc;
  if ((rc = gtc_alloc(e, (void **)&e->g_dhw, sizeof(float) * (size_t)c.max_batch * e->D))) return rc;
  if (!e->g_wimg || !e->g_hprev_img || !e->g_ximg[0]) REC_FAIL(e, REC_EINVAL, "GRU backward without a forward pass on this engine");
  const size_t used_big = (size_t)d.Prb * d.KG * tck::BLK2;  // per direction, for THIS batch
  if (stages & 1) {
    for (int dir = 0; dir < d.dirs; ++dir) {
      REC_CUDA(e, cudaMemsetAsync(e->g_dgi_img + dir * used_big, 0, used_big, e->stream));
      REC_CUDA(e, cudaMemsetAsync(e->g_dgh_img + dir * used_big, 0, used_big, e->stream));
    }
    REC_CUDA(e, cudaMemsetAsync(e->g_dstep[(d.L - 1 + 1) & 1], 0, (size_t)d.dirs * d.n_sb * d.KG * tck::BLK2, e->stream));
    REC_CUDA(e, cudaMemcpyAsync(e->g_dhw, dh, sizeof(float) * (size_t)B * e->D, cudaMemcpyDeviceToDevice, e->stream));
    gtc::BpttParams bp = {};
    for (int dir = 0; dir < d.dirs; ++dir) bp.whh_img[dir] = gtc_wslot(e, d, net_id, dir) + d.wih_bytes;
    bp.dhw = e->g_dhw; bp.gates_save = e->gates_save; bp.hprev_save = e->hprev_save; bp.lens = lengths;
    bp.dgi_img = e->g_dgi_img; bp.dgh_img = e->g_dgh_img;
    bp.Prb = d.Prb; bp.KG = d.KG; bp.KBh = d.KBh; bp.B = B; bp.L = d.L; bp.H = d.H; bp.dirs = d.dirs; bp.packed = c.use_packed_seq;
    bp.n_sb = d.n_sb;
    bp.dstep[0] = e->g_dstep[0]; bp.dstep[1] = e->g_dstep[1];  // step L - 1 reads the zeroed image [L & 1]
    if ((rc = tck::launch_tck<gtc::GruBptt>(e, dim3(d.H / 32, d.n_sb, d.dirs), bp))) return rc;
    // dx[p, dir, :] = dgi[p, :] . W_ih
    gtc::GemmParams g = {};
    for (int dir = 0; dir < d.dirs; ++dir) {
      g.A[dir] = gtc_wslot(e, d, net_id, dir);
      g.B[dir] = e->g_dgi_img + dir * used_big;
      g.C[dir] = e->dx + (int64_t)dir * d.E;
      g.bias[dir] = nullptr;
    }
    // transposed: accumulator rows = embedding column (A = MN-major view of the W_ih image), columns = positions
    g.a_mn = 1; g.b_mn = 0; g.a_cbs = d.KBe; g.b_cbs = d.KG; g.c_trans = 1;
    g.NT = d.Prb % 2 == 0 ? 256 : 128;
    g.m_tiles = d.E / 128; g.n_tiles = (d.Prb * 128) / g.NT; g.k_total = d.KG; g.n_split = 1;
    g.M = (int)d.P; g.ldc = (int64_t)d.dirs * d.E; g.c_split_stride = 0;
    const int total = g.m_tiles * g.n_tiles;
    int n_cta = e->sm_count / d.dirs;
    if (n_cta > total) n_cta = total;
    if ((rc = tck::launch_tck<gtc::Gemm>(e, dim3(n_cta, 1, d.dirs), g))) return rc;
  }
  if (stages & 2) {
    const int KS = (d.E > d.H ? d.E : d.H) + 1;
    int splits = 2 * d.Prb < 12 ? 2 * d.Prb : 12;   // split-K slices actually written (gru_adam_kernel sums e->wgrad_used)
    if (splits > e->wgrad_splits) splits = e->wgrad_splits;
    e->wgrad_used = splits;
    for (int which = 0; which < 2; ++which) {
      gtc::GemmParams g = {};
      const int N = which ? d.H : d.E;
      for (int dir = 0; dir < d.dirs; ++dir) {
        g.A[dir] = which ? e->g_hprev_img + (size_t)dir * d.Prb * d.KBh * tck::BLK2 : e->g_ximg[0];
        g.B[dir] = (which ? e->g_dgh_img : e->g_dgi_img) + dir * used_big;
        g.C[dir] = e->wgrad_part + ((int64_t)dir * 2 + which) * d.G * KS;
        g.bias[dir] = nullptr;
      }
      // transposed: accumulator rows = input feature k, columns = gate row j -> part[j * KS + k], coalesced along k
      g.a_mn = 1; g.b_mn = 1; g.a_cbs = N / 64; g.b_cbs = d.KG; g.c_trans = 1;
      g.NT = d.G % 256 == 0 ? 256 : 128;
      g.m_tiles = N / 128; g.n_tiles = d.G / g.NT; g.k_total = d.Prb * 2; g.n_split = splits;
      g.M = d.G; g.ldc = KS; g.c_split_stride = (int64_t)d.dirs * 2 * d.G * KS;
      const int total = g.m_tiles * g.n_tiles * splits;
      int n_cta = e->sm_count / d.dirs;
      if (n_cta > total) n_cta = total;
      if ((rc = tck::launch_tck<gtc::Gemm>(e, dim3(n_cta, 1, d.dirs), g))) return rc;
    }
    gtc::gru_bias_colsum_kernel<<<dim3(splits, d.KG, 2 * d.dirs), 256, 0, e->stream>>>(e->g_dgi_img, e->g_dgh_img, d.Prb, d.KG, d.G, d.dirs,
                                                                                  splits, KS, e->wgrad_part);
    REC_LAUNCH_CHECK(e);
  }
  return REC_OK;
}
