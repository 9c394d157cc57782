// Skeleton of the K-loop tcgen05 GEMM pipelines over packed bf16 hi/lo operand images (see heads_tck.cu for the image
// format): packing kernels, the warp-specialised kernel template and its launcher.  Included by heads_tck.cu (wide
// heads) and gru_tc.cu (GRU trunk for E, H >= 128).
#pragma once
#include <type_traits>
#include "common.cuh"
#include "tc.cuh"

namespace tck {

constexpr int BLK = 16384;      // one [128][64] bf16 operand block
constexpr int BLK2 = 2 * BLK;   // hi + lo
constexpr int HALF = 8192;      // 64 rows of a block
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = EPI_THREADS + 64;  // + issuer warp + loader warp
constexpr float LOG2E = 1.4426950408889634f;
constexpr int ROW_STRIDE_ = 8;  // row_stats stride (heads.cu)
constexpr int TOPK_OFF = 5;     // part record layout (heads.cu / heads_tc.cu)

// ------------------------------------------------------------------------------------------------------------
// Packing
// ------------------------------------------------------------------------------------------------------------
struct PackSrc {
  const float *p[3];
  float w[3];
  int n;
};

// fp32 row-major [R, C] (leading dimension ld) -> image; rows in [R, 128 * ceil(R/128)) are written as zeros.
static __global__ void __launch_bounds__(256) pack_img_kernel(PackSrc s, int R, int C, int64_t ld, uint8_t *__restrict__ img) {
  const int c8n = C >> 3, KB = C >> 6;
  const int64_t n_chunks = (int64_t)((R + 127) / 128) * 128 * c8n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // two chunks (4 x 16 B loads per source) in flight per thread: the grid stays small (4 CTAs per SM) so that the
  // latency-bound kernels of other graph branches find free warp slots next to this HBM-bound sweep
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_chunks; i0 += 2 * stride) {
    float4 a[2], b[2];
    int64_t row[2];
    int c8[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = i0 + u * stride;
      row[u] = i / c8n;
      c8[u] = (int)(i - row[u] * c8n);
      a[u] = make_float4(0.f, 0.f, 0.f, 0.f); b[u] = a[u];
      if (i < n_chunks && row[u] < R) {
        const float4 *p0 = reinterpret_cast<const float4 *>(s.p[0] + row[u] * ld + c8[u] * 8);
        a[u] = p0[0]; b[u] = p0[1];
        if (s.n > 1 || s.w[0] != 1.f) {
          const float w0 = s.w[0];
          a[u].x *= w0; a[u].y *= w0; a[u].z *= w0; a[u].w *= w0; b[u].x *= w0; b[u].y *= w0; b[u].z *= w0; b[u].w *= w0;
          for (int j = 1; j < s.n; ++j) {
            const float4 *pj = reinterpret_cast<const float4 *>(s.p[j] + row[u] * ld + c8[u] * 8);
            const float4 x = pj[0], y = pj[1];
            const float w = s.w[j];
            a[u].x = fmaf(w, x.x, a[u].x); a[u].y = fmaf(w, x.y, a[u].y); a[u].z = fmaf(w, x.z, a[u].z); a[u].w = fmaf(w, x.w, a[u].w);
            b[u].x = fmaf(w, y.x, b[u].x); b[u].y = fmaf(w, y.y, b[u].y); b[u].z = fmaf(w, y.z, b[u].z); b[u].w = fmaf(w, y.w, b[u].w);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (i0 + u * stride >= n_chunks) continue;
      uint8_t *blk = img + ((row[u] >> 7) * KB + (c8[u] >> 3)) * (int64_t)BLK2;
      tc::store_split8(blk, blk + BLK, (int)(row[u] & 127), c8[u] & 7, a[u], b[u]);
    }
  }
}

// h [B, D] -> image of h^T: rows = state dimension d (D % 128 == 0), columns = sessions (KBS = ceil(B/64) blocks,
// sessions beyond B are zeros).  The matrix is tiny (<= 512 KB): strided reads are fine.
static __global__ void __launch_bounds__(256) pack_img_T_kernel(const float *__restrict__ h, int B, int D, int KBS,
                                                        uint8_t *__restrict__ img) {
  const int c8n = KBS * 8;
  const int n_chunks = D * c8n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_chunks; i += gridDim.x * blockDim.x) {
    const int c8 = i / D, row = i - c8 * D;  // consecutive threads: consecutive d -> coalesced reads of h rows
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = c8 * 8 + j;
      v[j] = b < B ? h[(int64_t)b * D + row] : 0.f;
    }
    uint8_t *blk = img + ((int64_t)(row >> 7) * KBS + (c8 >> 3)) * BLK2;
    tc::store_split8(blk, blk + BLK, row & 127, c8 & 7, make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]));
  }
}

static __global__ void bias_combine_kernel(PackSrc s, int n, float *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = s.w[0] * s.p[0][i];
  for (int j = 1; j < s.n; ++j) acc = fmaf(s.w[j], s.p[j][i], acc);
  out[i] = acc;
}

// ---- thread-block-cluster helpers (time-step loops: the CTAs of a cluster exchange a step's result through global
// memory and signal each other through mbarriers in distributed shared memory) ---------------------------------------
__device__ __forceinline__ uint32_t cluster_size_x() { uint32_t v; asm volatile("mov.u32 %0, %%cluster_nctaid.x;" : "=r"(v)); return v; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (release, cluster scope) on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(tc::smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
    if (ok) break;
    if (spins > (1u << 24)) __trap();
  }
}
// generic-proxy writes (of other CTAs, made visible by an acquire) -> async-proxy (TMA) reads of this thread
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------------------
// Kernel skeleton
// ------------------------------------------------------------------------------------------------------------
// OP supplies:  STAGES, STAGE_BYTES, ACC_COLS (TMEM columns of one accumulator), TMEM_COLS (allocation, power of 2),
//   units(p, lo, hi)            this CTA's range of output units (an accumulator's worth of output each)
//   k_steps(p, u)               pipeline stages consumed by unit u
//   load(p, u, ks, stage, bar)  ONE thread: expect_tx + TMA bulk copies of stage (u, ks)
//   mma(p, u, ks, saddr, tacc, first)  ONE thread: the tcgen05.mma's of stage (u, ks) into accumulator tacc
//   Epi                         per-thread epilogue object: tile(p, u, i, tacc) per unit, finish(p) at the end
//   CLUSTERED                   true: the grid's x extent is ONE thread-block cluster and units are TIME STEPS: unit u + 1
//                               of every CTA reads (load) what the epilogues of ALL CTAs of the cluster wrote to global
//                               memory in unit u: the loader waits for the cluster-wide "step done" barrier per unit.
//   RESIDENT_BYTES              > 0: an operand that is the same for every unit (the recurrent weights of this CTA's
//                               hidden units) is loaded ONCE by load_resident(p, smem, bar) and stays in shared memory;
//                               mma() receives its address
// Optional clock64 phase stamps: an OP may define  static void trace(const Params &, int kind, int i)  (kinds: 0 issuer
// has its accumulator, 1 issuer has its stage, 2 MMAs issued, 3 / 4 epilogue warp 0 starts / ends a unit, 5 loader has
// a free stage, 6 / 7 the last epilogue warp).  OPs without it compile to nothing.
template <class OP, class = void>
struct op_has_trace : std::false_type {};
template <class OP>
struct op_has_trace<OP, std::void_t<decltype(&OP::trace)>> : std::true_type {};
template <class OP>
__device__ __forceinline__ void op_trace(const typename OP::Params &p, int kind, int i) {
  if constexpr (op_has_trace<OP>::value) OP::trace(p, kind, i);
}

// Optional  static constexpr int ISSUERS = 2:  two MMA-issuer warps take alternate units (= alternate accumulators).  The
// issuing thread blocks in tcgen05.mma for about as long as the tensor pipe is busy and then spends ~1000 cycles per unit
// in commit / mbarrier-wait / fence round trips (clock64 stamps of HeadCmaxPair: 26 MMAs = 1970 cycles issuing + 1080
// cycles of synchronisation during which the tensor pipe idles); with two issuers one warp's round trips overlap the
// other's MMAs.  Units are independent (different accumulators, per-thread tcgen05.commit): no ordering between them.
// Only for OPs with ONE k-step per unit and STAGES % ISSUERS == 0: every stage is then consumed by the same issuer each
// time round the ring, so its parity waits cannot alias (with 3 stages and 8 k-steps per unit the second issuer, which
// starts at stage 8 % 3 of round 2, would take the FIRST completion of that stage's barrier for its own -- measured: a
// launch failure in HeadFwd).
template <class OP, class = void>
struct op_issuers : std::integral_constant<int, 1> {};
template <class OP>
struct op_issuers<OP, std::void_t<decltype(OP::ISSUERS)>> : std::integral_constant<int, OP::ISSUERS> {};

template <class OP>
__global__ void __launch_bounds__(OP::EPI_WARPS * 32 + 32 + 32 * op_issuers<OP>::value, 1) tck_kernel(const typename OP::Params p) {
  constexpr int EPI_WARPS = OP::EPI_WARPS;  // 8 (default) or 16 epilogue warps + issuer warp(s) + loader warp
  constexpr int ISSUERS = op_issuers<OP>::value;
  static_assert(ISSUERS == 1 || OP::STAGES % ISSUERS == 0, "two issuers: each stage must always belong to the same issuer");
  extern __shared__ uint8_t raw[];
  uint8_t *sm = raw + ((1024u - (tc::smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t full[OP::STAGES], empty[OP::STAGES], tfull[2], tempty[2], stepbar, resbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int u_lo = 0, u_hi = 0;
  OP::units(p, u_lo, u_hi);
  const uint32_t csize = OP::CLUSTERED ? cluster_size_x() : 1u;
  if (tid == 0) {
    for (int s = 0; s < OP::STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&tfull[b], 1); tc::mbar_init(&tempty[b], EPI_WARPS); }
    tc::mbar_init(&stepbar, csize * EPI_WARPS);
    tc::mbar_init(&resbar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, OP::TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (OP::CLUSTERED) cluster_sync_all();  // every CTA's barriers are initialised before a peer may arrive on them
  const uint32_t tmem = tmem_base_s;
  uint8_t *resident = sm + OP::STAGES * OP::STAGE_BYTES;
  uint8_t *extra = resident + OP::RESIDENT_BYTES;

  if (warp == EPI_WARPS + ISSUERS) {
    // ---- TMA loader ----
    if (lane == 0) {
      int s = 0, i = 0;
      uint32_t round = 0;
      if constexpr (!OP::CLUSTERED && OP::RESIDENT_BYTES > 0) OP::load_resident(p, resident, &resbar);
      for (int u = u_lo; u < u_hi; ++u, ++i) {
        const int nk = OP::k_steps(p, u);
        if constexpr (OP::CLUSTERED) {
          if (i == 0) {
            if constexpr (OP::RESIDENT_BYTES > 0) OP::load_resident(p, resident, &resbar);
          } else {  // the previous step of every CTA of the cluster is in global memory
            mbar_wait_cluster(&stepbar, (uint32_t)(i - 1) & 1u);
            fence_proxy_async_all();
          }
          for (int ks = 0; ks < nk; ++ks) {
            if (round > 0) tc::mbar_wait(&empty[s], (round - 1) & 1);
            OP::load(p, u, ks, sm + s * OP::STAGE_BYTES, &full[s]);
            if (++s == OP::STAGES) { s = 0; ++round; }
          }
        } else {
          for (int ks = 0; ks < nk; ++ks) {
            if (round > 0) tc::mbar_wait(&empty[s], (round - 1) & 1);  // the MMAs that read this stage are complete
            op_trace<OP>(p, 5, i);
            OP::load(p, u, ks, sm + s * OP::STAGE_BYTES, &full[s]);
            if (++s == OP::STAGES) { s = 0; ++round; }
          }
        }
      }
    }
  } else if (warp >= EPI_WARPS) {
    // ---- MMA issuer(s) ----
    int s = 0, i = 0;
    uint32_t round = 0;
    for (int u = u_lo; u < u_hi; ++u, ++i) {
      const int b = i & 1;
      const int nk = OP::k_steps(p, u);
      if (ISSUERS > 1 && (i % ISSUERS) != warp - EPI_WARPS) {  // the other issuer's unit: only track the stage ring
        for (int ks = 0; ks < nk; ++ks) if (++s == OP::STAGES) { s = 0; ++round; }
        continue;
      }
      if constexpr (OP::RESIDENT_BYTES > 0) { if (i < ISSUERS) tc::mbar_wait(&resbar, 0); }
      for (int ks = 0; ks < nk; ++ks) {
        // the stage has usually landed long ago: its wait (an mbarrier round trip of ~150 cycles even when satisfied)
        // comes BEFORE the wait for the accumulator, which is the one on the critical path
        tc::mbar_wait(&full[s], round & 1);
        if (ks == 0) {
          if (lane == 0) op_trace<OP>(p, 0, i);
          if (i >= 2) tc::mbar_wait(&tempty[b], ((i - 2) >> 1) & 1);  // the epilogue has read this accumulator
        }
        tc::tc_fence_after();
        if (lane == 0) {
          op_trace<OP>(p, 1, i);
          if constexpr (OP::RESIDENT_BYTES > 0)
            OP::mma(p, u, ks, tc::smem_u32(sm + s * OP::STAGE_BYTES), tc::smem_u32(resident), tmem + (uint32_t)(b * OP::ACC_COLS), ks == 0);
          else
            OP::mma(p, u, ks, tc::smem_u32(sm + s * OP::STAGE_BYTES), tmem + (uint32_t)(b * OP::ACC_COLS), ks == 0);
          op_trace<OP>(p, 2, i);
          tc::mma_commit(&empty[s]);
        }
        __syncwarp();
        if (++s == OP::STAGES) { s = 0; ++round; }
      }
      if (lane == 0) tc::mma_commit(&tfull[b]);
      __syncwarp();
    }
  } else {
    // ---- epilogue warps ----
    typename OP::Epi epi(p, extra, tid);
    int i = 0;
    for (int u = u_lo; u < u_hi; ++u, ++i) {
      const int b = i & 1;
      tc::mbar_wait(&tfull[b], (i >> 1) & 1);
      tc::tc_fence_after();
      if (tid == 0) op_trace<OP>(p, 3, i);
      if (tid == EPI_WARPS * 32 - 32) op_trace<OP>(p, 6, i);
      epi.tile(p, u, i, tmem + (uint32_t)(b * OP::ACC_COLS));
      if (tid == 0) op_trace<OP>(p, 4, i);
      if (tid == EPI_WARPS * 32 - 32) op_trace<OP>(p, 7, i);
      tc::tc_fence_before();
      if (OP::CLUSTERED) __threadfence();
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(&tempty[b]);
        if (OP::CLUSTERED)  // this warp's share of the step is in global memory: tell every CTA of the cluster
          for (uint32_t r = 0; r < csize; ++r) mbar_arrive_remote(&stepbar, r);
      }
    }
    epi.finish(p);
  }
  if (OP::CLUSTERED) cluster_sync_all();  // no CTA leaves (its barriers vanish) while a peer may still arrive on them
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, OP::TMEM_COLS);
}

template <int NTHREADS = EPI_THREADS>
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

template <class OP>
static int launch_tck(rec_engine *e, dim3 grid, const typename OP::Params &p) {
  const size_t smem = 1024 + (size_t)OP::STAGES * OP::STAGE_BYTES + OP::RESIDENT_BYTES + OP::EXTRA_BYTES;
  constexpr int NTHREADS = OP::EPI_WARPS * 32 + 32 + 32 * op_issuers<OP>::value;
  static bool attr_set[REC_MAX_DEVICES] = {};
  if (!attr_set[e->dev]) {
    REC_CUDA(e, cudaFuncSetAttribute(tck_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[e->dev] = true;
  }
  if (OP::CLUSTERED) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = e->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = grid.x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    REC_CUDA(e, cudaLaunchKernelEx(&cfg, tck_kernel<OP>, p));
  } else {
    tck_kernel<OP><<<grid, NTHREADS, smem, e->stream>>>(p);
  }
  e->launches++;
  if (e->tl_on) rec_timeline_record(e, OP::NAME, 0);
  cudaError_t st = cudaGetLastError();
  if (st != cudaSuccess) REC_FAIL(e, REC_ECUDA, "kernel launch failed: %s (%s)", cudaGetErrorString(st), OP::NAME);
  return REC_OK;
}

}  // namespace tck
