// Blackwell (sm_100a) tensor-core plumbing written as inline PTX: tcgen05.mma with TMEM accumulators,
// shared-memory matrix descriptors, mbarrier completion, TMEM allocation and tcgen05.ld.
//
// Operand tiles in shared memory all have ONE physical form: blocks of [rows][64 bf16] (128-byte rows)
// with the 128-byte swizzle (16-byte chunk index XOR (row % 8)), 1024-byte aligned.  The same bytes can
// be described to the tensor core either as a K-major operand (MN = rows, K = the 64 columns) or as an
// MN-major operand (MN = the 64 columns, K = rows) -- which is what lets h, W and dlogits each live in
// shared memory once while feeding three different GEMMs (logits, dW, dh).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a lost completion traps (fails the launch) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try(bar, parity); ++spins)
    if (spins > (1u << 24)) __trap();
}

// ---- TMA 1-D bulk copy global -> shared with mbarrier transaction count ---------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// dst/src 16-byte aligned, bytes a multiple of 16; completes on `bar` (complete_tx)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// L2 prefetch of a contiguous range (16-byte aligned, bytes a multiple of 16); no completion tracking
__device__ __forceinline__ void l2_prefetch(const void *p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- proxies / fences ----------------------------------------------------------------------------
// generic-proxy smem writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ----------------------------------------------------------------------------------------
// one full warp allocates `cols` (power of two >= 32) columns; base address is written to *dst (smem)
__device__ __forceinline__ void tmem_alloc(uint32_t *dst, uint32_t cols) {
  __syncwarp();
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  __syncwarp();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// TMEM -> registers.  32 lanes x N consecutive 32-bit columns: thread i of the warp gets lane
// (32*(warp%4) + i).  tcgen05.ld is ASYNCHRONOUS: the destination registers are only valid after
// tcgen05.wait::ld.  Both instructions live in ONE asm statement so that the compiler can never move,
// copy or spill the destination registers between the load and the wait (it cannot see the asynchrony).
// (.sync.aligned: the whole warp must execute them convergently -> __syncwarp first.)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  __syncwarp();
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  __syncwarp();
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// kept for call sites that still pair an explicit wait with the loads above (now a no-op barrier)
__device__ __forceinline__ void tmem_ld_wait() { __syncwarp(); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ---------------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor (PTX "matrix descriptor", Blackwell version = 1, SWIZZLE_128B)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // layout type: SWIZZLE_128B
  return d;
}
// Un-swizzled K-major operand [rows][16 bf16] (one K = 16 slice, 32 bytes per row): core matrices of 8 rows x 16 bytes
// are contiguous (128 B); the two K chunks of a row group sit `lbo` bytes apart, row groups `sbo` bytes apart
// (layout type 0 = SWIZZLE_NONE).  Used for the 4 KB "bias as one more K slice" operands (nosw_off below).
__device__ __forceinline__ uint64_t smem_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// byte offset of element (row, k < 16) in the un-swizzled [rows][16 bf16] operand described with lbo = 128, sbo = 256
__device__ __host__ __forceinline__ uint32_t nosw_off(int row, int k) {
  return (uint32_t)((row >> 3) * 256 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2);
}
// K-major operand: MN = rows of the block (8-row groups 1024 B apart), K = 64 columns of the block
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t block_saddr, int k16) {
  return smem_desc(block_saddr + k16 * 32, 16, 1024);
}
// MN-major operand: MN = columns (next 64 at `mn_block_stride` bytes), K = rows (16 rows per MMA = 2048 B)
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t block_saddr, int k16, uint32_t mn_block_stride) {
  return smem_desc(block_saddr + k16 * 2048, mn_block_stride, 1024);
}
// 32-bit instruction descriptor: bf16 x bf16 -> fp32, dense
__device__ __forceinline__ uint32_t instr_desc(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;                         // D format F32
  d |= 1u << 7;                         // A format BF16
  d |= 1u << 10;                        // B format BF16
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 2^x, flush-to-zero: ONE MUFU.EX2 (the default ex2.approx adds a denormal range check and two multiplies)
__device__ __forceinline__ float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// 2^x on the FMA pipe (no MUFU): round-to-nearest split x = n + f with the 1.5 * 2^23 trick, a cubic for 2^f on
// [-0.5, 0.5] (relative error 8.1e-5: enough for the terms of a log-sum-exp whose SUM is needed to 1e-3), n added to the
// exponent field by one integer multiply-add ((bits of 1.5 * 2^23 + n) << 23 == n << 23 mod 2^32).  x <= ~0; x below
// -125 (incl. -inf of masked columns) is clamped: 2^-125 vanishes next to the row maximum's 2^0.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 0.05533474684f, 0.2426231205f);  // minimax cubic (relative error), scratch fit: max 8.1e-5
  p = fmaf(p, f, 0.6932122111f);
  p = fmaf(p, f, 0.9999210238f);
  return __int_as_float(__float_as_int(t) * 0x800000 + __float_as_int(p));
}

// ---- operand staging -----------------------------------------------------------------------------
// byte offset of element (row, col) inside a [rows][64 bf16] SWIZZLE_128B block
__device__ __forceinline__ uint32_t sw128_off(int row, int col) {
  return (uint32_t)(row * 128 + ((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1)));
}
// split fp32 into bf16 hi + bf16 lo (x ~= hi + lo, |err| ~ 2^-17 |x|)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16 &hi, __nv_bfloat16 &lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// two fp32 -> packed bf16x2 hi and lo words (element `a` in the low half).  cvt.rn.bf16x2.f32 converts a pair
// per instruction; the scalar F2F.BF16.F32 it replaces runs at conversion-unit rate (16 lanes/clk/SM).
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t &hi, uint32_t &lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hb), "f"(a - ha));
}
// 8 consecutive fp32 -> one 16-byte bf16 hi chunk + one lo chunk (registers)
__device__ __forceinline__ void split8(const float4 &a, const float4 &b, uint4 &h, uint4 &l) {
  split_bf16x2(a.x, a.y, h.x, l.x);
  split_bf16x2(a.z, a.w, h.y, l.y);
  split_bf16x2(b.x, b.y, h.z, l.z);
  split_bf16x2(b.z, b.w, h.w, l.w);
}
// one 16-byte chunk -> position (row, chunk) of a SWIZZLE_128B block
__device__ __forceinline__ void store_chunk(uint8_t *blk, int row, int chunk, const uint4 &v) {
  *reinterpret_cast<uint4 *>(blk + (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4))) = v;
}
// 8 consecutive fp32 (one 16-byte bf16 chunk) -> hi/lo chunks at (row, col8*8) of two blocks
__device__ __forceinline__ void store_split8(uint8_t *blk_hi, uint8_t *blk_lo, int row, int chunk, const float4 &a,
                                             const float4 &b) {
  uint4 h, l;
  split_bf16x2(a.x, a.y, h.x, l.x);
  split_bf16x2(a.z, a.w, h.y, l.y);
  split_bf16x2(b.x, b.y, h.z, l.z);
  split_bf16x2(b.z, b.w, h.w, l.w);
  uint32_t off = (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
  *reinterpret_cast<uint4 *>(blk_hi + off) = h;
  *reinterpret_cast<uint4 *>(blk_lo + off) = l;
}

}  // namespace tc
