// Self-test of the tcgen05 plumbing (tc.cuh): one CTA computes a small GEMM through each operand
// "major-ness" used by the head kernels, with the bf16x3 split (hi*hi + hi*lo + lo*hi, fp32 accumulate in
// TMEM).  Exposed as rec_debug_tc_gemm for tests/test_gpu_tc.py; not part of the hot path.
//   mode 0: C[128,128] = A[128,64] . B[128,64]^T        (A, B K-major)            -- logits GEMM
//   mode 1: C[128, 64] = P[128k,128m]^T . Q[128k,64n]   (A, B MN-major)            -- dW GEMM
//   mode 2: C[128, 64] = P[128m,128k] . R[128k,64n]     (A K-major, B MN-major)    -- dh GEMM
//   mode 3 (+8: lbo / sbo swapped, must FAIL): mode 0 + bias[j] added by ONE more K = 16 MMA over two un-swizzled 4 KB
//           operands (ones[128][16] . (b_hi, b_lo, b_lo2, 0...)[128][16]^T); bias = B + 128 * 64
#include "common.cuh"
#include "tc.cuh"

__global__ void __launch_bounds__(128) tc_selftest_kernel(int mode, const float *__restrict__ A,
                                                          const float *__restrict__ B, float *__restrict__ C) {
  extern __shared__ uint8_t raw[];
  uint8_t *sm = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int BLK = 128 * 128;  // bytes of one [128][64 bf16] block
  // operand A: up to two blocks (hi) + two blocks (lo); operand B likewise
  uint8_t *a_hi = sm, *a_lo = sm + 2 * BLK, *b_hi = sm + 4 * BLK, *b_lo = sm + 6 * BLK;
  const int variant = mode >> 3;
  mode &= 7;
  const bool with_bias = mode == 3;
  if (with_bias) mode = 0;
  const int a_cols = (mode == 0) ? 64 : 128;   // fp32 columns of the A source matrix (128 rows)
  const int b_rows = 128, b_cols = 64;
  // stage A: [128][a_cols] fp32 -> blocks of [128][64]
  for (int e = tid; e < 128 * (a_cols / 8); e += 128) {
    int row = e / (a_cols / 8), ch = e % (a_cols / 8);
    const float4 *src = reinterpret_cast<const float4 *>(A + (size_t)row * a_cols + ch * 8);
    int blk = ch / 8;
    tc::store_split8(a_hi + blk * BLK, a_lo + blk * BLK, row, ch % 8, src[0], src[1]);
  }
  for (int e = tid; e < b_rows * (b_cols / 8); e += 128) {
    int row = e / (b_cols / 8), ch = e % (b_cols / 8);
    const float4 *src = reinterpret_cast<const float4 *>(B + (size_t)row * b_cols + ch * 8);
    tc::store_split8(b_hi, b_lo, row, ch, src[0], src[1]);
  }
  uint8_t *ones = sm + BLK, *bblk = sm + 3 * BLK;  // second blocks of the A regions (unused by mode 0)
  if (with_bias) {
    for (int e = tid; e < 4096 / 4; e += 128) { reinterpret_cast<uint32_t *>(ones)[e] = 0u; reinterpret_cast<uint32_t *>(bblk)[e] = 0u; }
    __syncthreads();
    const float b = B[128 * 64 + tid];
    const __nv_bfloat16 one = __float2bfloat16_rn(1.f);
    const __nv_bfloat16 h0 = __float2bfloat16_rn(b);
    const float r1 = b - __bfloat162float(h0);
    const __nv_bfloat16 h1 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 h2 = __float2bfloat16_rn(r1 - __bfloat162float(h1));
    const __nv_bfloat16 hv[3] = {h0, h1, h2};
    for (int k = 0; k < 3; ++k) {
      *reinterpret_cast<__nv_bfloat16 *>(ones + tc::nosw_off(tid, k)) = one;
      *reinterpret_cast<__nv_bfloat16 *>(bblk + tc::nosw_off(tid, k)) = hv[k];
    }
  }
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 128);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int N = (mode == 0) ? 128 : 64;
  if (tid == 0) {
    const uint32_t ah = tc::smem_u32(a_hi), al = tc::smem_u32(a_lo), bh = tc::smem_u32(b_hi), bl = tc::smem_u32(b_lo);
    bool acc = false;
    if (mode == 0) {
      const uint32_t id = tc::instr_desc(128, 128, 0, 0);
      for (int pass = 0; pass < 3; ++pass) {
        uint32_t a = pass == 2 ? al : ah, b = pass == 1 ? bl : bh;
        for (int k = 0; k < 4; ++k) { tc::mma_bf16(tmem, tc::desc_kmajor(a, k), tc::desc_kmajor(b, k), id, acc); acc = true; }
      }
      if (with_bias) {
        const uint32_t lbo = variant ? 256 : 128, sbo = variant ? 128 : 256;
        tc::mma_bf16(tmem, tc::smem_desc_nosw(tc::smem_u32(ones), lbo, sbo), tc::smem_desc_nosw(tc::smem_u32(bblk), lbo, sbo), id, true);
      }
    } else if (mode == 1) {
      const uint32_t id = tc::instr_desc(128, 64, 1, 1);
      for (int pass = 0; pass < 3; ++pass) {
        uint32_t a = pass == 2 ? al : ah, b = pass == 1 ? bl : bh;
        for (int k = 0; k < 8; ++k) { tc::mma_bf16(tmem, tc::desc_mnmajor(a, k, BLK), tc::desc_mnmajor(b, k, BLK), id, acc); acc = true; }
      }
    } else {
      const uint32_t id = tc::instr_desc(128, 64, 0, 1);
      for (int pass = 0; pass < 3; ++pass) {
        uint32_t a = pass == 2 ? al : ah, b = pass == 1 ? bl : bh;
        for (int k = 0; k < 8; ++k) {
          // A: K = 128 spans two 64-column blocks; B: K = 128 rows of one block
          tc::mma_bf16(tmem, tc::desc_kmajor(a + (k / 4) * BLK, k % 4), tc::desc_mnmajor(b, k, BLK), id, acc);
          acc = true;
        }
      }
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  const int row = tid;  // TMEM lane == output row
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) C[(size_t)row * N + c0 + j] = v[j];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

extern "C" int rec_debug_tc_gemm(int mode, const float *A, const float *B, float *C, void *stream) {
  if (mode < 0 || (mode > 3 && mode != 11) || !A || !B || !C) return REC_EINVAL;
  const size_t smem = 8 * 128 * 128 + 1024;
  cudaError_t st = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (st != cudaSuccess) return REC_ECUDA;
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mode, A, B, C);
  return cudaGetLastError() == cudaSuccess ? REC_OK : REC_ECUDA;
}
