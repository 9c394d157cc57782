// Full-vocabulary heads on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), D = 64.
//
// head_stats_tc_kernel -- forward statistics of one 128-session block against a range of 128-row
// vocabulary tiles, logits never leave TMEM/registers:
//     logits tile  L[128 x 128] = h[128 x 64] . W_tile[128 x 64]^T          (tcgen05.mma, K-major x K-major)
//     epilogue     online (max, sum-exp) + target logit | running top-k (score desc, id asc)
//                  | running argmax of sum_j w_j Q_j (the heads accumulate into ONE TMEM tile, W_j is
//                  scaled by w_j while it is converted)
// fp32 master weights stream from HBM once per pass; each thread converts its slice to a bf16 hi + bf16 lo
// pair on the way into shared memory (128-byte-swizzled UMMA layout) and the product is evaluated as
// hi*hi + hi*lo + lo*hi with fp32 accumulation ("bf16x3", ~1e-5 relative error -- parity with the fp32
// oracle at 1e-3 needs more than one bf16 pass, and with K = 64 the extra MMAs are nearly free).
// Software pipeline per CTA (all 8 warps cooperate, one elected thread issues the MMAs):
//     convert W(u+1) -> smem[(u+1)&1]  ||  tensor core runs MMA(u)  ;  barrier ;
//     issue MMA(u+1) -> TMEM[(g+1)&1]  ;  wait MMA(u) ; epilogue(u) from TMEM[g&1]
#include "common.cuh"
#include "tc.cuh"

#define PART_TOPK_OFF 5
#define LOG2E_F 1.4426950408889634f
constexpr int BLK = 128 * 128;  // bytes of one [128 rows][64 bf16] operand block

struct TcHeadPtrs {
  const float *w[REC_MAX_HEADS];
  const float *b[REC_MAX_HEADS];
};

// fp32 [rows, 64] row-major (rows r0.. of a matrix with nrows rows) -> hi/lo swizzled blocks, 256 threads
__device__ __forceinline__ void stage_rows64(uint8_t *blk_hi, uint8_t *blk_lo, const float *__restrict__ src, int r0,
                                             int nrows, float scale, int tid) {
  float4 a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = tid + 256 * i, row = c >> 3, ch = c & 7;
    a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    b[i] = a[i];
    if (r0 + row < nrows) {
      const float4 *p = reinterpret_cast<const float4 *>(src + (int64_t)(r0 + row) * 64 + ch * 8);
      a[i] = p[0];
      b[i] = p[1];
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int c = tid + 256 * i, row = c >> 3, ch = c & 7;
    if (scale != 1.f) {
      a[i].x *= scale; a[i].y *= scale; a[i].z *= scale; a[i].w *= scale;
      b[i].x *= scale; b[i].y *= scale; b[i].z *= scale; b[i].w *= scale;
    }
    tc::store_split8(blk_hi, blk_lo, row, ch, a[i], b[i]);
  }
}

// Slow path of the running top-k: insert ONE element that beats the k-th best into the thread-private sorted list
// (shared memory, stride `stride` between ranks).  Callers walk a chunk in ascending column order; together with the
// strict comparisons that keeps equal scores ordered by ascending id.  (A per-chunk routine that re-scanned all 32
// columns from a local-memory copy cost ~350 instructions per call; with 32 independent rows per warp some lane
// takes the slow path in most chunks -- k ln(n/k) insertions per row and thread -- so it doubled the evaluation
// epilogue.)
__device__ __noinline__ void topk_insert_one(float v, int id, float *lv, int *li, int stride, int topk, int &cnt, float &tau) {
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

// NB = 128-session blocks per CTA (they share every converted W tile), NT = compute threads (256: two
// 64-column halves per row, 512: four 32-column quarters per row) + ONE extra warp that only issues MMAs.
// ARG = greedy-action mode: the `nh` Q heads are pre-combined while they are converted,
//   sum_j w_j (h . W_j[a] + b_j[a]) = h . (sum_j w_j W_j[a]) + sum_j w_j b_j[a],
// so a vocabulary tile costs ONE set of MMAs whatever the number of heads.
// TMEM: 2 x NB accumulator tiles of 128 columns (double-buffered over units).
//
// Warp roles (no block-wide barrier inside the main loop -- tcgen05.mma issue blocks the issuing thread
// for about as long as the tensor pipe needs, which used to stall every warp at the next __syncthreads):
//   compute warps: store(u+1) [regs -> bf16 hi/lo smem stage], fetch(u+2) [global -> regs], arrive full[(u+1)&1],
//                  wait done[u&1], epilogue(u) from TMEM[u&1]
//   issuer warp  : wait full[u&1] (all compute warps stored unit u), issue MMAs(u) -> commit done[u&1]
// Stage / TMEM buffer (u&1) is reused by unit u+2: every compute warp stores unit u+2 only after it has seen
// done[u&1] (MMA(u) complete) and finished epilogue(u) in program order.
// RING = the fp32 weight tiles arrive by TMA bulk copies into a ring of `n_slots` 32 KB shared-memory slots filled
// by a LOADER warp (per-thread global loads of a whole tile ran into the SM's outstanding-request limit: ~6000
// cycles per tile); the compute warps read their share of a slot, release it, and convert from registers.
template <int NB, int NT, bool ARG, bool RING>
__global__ void __launch_bounds__(NT + (RING ? (ARG ? 64 : 128) : 32), 1) head_stats_tc_kernel(TcHeadPtrs hp, const float *__restrict__ h, int B, int Vloc,
                                                                   int vocab_lo, int n_tiles, int do_stats, int first_head,
                                                                   int nh, float w0, float w1, float w2,
                                                                   const int64_t *__restrict__ target, int topk,
                                                                   float *__restrict__ part, int part_stride,
                                                                   long long *__restrict__ trace, int n_slots) {
  int tr_n = 0;
#define STRACE(tag) do { if (trace && (ARG || do_stats) && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && tr_n < 120) { asm volatile("" ::: "memory"); trace[2 * tr_n] = (tag); trace[2 * tr_n + 1] = clock64(); ++tr_n; asm volatile("" ::: "memory"); } } while (0)
  constexpr int CS = NT / 128;   // column splits per row
  constexpr int CW = 128 / CS;   // columns per thread and tile
  constexpr int TASKS = 1024 / NT;  // 16-byte chunk pairs per thread when staging a [128 x 64] fp32 tile
  constexpr int NH = (ARG && !RING) ? 3 : 1;   // head tiles a thread keeps in flight per unit (register-fetch path)
  extern __shared__ uint8_t raw[];
  uint8_t *sm = raw + ((1024u - (tc::smem_u32(raw) & 1023u)) & 1023u);  // 1024-aligned; offset arithmetic keeps the pointer provably shared (LDS/STS, not generic LD/ST)
  constexpr int NST = RING ? 1 : 2;  // bf16 W stages (RING: the TMEM double buffer alone overlaps MMA(u+1) with epilogue(u))
  uint8_t *h_blk = sm;                                            // [NB][hi|lo] x BLK
  uint8_t *w_st = sm + NB * 2 * BLK;                              // stage s: hi at w_st + s*2*BLK, lo at + BLK
  float *bias_g = reinterpret_cast<float *>(w_st + NST * 2 * BLK);  // [4][128] (a warp may run up to 3 units ahead of another)
  float *tv = bias_g + 512;                                       // [NB][topk][NT]
  int *ti = reinterpret_cast<int *>(tv + (size_t)NB * topk * NT); // [NB][topk][NT]
  float *stg = reinterpret_cast<float *>(ti + (size_t)NB * topk * NT);  // RING: [n_slots][128 x 64] fp32 (every region before it is a multiple of 128 B)
  float *xch = stg;                                               // [NB][128][CS][5] end-of-kernel exchange (RING: reuses the slots)
  __shared__ uint64_t mbar_done[2], mbar_full[2], mbar_tfree[2], slot_full[4], slot_free[4];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // CONV: two CONVERTER warps turn the ring's fp32 tiles into the bf16 hi/lo stage, so the compute warps run nothing
  // but epilogues (statistics mode; in greedy-action mode the compute warps combine the heads themselves)
  constexpr bool CONV = RING && !ARG;
  constexpr int NPROD = CONV ? 2 : NT / 32;  // warps that fill the bf16 stage / read the ring slots
  const bool issuer = warp == NT / 32, loader = RING && warp == NT / 32 + 1, conv = CONV && warp >= NT / 32 + 2;
  const int q = warp & 3, cq = (warp >> 2) & (CS - 1);  // TMEM lane quarter, column split
  const int sp = blockIdx.x, n_split = gridDim.x, bg = blockIdx.y;
  const int b0 = bg * NB * 128;
  const int per = (n_tiles + n_split - 1) / n_split;
  const int t_lo = sp * per, t_hi = min(n_tiles, t_lo + per);
  const int n_units = max(0, t_hi - t_lo);

  STRACE(19);
  if (tid == 0) {
    tc::mbar_init(&mbar_done[0], 1); tc::mbar_init(&mbar_done[1], 1);
    tc::mbar_init(&mbar_full[0], NPROD); tc::mbar_init(&mbar_full[1], NPROD);
    tc::mbar_init(&mbar_tfree[0], NT / 32); tc::mbar_init(&mbar_tfree[1], NT / 32);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, NB * 256);
  // The loader owns the slot barriers: it initialises them and requests the first n_slots tiles BEFORE the block-wide
  // prologue barrier, so the cold DRAM latency of the first tile overlaps the staging of h.
  int ld_sl = 0, ld_u = 0, ld_hh = 0;
  uint32_t ld_round = 0;
  auto loader_issue = [&]() {  // next (unit, head) tile -> slot ld_sl; one thread
    const int v0 = (t_lo + ld_u) * 128;
    const uint32_t bytes = (uint32_t)min(128, Vloc - v0) * 256u;
    tc::mbar_expect_tx(&slot_full[ld_sl], bytes);
    tc::bulk_g2s(stg + ld_sl * 8192, hp.w[first_head + ld_hh] + (int64_t)v0 * 64, bytes, &slot_full[ld_sl]);
    if (++ld_sl == n_slots) { ld_sl = 0; ++ld_round; }
    if (++ld_hh == nh) { ld_hh = 0; ++ld_u; }
  };
  if (loader && lane == 0) {
    for (int i = 0; i < 4; ++i) { tc::mbar_init(&slot_full[i], 1); tc::mbar_init(&slot_free[i], NPROD); }
    tc::fence_barrier_init();
    while (ld_round == 0 && ld_u < n_units) loader_issue();
  }
  // h blocks of this CTA -> bf16 hi/lo (zero rows beyond B)
  if (tid < NT) {
    for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
      for (int i = 0; i < TASKS; ++i) {
        const int c = tid + NT * i, row = c >> 3, c8 = c & 7;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (b0 + nb * 128 + row < B) {
          const float4 *p = reinterpret_cast<const float4 *>(h + (int64_t)(b0 + nb * 128 + row) * 64 + c8 * 8);
          a = p[0];
          b = p[1];
        }
        tc::store_split8(h_blk + nb * 2 * BLK, h_blk + nb * 2 * BLK + BLK, row, c8, a, b);
      }
    }
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  STRACE(20);

  if (issuer) {
    // ---- MMA issuer warp: NB logits tiles against the same W stage per unit ----
    const uint32_t id_l = tc::instr_desc(128, 128, 0, 0);
    // weight tiles -> L2 a few units ahead of the compute warps' register loads (units 0 and 1 are fetched directly)
    constexpr int PD = 3;
    auto l2_ahead = [&](int u) {
      const int v0 = (t_lo + u) * 128;
      const uint32_t bytes = (uint32_t)min(128, Vloc - v0) * 256u;
      for (int hh = 0; hh < nh; ++hh) tc::l2_prefetch(hp.w[first_head + hh] + (int64_t)v0 * 64, bytes);
    };
    if (!RING && lane == 0)
      for (int u = 2; u < min(n_units, 2 + PD); ++u) l2_ahead(u);
    for (int u = 0; u < n_units; ++u) {
      const int s = u & 1, ws = u & (NST - 1);
      if (!RING && lane == 0 && u + 2 + PD < n_units) l2_ahead(u + 2 + PD);
      tc::mbar_wait(&mbar_full[s], (u >> 1) & 1);
      if (CONV && u >= 2) tc::mbar_wait(&mbar_tfree[s], ((u - 2) >> 1) & 1);  // every compute warp has read logits(u-2)
      tc::tc_fence_after();
      if (lane == 0) {
        const uint64_t bh = tc::desc_kmajor(tc::smem_u32(w_st + ws * 2 * BLK), 0), bl = tc::desc_kmajor(tc::smem_u32(w_st + ws * 2 * BLK + BLK), 0);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const uint32_t d = tmem_base + (uint32_t)(s * NB + nb) * 128;
          const uint64_t ah = tc::desc_kmajor(tc::smem_u32(h_blk + nb * 2 * BLK), 0), al = tc::desc_kmajor(tc::smem_u32(h_blk + nb * 2 * BLK + BLK), 0);
          bool acc = false;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint64_t a = pass == 2 ? al : ah, b = pass == 1 ? bl : bh;
#pragma unroll
            for (int k = 0; k < 4; ++k) { tc::mma_bf16(d, a + (uint64_t)(k * 2), b + (uint64_t)(k * 2), id_l, acc); acc = true; }
          }
        }
        tc::mma_commit(&mbar_done[s]);
      }
      __syncwarp();
    }
  } else if (loader) {
    // ---- TMA loader warp: (unit, head) tiles in order into the slot ring ----
    if (lane == 0) {
      while (ld_u < n_units) {
        tc::mbar_wait(&slot_free[ld_sl], (ld_round - 1) & 1);  // every compute warp has read the slot's previous tile
        loader_issue();
      }
    }
  } else if (conv) {
    // ---- converter warps (64 threads): slot ring -> bf16 hi/lo stage + bias tile, one unit ahead of the MMAs ----
    const int ctid = tid - (NT + 64);
    const float *bsrc = hp.b[first_head];
    int sl = 0;
    uint32_t round = 0;
    float bq[2] = {0.f, 0.f};
    auto bias_fetch_c = [&](int u) {
      const int v0 = (t_lo + u) * 128;
#pragma unroll
      for (int j = 0; j < 2; ++j) bq[j] = (v0 + ctid + 64 * j < Vloc) ? __ldg(bsrc + v0 + ctid + 64 * j) : 0.f;
    };
    if (n_units > 0) bias_fetch_c(0);
    for (int u = 0; u < n_units; ++u) {
      const int v0 = (t_lo + u) * 128;
      tc::mbar_wait(&slot_full[sl], round & 1);
      const float *src = stg + sl * 8192;
#pragma unroll 1
      for (int qd = 0; qd < 4; ++qd) {  // 32 rows per pass: 4 chunk pairs per thread in registers
        float4 xa[4], xb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = ctid + 64 * (qd * 4 + i), row = c >> 3, c8 = c & 7, sw = (c8 >> 2) & 1;
          float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
          if (v0 + row < Vloc) {
            x0 = *reinterpret_cast<const float4 *>(src + row * 64 + c8 * 8 + 4 * sw);
            x1 = *reinterpret_cast<const float4 *>(src + row * 64 + c8 * 8 + 4 * (sw ^ 1));
          }
          xa[i] = sw ? x1 : x0;
          xb[i] = sw ? x0 : x1;
        }
        if (qd == 0 && u > 0) tc::mbar_wait(&mbar_done[(u - 1) & 1], ((u - 1) >> 1) & 1);  // MMA(u-1) complete: the stage is free
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = ctid + 64 * (qd * 4 + i);
          tc::store_split8(w_st, w_st + BLK, c >> 3, c & 7, xa[i], xb[i]);
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&slot_free[sl]);
      if (++sl == n_slots) { sl = 0; ++round; }
#pragma unroll
      for (int j = 0; j < 2; ++j) bias_g[(u & 3) * 128 + ctid + 64 * j] = bq[j];
      tc::fence_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&mbar_full[u & 1]);
      if (u + 1 < n_units) bias_fetch_c(u + 1);
    }
  } else {
    // ---- compute warps ----
    // unit u = vocabulary tile t_lo + u.  fetch(): global -> registers (in flight across a whole pipeline step);
    // store(): registers -> (combined heads) -> bf16 hi/lo swizzled smem stage + bias tile.
    float4 fa[NH][TASKS], fb[NH][TASKS];
    float fbias[NH];
    const float wsc[3] = {(ARG && nh > 1) ? w0 : 1.f, ARG ? w1 : 0.f, ARG ? w2 : 0.f};
    auto fetch = [&](int u) {
      const int v0 = (t_lo + u) * 128;
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {
        fbias[hh] = 0.f;
#pragma unroll
        for (int i = 0; i < TASKS; ++i) { fa[hh][i] = make_float4(0.f, 0.f, 0.f, 0.f); fb[hh][i] = fa[hh][i]; }
        if (hh < nh) {
          const float *src = hp.w[first_head + hh];
#pragma unroll
          for (int i = 0; i < TASKS; ++i) {
            const int c = tid + NT * i, row = c >> 3, c8 = c & 7;
            if (v0 + row < Vloc) {
              const float4 *p = reinterpret_cast<const float4 *>(src + (int64_t)(v0 + row) * 64 + c8 * 8);
              fa[hh][i] = p[0];
              fb[hh][i] = p[1];
            }
          }
          if (tid < 128 && v0 + tid < Vloc) fbias[hh] = __ldg(hp.b[first_head + hh] + v0 + tid);
        }
      }
    };
    auto store = [&](int u) {
      const int s = u & 1;
      uint8_t *bh = w_st + s * 2 * BLK, *bl = bh + BLK;
#pragma unroll
      for (int i = 0; i < TASKS; ++i) {
        const int c = tid + NT * i;
        float4 a = fa[0][i], b = fb[0][i];
        if (ARG) {
          a.x *= wsc[0]; a.y *= wsc[0]; a.z *= wsc[0]; a.w *= wsc[0];
          b.x *= wsc[0]; b.y *= wsc[0]; b.z *= wsc[0]; b.w *= wsc[0];
#pragma unroll
          for (int hh = 1; hh < NH; ++hh) {  // heads beyond nh were fetched as zeros
            a.x = fmaf(wsc[hh], fa[hh][i].x, a.x); a.y = fmaf(wsc[hh], fa[hh][i].y, a.y);
            a.z = fmaf(wsc[hh], fa[hh][i].z, a.z); a.w = fmaf(wsc[hh], fa[hh][i].w, a.w);
            b.x = fmaf(wsc[hh], fb[hh][i].x, b.x); b.y = fmaf(wsc[hh], fb[hh][i].y, b.y);
            b.z = fmaf(wsc[hh], fb[hh][i].z, b.z); b.w = fmaf(wsc[hh], fb[hh][i].w, b.w);
          }
        }
        tc::store_split8(bh, bl, c >> 3, c & 7, a, b);
      }
      if (tid < 128) {
        float bsum = fbias[0] * wsc[0];
        if (ARG) {
#pragma unroll
          for (int hh = 1; hh < NH; ++hh) bsum = fmaf(wsc[hh], fbias[hh], bsum);
        }
        bias_g[(u & 3) * 128 + tid] = bsum;
      }
    };
    // RING: this thread's chunk pairs of unit u from the slot ring (heads combined on the fly) -> bf16 stage.
    // Bank-conflict-free slot reads: the 8 lanes of a row read alternating 16-byte halves of their 32-byte chunk.
    int g_sl = 0;
    uint32_t g_round = 0;
    float4 ga[TASKS], gb[TASKS];
    auto gather = [&](int u) {
      const int v0 = (t_lo + u) * 128;
      for (int hh = 0; hh < nh; ++hh) {
        tc::mbar_wait(&slot_full[g_sl], g_round & 1);
        STRACE(41);
        const float *src = stg + g_sl * 8192;
        const float w = wsc[hh];
#pragma unroll
        for (int i = 0; i < TASKS; ++i) {
          const int c = tid + NT * i, row = c >> 3, c8 = c & 7, sw = (c8 >> 2) & 1;
          float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
          if (v0 + row < Vloc) {
            x0 = *reinterpret_cast<const float4 *>(src + row * 64 + c8 * 8 + 4 * sw);
            x1 = *reinterpret_cast<const float4 *>(src + row * 64 + c8 * 8 + 4 * (sw ^ 1));
          }
          const float4 x = sw ? x1 : x0, y = sw ? x0 : x1;
          if (hh == 0) {
            ga[i] = make_float4(x.x * w, x.y * w, x.z * w, x.w * w);
            gb[i] = make_float4(y.x * w, y.y * w, y.z * w, y.w * w);
          } else {
            ga[i].x = fmaf(w, x.x, ga[i].x); ga[i].y = fmaf(w, x.y, ga[i].y); ga[i].z = fmaf(w, x.z, ga[i].z); ga[i].w = fmaf(w, x.w, ga[i].w);
            gb[i].x = fmaf(w, y.x, gb[i].x); gb[i].y = fmaf(w, y.y, gb[i].y); gb[i].z = fmaf(w, y.z, gb[i].z); gb[i].w = fmaf(w, y.w, gb[i].w);
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&slot_free[g_sl]);  // the values are in registers: the loader may refill the slot
        STRACE(42);
        if (++g_sl == n_slots) { g_sl = 0; ++g_round; }
      }
    };
    auto store_gathered = [&]() {  // registers -> the (single) bf16 stage; its last reader, MMA(u-1), must be complete
#pragma unroll
      for (int i = 0; i < TASKS; ++i) {
        const int c = tid + NT * i;
        tc::store_split8(w_st, w_st + BLK, c >> 3, c & 7, ga[i], gb[i]);
      }
    };
    float rb[3] = {0.f, 0.f, 0.f};  // RING: bias values of the next unit (tid < 128), one per head
    auto bias_fetch = [&](int u) {
      const int v0 = (t_lo + u) * 128;
      if (tid < 128) {
#pragma unroll
        for (int hh = 0; hh < 3; ++hh) rb[hh] = (hh < nh && v0 + tid < Vloc) ? __ldg(hp.b[first_head + hh] + v0 + tid) : 0.f;
      }
    };
    auto bias_store = [&](int u) {
      if (tid < 128) bias_g[(u & 3) * 128 + tid] = fmaf(wsc[2], rb[2], fmaf(wsc[1], rb[1], wsc[0] * rb[0]));
    };
    auto publish = [&](int u) {  // this warp's share of unit u is in shared memory
      tc::fence_async_smem();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&mbar_full[u & 1]);
    };

    // per-thread running state per block: row (b0 + nb*128 + q*32 + lane), tile columns [cq*CW, cq*CW + CW)
    float m_run[NB], s_run[NB], tgt[NB], av[NB], tau[NB];
    int ai[NB], cnt[NB], trow[NB];
    float r_v0[NB], r_v1[NB];  // top-k <= 2 lives in registers (no warm-up cost when a CTA only sees a few tiles)
    int r_i0[NB], r_i1[NB];
    const bool smallk = topk > 0 && topk <= 2;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      m_run[nb] = REC_NEG_INF; s_run[nb] = 0.f; tgt[nb] = REC_NEG_INF; av[nb] = REC_NEG_INF; tau[nb] = REC_NEG_INF;
      ai[nb] = 0x7fffffff; cnt[nb] = 0;
      r_v0[nb] = REC_NEG_INF; r_v1[nb] = REC_NEG_INF; r_i0[nb] = 0x7fffffff; r_i1[nb] = 0x7fffffff;
      const int row = b0 + nb * 128 + q * 32 + lane;
      trow[nb] = (!ARG && do_stats && target && row < B) ? (int)(target[row] - vocab_lo) : -1;
    }

    if (CONV) {
      // nothing to stage: the converter warps fill the bf16 stage
    } else if (RING) {
      if (n_units > 0) { bias_fetch(0); gather(0); store_gathered(); bias_store(0); publish(0); }
      if (n_units > 1) bias_fetch(1);
    } else {
      if (n_units > 0) { fetch(0); store(0); publish(0); }
      if (n_units > 1) fetch(1);
    }
    STRACE(21);

    for (int u = 0; u < n_units; ++u) {
      STRACE(1);
      if (CONV) {
        tc::mbar_wait(&mbar_done[u & 1], (u >> 1) & 1);
        tc::tc_fence_after();
      } else if (RING) {
        if (u + 1 < n_units) gather(u + 1);  // slot ring -> registers (heads combined)
        STRACE(3);
        tc::mbar_wait(&mbar_done[u & 1], (u >> 1) & 1);  // MMA(u) complete: logits(u) in TMEM, the bf16 stage is free
        tc::tc_fence_after();
        STRACE(5);
        if (u + 1 < n_units) {
          store_gathered();
          bias_store(u + 1);
          publish(u + 1);  // MMA(u+1) runs during epilogue(u)
          if (u + 2 < n_units) bias_fetch(u + 2);
        }
      } else {
        if (u + 1 < n_units) {
          // stage (u+1)&1 was last read by MMA(u-1), whose completion this warp observed in iteration u-1
          store(u + 1);
          STRACE(3);
          publish(u + 1);  // BEFORE the next fetch: the proxy fence would otherwise wait for those loads to land
          if (u + 2 < n_units) fetch(u + 2);  // lands while this step's MMA wait / epilogue run
        }
        STRACE(5);
        tc::mbar_wait(&mbar_done[u & 1], (u >> 1) & 1);
        tc::tc_fence_after();
      }
      STRACE(6);
      // ---- epilogue of unit u: NB blocks x CW columns in chunks of 32 ----
      const int v0 = (t_lo + u) * 128;
      if (ARG) {
        // greedy action: chunk maximum first, the index search only when it beats the running maximum
        constexpr int EC = RING ? 32 : 16;  // register-fetch path: the raw head tiles of unit u+2 are live in registers
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
#pragma unroll 1
          for (int ch = 0; ch < CW / EC; ++ch) {
            const int c_lo = v0 + cq * CW + ch * EC;
            float l[EC];
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(((u & 1) * NB + nb) * 128 + cq * CW + ch * EC);
            if (EC == 32) tc::tmem_ld32(ta, l); else tc::tmem_ld16(ta, l);
            const float *bgp = bias_g + (u & 3) * 128 + cq * CW + ch * EC;
#pragma unroll
            for (int j = 0; j < EC; j += 4) {
              float4 b4 = *reinterpret_cast<const float4 *>(bgp + j);
              l[j] += b4.x; l[j + 1] += b4.y; l[j + 2] += b4.z; l[j + 3] += b4.w;
            }
            if (c_lo + EC > Vloc) {
#pragma unroll
              for (int j = 0; j < EC; ++j) if (c_lo + j >= Vloc) l[j] = REC_NEG_INF;
            }
            float cm[4] = {l[0], l[1], l[2], l[3]};
#pragma unroll
            for (int j = 4; j < EC; ++j) cm[j & 3] = fmaxf(cm[j & 3], l[j]);
            if (fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3])) > av[nb]) {  // rare once the running maximum has warmed up
#pragma unroll
              for (int j = 0; j < EC; ++j)
                if (l[j] > av[nb]) { av[nb] = l[j]; ai[nb] = vocab_lo + c_lo + j; }  // ascending ids: strict > keeps the lowest id
            }
          }
        }
        continue;
      }
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
#pragma unroll 1
        for (int half = 0; half < CW / 32; ++half) {
          const int c_lo = v0 + cq * CW + half * 32;
          float l[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(((u & 1) * NB + nb) * 128 + cq * CW + half * 32), l);
          const float *bgp = bias_g + (u & 3) * 128 + cq * CW + half * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 b4 = *reinterpret_cast<const float4 *>(bgp + j);
            l[j] += b4.x; l[j + 1] += b4.y; l[j + 2] += b4.z; l[j + 3] += b4.w;
          }
          if (c_lo + 32 > Vloc) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (c_lo + j >= Vloc) l[j] = REC_NEG_INF;
          }
          float tmax = fmaxf(l[0], l[1]);  // chunk maximum (also the gate of the top-k fast path)
#pragma unroll
          for (int j = 2; j < 32; ++j) tmax = fmaxf(tmax, l[j]);
          if (!ARG && do_stats) {
            // (REC_NEG_INF is the finite -FLT_MAX: the clamp keeps nm * log2e finite for an all-masked chunk)
            const float nm = fmaxf(m_run[nb], tmax), nml = -fmaxf(nm, -1e30f) * LOG2E_F;
            float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 32; ++j) ps[j & 3] += tc::ex2_ftz(fmaf(l[j], LOG2E_F, nml));
            s_run[nb] = fmaf(s_run[nb], tc::ex2_ftz((m_run[nb] - nm) * LOG2E_F), (ps[0] + ps[1]) + (ps[2] + ps[3]));
            m_run[nb] = nm;
            if (trow[nb] >= c_lo && trow[nb] < c_lo + 32) {
              const int tj = trow[nb] - c_lo;
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j == tj) tgt[nb] = l[j];
            }
          }
          if (!ARG && topk > 0) {
            // fast path: nothing in this chunk beats the current k-th best (one compare against the chunk max);
            // the insertion code exists once, out of line, and walks the chunk from local memory
            const float cmax = tmax;
            if (smallk) {
              if (cmax > tau[nb]) {
                // 8-column groups behind their own maximum: with 32 independent rows per warp SOME lane enters here in
                // most chunks (k ln(n/k) entries per row), but rarely more than one or two groups are live
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                  float gm = fmaxf(fmaxf(l[g8 * 8], l[g8 * 8 + 1]), fmaxf(l[g8 * 8 + 2], l[g8 * 8 + 3]));
                  gm = fmaxf(gm, fmaxf(fmaxf(l[g8 * 8 + 4], l[g8 * 8 + 5]), fmaxf(l[g8 * 8 + 6], l[g8 * 8 + 7])));
                  if (gm > tau[nb]) {
#pragma unroll
                    for (int j = g8 * 8; j < g8 * 8 + 8; ++j) {
                      const float v = l[j];
                      const int id = vocab_lo + c_lo + j;
                      const bool g0 = v > r_v0[nb], g1 = v > r_v1[nb];
                      r_v1[nb] = g0 ? r_v0[nb] : (g1 ? v : r_v1[nb]);
                      r_i1[nb] = g0 ? r_i0[nb] : (g1 ? id : r_i1[nb]);
                      r_v0[nb] = g0 ? v : r_v0[nb];
                      r_i0[nb] = g0 ? id : r_i0[nb];
                    }
                    tau[nb] = topk == 1 ? r_v0[nb] : r_v1[nb];
                  }
                }
              }
            } else if (cmax > tau[nb]) {
              float *lv = tv + (size_t)nb * topk * NT + tid;
              int *li = ti + (size_t)nb * topk * NT + tid;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (l[j] > tau[nb]) topk_insert_one(l[j], vocab_lo + c_lo + j, lv, li, NT, topk, cnt[nb], tau[nb]);
            }
          }
        }
      }
      if (CONV) {  // logits(u) are consumed: the issuer may overwrite this TMEM buffer with unit u+2
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&mbar_tfree[u & 1]);
      }
    }

    STRACE(30);
    if (RING) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");  // every compute warp is done with the slot ring
    // ---- combine the CS column splits of every row inside the CTA, then publish ONE record per row ----
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      float *x = xch + ((nb * 128 + q * 32 + lane) * CS + cq) * 5;
      x[0] = m_run[nb]; x[1] = s_run[nb]; x[2] = tgt[nb]; x[3] = av[nb]; x[4] = __int_as_float(ai[nb]);
      if (!ARG && topk > 0) {  // pad the private list so that the merge below can read topk entries
        float *lv = tv + (size_t)nb * topk * NT + tid;
        int *li = ti + (size_t)nb * topk * NT + tid;
        if (smallk) {
          lv[0] = r_v0[nb]; li[0] = r_i0[nb];
          if (topk > 1) { lv[NT] = r_v1[nb]; li[NT] = r_i1[nb]; }
        } else {
          for (int k = cnt[nb]; k < topk; ++k) { lv[k * NT] = REC_NEG_INF; li[k * NT] = 0x7fffffff; }
        }
      }
    }
  }
  __syncthreads();
  if (tid < NT && cq == 0) {
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int r = q * 32 + lane, row = b0 + nb * 128 + r;
      if (row >= B) continue;
      const float *x = xch + ((nb * 128 + r) * CS) * 5;
      float m = REC_NEG_INF, tg = REC_NEG_INF, bv = REC_NEG_INF, bv2 = REC_NEG_INF;
      int bi = 0x7fffffff, bi2 = 0x7fffffff;
#pragma unroll
      for (int c = 0; c < CS; ++c) {
        m = fmaxf(m, x[c * 5]);
        tg = fmaxf(tg, x[c * 5 + 2]);
        const float v = x[c * 5 + 3];
        const int i = __float_as_int(x[c * 5 + 4]);
        if (better(v, i, bv, bi)) { bv2 = bv; bi2 = bi; bv = v; bi = i; }
        else if (better(v, i, bv2, bi2)) { bv2 = v; bi2 = i; }
      }
      float ssum = 0.f;
#pragma unroll
      for (int c = 0; c < CS; ++c) if (x[c * 5 + 1] > 0.f) ssum += x[c * 5 + 1] * __expf(x[c * 5] - m);
      float *o = part + ((int64_t)sp * B + row) * part_stride;
      o[0] = m; o[1] = ssum; o[2] = tg; o[3] = bv; o[4] = __int_as_float(bi);
      if (ARG) {
        // greedy action: the scores are bf16x3 products -- publish the two best of the CS per-thread maxima as
        // candidates; the merge re-scores the best REC_ARG_CAND of all records in fp32 (common.cuh: rec_kpub)
        o[PART_TOPK_OFF] = bv; o[PART_TOPK_OFF + REC_MAX_TOPK] = __int_as_float(bi);
        o[PART_TOPK_OFF + 1] = bv2; o[PART_TOPK_OFF + REC_MAX_TOPK + 1] = __int_as_float(bi2);
      }
      if (!ARG && topk > 0) {
        // CS-way merge of the sorted private lists (ids of different splits are disjoint); rec_kpub(topk, CS) >= topk
        // entries leave the CTA so that the fp32 re-score of the final merge has a margin of candidates
        const int kpub = rec_kpub(topk, CS);
        int pos[CS];
#pragma unroll
        for (int c = 0; c < CS; ++c) pos[c] = 0;
        for (int k = 0; k < kpub; ++k) {
          float cv = REC_NEG_INF;
          int ci = 0x7fffffff, cc = 0;
#pragma unroll
          for (int c = 0; c < CS; ++c) {
            if (pos[c] < topk) {
              const int t2 = (c * 4 + q) * 32 + lane;  // thread id of (q, cq = c, lane)
              const float v = tv[((size_t)nb * topk + pos[c]) * NT + t2];
              const int i = ti[((size_t)nb * topk + pos[c]) * NT + t2];
              if (better(v, i, cv, ci)) { cv = v; ci = i; cc = c; }
            }
          }
#pragma unroll
          for (int c = 0; c < CS; ++c) if (c == cc) ++pos[c];
          o[PART_TOPK_OFF + k] = cv;
          o[PART_TOPK_OFF + REC_MAX_TOPK + k] = __int_as_float(ci);
        }
      }
    }
  }
  STRACE(31);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, NB * 256);
}

// ------------------------------------------------------------------------------------------------
#include <stdlib.h>
// debug: REC_TRACE_SEL=1 routes the phase trace (rec_debug_set_trace) to the statistics kernel instead of the
// backward kernel
static int trace_sel() {
  static int sel = -1;
  if (sel < 0) { const char *v = getenv("REC_TRACE_SEL"); sel = v ? atoi(v) : 0; }
  return sel;
}

static TcHeadPtrs tc_head_ptrs(const rec_engine *e, int net_id) {
  TcHeadPtrs hp;
  for (int i = 0; i < REC_MAX_HEADS; ++i) { hp.w[i] = e->nets[net_id].p.head_w[i]; hp.b[i] = e->nets[net_id].p.head_b[i]; }
  return hp;
}

bool tc_heads_supported(const rec_engine *e) { return e->D == 64 && e->use_tc; }

// Same contract as launch_head_stats (heads.cu); *n_split_out counts records per row (one per CTA column).
template <int NB, int NT, bool ARG, bool RING>
static int launch_stats_tc_kernel(rec_engine *e, const HeadStatsArgs &a, dim3 grid, size_t smem, int n_tiles, int topk, int n_slots) {
  static size_t attr_smem[REC_MAX_DEVICES] = {};
  if (smem > attr_smem[e->dev]) {
    REC_CUDA(e, cudaFuncSetAttribute(head_stats_tc_kernel<NB, NT, ARG, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[e->dev] = smem;
  }
  head_stats_tc_kernel<NB, NT, ARG, RING><<<grid, NT + (RING ? (ARG ? 64 : 128) : 32), smem, e->stream>>>(
      tc_head_ptrs(e, a.net_id), a.h, a.B, e->Vloc, e->cfg.vocab_lo, n_tiles, ARG ? 0 : a.do_stats, ARG ? 1 + a.arg_shift : a.stats_head,
      ARG ? a.n_arg : 1, a.w[0], a.w[1], a.w[2], a.target, topk, e->part, e->part_stride, trace_sel() == (ARG ? 2 : 1) ? e->trace : nullptr,
      n_slots);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

// Same contract as launch_head_stats (heads.cu); *n_split_out counts records per row (one per CTA column).
// RING_OK: the 512-thread (training) variants stream the weights through the TMA slot ring when at least one
// 32 KB slot fits next to the top-k lists; the evaluation variants keep the register-fetch path.
template <int NB, int NT, bool ARG, bool RING_OK>
static int launch_stats_tc_variant(rec_engine *e, const HeadStatsArgs &a, int *n_split_out) {
  const int n_tiles = cdiv(e->Vloc, 128), ng = cdiv(a.B, 128 * NB);
  int n_split = e->sm_count / ng;  // one CTA per SM, a single wave
  if (n_split > n_tiles) n_split = n_tiles;
  if (n_split < 1) n_split = 1;
  int per = cdiv(n_tiles, n_split);
  n_split = cdiv(n_tiles, per);
  const int topk = ARG ? 0 : a.topk;
  const size_t xch_bytes = (size_t)NB * 128 * (NT / 128) * 5 * 4, lists = (size_t)NB * topk * NT * 8;
  const size_t base_ring = 1024 + (size_t)(NB * 2 + 2) * BLK + 2048 + lists + 128;  // single bf16 stage, xch inside the slots
  const size_t smem_plain = 1024 + (size_t)(NB * 2 + 4) * BLK + 2048 + lists + 128 + xch_bytes;
  if (smem_plain > 227 * 1024) REC_FAIL(e, REC_EINVAL, "head statistics kernel needs %zu B of shared memory (top-k %d)", smem_plain, topk);
  static int no_ring = -1;
  if (no_ring < 0) { const char *v = getenv("REC_NO_RING"); no_ring = v ? atoi(v) : 0; }
  int n_slots = RING_OK && !no_ring ? (int)((227 * 1024 - base_ring) / 32768) : 0;
  if (n_slots > 4) n_slots = 4;
  dim3 grid(n_split, ng);
  *n_split_out = n_split;
  e->st_approx = true;
  e->st_kpub = topk > 0 ? rec_kpub(topk, NT / 128) : 0;
  e->st_apub = ARG ? 2 : 0;
  if (RING_OK && n_slots > 0)
    return launch_stats_tc_kernel<NB, NT, ARG, RING_OK>(e, a, grid, base_ring + (size_t)n_slots * 32768, n_tiles, topk, n_slots);
  return launch_stats_tc_kernel<NB, NT, ARG, false>(e, a, grid, smem_plain, n_tiles, topk, 0);
}

// Same contract as launch_head_stats (heads.cu); *n_split_out = records per row.
int launch_head_stats_tc(rec_engine *e, const HeadStatsArgs &a, int *n_split_out) {
  if (a.n_arg > 0) {
    if (a.n_arg > 3) REC_FAIL(e, REC_EINVAL, "greedy-action pass supports at most 3 Q heads (got %d)", a.n_arg);
    if (a.B <= 128) return launch_stats_tc_variant<1, 512, true, true>(e, a, n_split_out);
    return launch_stats_tc_variant<2, 512, true, true>(e, a, n_split_out);
  }
  if (a.B <= 128) return launch_stats_tc_variant<1, 512, false, true>(e, a, n_split_out);
  if (a.topk <= 8) return launch_stats_tc_variant<2, 512, false, true>(e, a, n_split_out);   // training: 16 warps, 32 columns/thread
  static int eval_wide = -1;
  if (eval_wide < 0) { const char *v = getenv("REC_EVAL_WIDE"); eval_wide = v ? atoi(v) : 1; }
  // evaluation (8 < k <= 20): the per-thread top-k lists dominate shared memory.  For moderate (batch x catalogue) sizes
  // one 128-session block with 16 warps (32 logits per thread and tile) beats two blocks with 8 warps.
  // (measured: 70 852 items / B = 2000: 0.83 vs 0.97 ms per batch; 1 M items / B = 5000: 8.3 vs 6.7 ms -- with many
  // column groups the per-tile weight conversion, repeated by every group, outweighs the faster epilogue)
  if (a.topk <= 20 && (eval_wide == 2 || (eval_wide && (int64_t)cdiv(a.B, 128) * cdiv(e->Vloc, 128) < 200000)))
    return launch_stats_tc_variant<1, 512, false, true>(e, a, n_split_out);
  if (a.topk <= 20) return launch_stats_tc_variant<2, 256, false, true>(e, a, n_split_out);
  return launch_stats_tc_variant<1, 256, false, true>(e, a, n_split_out);
}

// ================================================================================================
// Supervised head backward + Adam on the tensor cores (D = 64).  Per 128-row vocabulary tile, with W, m, v read ONCE:
//   for each 128-session block bb (batches beyond 256 sessions: chunks of two blocks):
//     L   = h[bb] . W^T                      tcgen05.mma  K-major x K-major            -> TMEM[0,128) / [128,256)
//     dl  = (exp(L + b - lse) - onehot)/B    epilogue: TMEM -> regs -> bf16 hi/lo -> smem (swizzled)
//     dW += dl^T . h[bb]                     tcgen05.mma  MN-major x MN-major          -> TMEM[256,320)
//     db += dl^T . 1                         tcgen05.mma  MN-major x K-major (ones)    -> TMEM[320,336)
//     dh[bb] += dl . W                       tcgen05.mma  K-major x MN-major           -> TMEM[384+64bb, ...)
//   dW, db: TMEM -> registers / per-warp smem transpose -> Adam in coalesced 64-byte pieces -> W, m, v written once.
// For B <= 256 dh stays in TMEM across ALL tiles of the CTA and is written once at the end (one slice per CTA).
// The three operand roles of h, W and dl use the SAME shared-memory bytes (see tc.cuh).
// ================================================================================================
struct TcTrainPtrs {
  float *w, *wm, *wv, *b, *bm, *bv;
};

__device__ __forceinline__ void adam_f(float &p, float &m, float &v, float g, float b1, float b2, float eps, float step_size,
                                       float inv_bc2_sqrt) {
  adam_elem(p, m, v, g, b1, b2, eps, step_size, inv_bc2_sqrt);
}

// h [B, 64] fp32 -> packed bf16 hi/lo image: block bb (128 rows) = [hi 16 KB | lo 16 KB], swizzled exactly as
// the kernels want it in shared memory, so that (re)loading a block is a raw TMA bulk copy.
__global__ void __launch_bounds__(256) h_prepack_kernel(const float *__restrict__ h, int B, uint8_t *__restrict__ out) {
  const int bb = blockIdx.x, tid = threadIdx.x;
  uint8_t *hi = out + (size_t)bb * 2 * BLK, *lo = hi + BLK;
  stage_rows64(hi, lo, h, bb * 128, B, 1.f, tid);
}

// ================================================================================================
// head_bwd_adam_tc2_kernel -- warp-specialised into four roles that only meet at mbarriers (the first, single-role version
// of this kernel -- every warp did every phase behind block barriers -- ran at 31 % of the HBM roofline at 1 M items):
//   16 compute warps : wait W(t) staged -> convert to bf16 hi/lo -> arrive WREADY
//                      wait L[bb] -> dlogits in registers (ONE ex2 per logit) -> (bb>0: wait G) -> smem -> arrive DL
//                      wait G (the W/dl operands are free again)
//   1 issuer warp    : TMA prefetch of W(t+1), L2 prefetch of the tile's Adam state, logits(bb=0,1) MMAs,
//                      dW,db,dh(bb) MMAs -> commit G (and GT after the tile's last block)
//   8 Adam warps     : (TMEM lane quarter = 32 vocabulary rows) x (32 columns): wait GT -> dW/db TMEM -> registers /
//                      per-warp smem transpose -> arrive DWFREE -> coalesced Adam on p, m, v (L2 hits).
//                      They run one tile behind the others: the update of tile t overlaps the GEMMs of tile t+1.
// Why: tcgen05.mma issue blocks the issuing thread for as long as the tensor pipe is busy (the gradient GEMMs with
// N = 64 are shared-memory-bandwidth bound, ~48 cycles per MMA), and the Adam stores are bound by the SM's
// store port (one 32-byte sector per cycle: 3072 cycles per tile) -- neither may sit on the critical path.
// ================================================================================================
#define BWD2_COMPUTE 512
#define BWD2_THREADS (BWD2_COMPUTE + 32 + 256)

// CHUNKED = more than 256 sessions (a compile-time split: the resident instantiation -- every single-GPU configuration of the
// reference -- carries none of the chunk bookkeeping in its registers).
template <bool CHUNKED>
__global__ void __launch_bounds__(BWD2_THREADS, 1) head_bwd_adam_tc2_kernel(TcTrainPtrs hp, const uint8_t *__restrict__ hpack,
                                                                           const int64_t *__restrict__ target,
                                                                           const float *__restrict__ row_stats, int B, int Vloc,
                                                                           int vocab_lo, int n_tiles, float inv_B,
                                                                           float *__restrict__ dh_part, float b1, float b2,
                                                                           float eps, float step_size, float inv_bc2_sqrt,
                                                                           long long *__restrict__ trace,
                                                                           const float *__restrict__ sc,
                                                                           const float *__restrict__ extra) {
  // extra (optional, [B]): a per-row gradient added to dlogits at the target column (SARM: head 0 is a Q head too)
  if (sc) { step_size = sc[0]; inv_bc2_sqrt = sc[1]; }
  extern __shared__ uint8_t raw[];
  uint8_t *sm = raw + ((1024u - (tc::smem_u32(raw) & 1023u)) & 1023u);  // 1024-aligned, provably shared
  int tr_n = 0;
  uint8_t *h_blk = sm;                 // 256 sessions: [bb][hi|lo] x BLK      (64 KB)
  uint8_t *w_hi = sm + 4 * BLK, *w_lo = sm + 5 * BLK;            // (32 KB)
  uint8_t *dl_hi = sm + 6 * BLK, *dl_lo = sm + 8 * BLK;          // each 2 blocks (v halves)  (64 KB)
  float *w_stage = reinterpret_cast<float *>(sm + 10 * BLK);     // fp32 tile [128][64], TMA destination (32 KB)
  uint8_t *ones = sm + 12 * BLK;                                  // 4 KB of bf16 1.0
  float *bias_s = reinterpret_cast<float *>(ones + 4096);        // [128]
  float *lse_s = bias_s + 128;                                    // [2][256] per-chunk log-sum-exp (double-buffered)
  int *tgt_s = reinterpret_cast<int *>(lse_s + 512);             // [2][256] target column relative to vocab_lo
  float *xpose = reinterpret_cast<float *>(tgt_s + 512);         // [8 warps][32 rows][20] Adam warps' transpose staging
  enum { MB_L0 = 0, MB_L1, MB_W, MB_H, MB_G, MB_GT, MB_WREADY, MB_DL, MB_DWFREE, MB_N };
  __shared__ uint64_t mbar[MB_N];
  __shared__ uint32_t tmem_base_s;
  constexpr uint32_t T_L = 0 /* 2 x 128 */, T_DW = 256, T_DB = 320, T_DH = 384 /* 2 x 64 */;
  constexpr int NT = BWD2_COMPUTE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = warp == NT / 32, adam = warp > NT / 32;
  const int q = warp & 3, cq = (warp >> 2) & 3;  // TMEM lane quarter; column quarter (4 warps share a lane quarter)
  // batches beyond 256 sessions run in chunks of 256 per tile: the h chunk is re-read (TMA, from L2) per (tile, chunk),
  // dW/db accumulate over the chunks in TMEM, dh leaves TMEM per (tile, chunk) into this CTA's slice
  const int n_chunks = CHUNKED ? (B + 255) / 256 : 1;
  constexpr bool resident = !CHUNKED;  // dh stays in TMEM across all tiles of the CTA
  const int nbb0 = min(2, (B + 127) / 128);  // 128-session blocks of chunk 0
  const float log2_inv_B = __log2f(inv_B);

  if (issuer && lane == 0) {
    for (int i = 0; i < MB_N; ++i)
      tc::mbar_init(&mbar[i], (i == MB_WREADY || i == MB_DL) ? NT / 32 : (i == MB_DWFREE ? 8 : 1));
    tc::fence_barrier_init();
    if ((int)blockIdx.x < n_tiles) {  // first weight tile + the packed h image: in flight during the prologue
      const uint32_t bytes = (uint32_t)min(128, Vloc - (int)blockIdx.x * 128) * 256u;
      tc::mbar_expect_tx(&mbar[MB_W], bytes);
      tc::bulk_g2s(w_stage, hp.w + (int64_t)blockIdx.x * 128 * 64, bytes, &mbar[MB_W]);
      tc::mbar_expect_tx(&mbar[MB_H], (uint32_t)nbb0 * 2 * BLK);
      tc::bulk_g2s(h_blk, hpack, (uint32_t)nbb0 * 2 * BLK, &mbar[MB_H]);
    }
  }
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  for (int i = tid; i < 4096 / 4; i += BWD2_THREADS) reinterpret_cast<uint32_t *>(ones)[i] = 0x3F803F80u;
  if (tid < 256 && resident) {
    lse_s[tid] = tid < B ? row_stats[(int64_t)tid * 8] : 0.f;
    tgt_s[tid] = tid < B ? (int)(target[tid] - vocab_lo) : -1;
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (issuer) {
    const uint32_t s_h = tc::smem_u32(h_blk), s_wh = tc::smem_u32(w_hi), s_wl = tc::smem_u32(w_lo);
    const uint32_t s_dh = tc::smem_u32(dl_hi), s_dl = tc::smem_u32(dl_lo), s_one = tc::smem_u32(ones);
    auto prefetch_w = [&](int t) {  // TMA bulk copy of the fp32 tile
      const uint32_t bytes = (uint32_t)min(128, Vloc - t * 128) * 256u;
      tc::mbar_expect_tx(&mbar[MB_W], bytes);
      tc::bulk_g2s(w_stage, hp.w + (int64_t)t * 128 * 64, bytes, &mbar[MB_W]);
    };
    const uint64_t d_w_k_hi = tc::desc_kmajor(s_wh, 0), d_w_k_lo = tc::desc_kmajor(s_wl, 0);
    const uint64_t d_dl_mn_hi = tc::desc_mnmajor(s_dh, 0, BLK), d_dl_mn_lo = tc::desc_mnmajor(s_dl, 0, BLK);
    const uint64_t d_dl_k_hi = tc::desc_kmajor(s_dh, 0), d_dl_k_lo = tc::desc_kmajor(s_dl, 0);
    const uint64_t d_w_mn_hi = tc::desc_mnmajor(s_wh, 0, BLK), d_w_mn_lo = tc::desc_mnmajor(s_wl, 0, BLK);
    const uint64_t d_one = tc::desc_kmajor(s_one, 0);
    auto issue_logits = [&](int bb) {
      const uint32_t id = tc::instr_desc(128, 128, 0, 0);
      const uint64_t d_h_hi = tc::desc_kmajor(s_h + bb * 2 * BLK, 0), d_h_lo = tc::desc_kmajor(s_h + bb * 2 * BLK + BLK, 0);
      bool acc = false;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint64_t a = pass == 2 ? d_h_lo : d_h_hi, b = pass == 1 ? d_w_k_lo : d_w_k_hi;
#pragma unroll
        for (int k = 0; k < 4; ++k) { tc::mma_bf16(tmem + T_L + bb * 128, a + (uint64_t)(k * 2), b + (uint64_t)(k * 2), id, acc); acc = true; }
      }
      tc::mma_commit(&mbar[MB_L0 + bb]);
    };
    auto issue_dW = [&](int bb, bool first_of_tile) {  // dW += dl^T . h[bb]  (A: dl MN-major, B: h MN-major, K = batch)
      const uint32_t id = tc::instr_desc(128, 64, 1, 1);
      const uint64_t d_h_hi = tc::desc_mnmajor(s_h + bb * 2 * BLK, 0, BLK), d_h_lo = tc::desc_mnmajor(s_h + bb * 2 * BLK + BLK, 0, BLK);
      bool acc = !first_of_tile;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint64_t a = pass == 2 ? d_dl_mn_lo : d_dl_mn_hi, b = pass == 1 ? d_h_lo : d_h_hi;
#pragma unroll
        for (int k = 0; k < 8; ++k) { tc::mma_bf16(tmem + T_DW, a + (uint64_t)(k * 128), b + (uint64_t)(k * 128), id, acc); acc = true; }
      }
    };
    auto issue_db = [&](bool first_of_tile) {  // db += dl^T . 1  (B: 16 rows of ones, K-major, two K blocks of 2 KB)
      const uint32_t id = tc::instr_desc(128, 16, 1, 0);
      bool acc = !first_of_tile;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const uint64_t a = pass == 1 ? d_dl_mn_lo : d_dl_mn_hi;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          tc::mma_bf16(tmem + T_DB, a + (uint64_t)(k * 128), d_one + (uint64_t)((k >> 2) * 128 + (k & 3) * 2), id, acc);
          acc = true;
        }
      }
    };
    auto issue_dh = [&](int bb, bool acc_dh) {  // dh[bb] += dl . W  (A: dl K-major (K = v, 2 blocks), B: W MN-major)
      const uint32_t id = tc::instr_desc(128, 64, 0, 1);
      bool acc = acc_dh;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint64_t a = pass == 2 ? d_dl_k_lo : d_dl_k_hi, b = pass == 1 ? d_w_mn_lo : d_w_mn_hi;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          tc::mma_bf16(tmem + T_DH + bb * 64, a + (uint64_t)((k >> 2) * (BLK >> 4) + (k & 3) * 2), b + (uint64_t)(k * 128), id, acc);
          acc = true;
        }
      }
    };
    uint32_t ph_dl = 0, ph_ready = 0, ph_h = 0;
    int k = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++k) {
      if (lane == 0) {  // Adam state of this tile -> L2 while the tile's GEMMs run
        const uint32_t bytes = (uint32_t)min(128, Vloc - t * 128) * 256u;
        tc::l2_prefetch(hp.wm + (int64_t)t * 128 * 64, bytes);
        tc::l2_prefetch(hp.wv + (int64_t)t * 128 * 64, bytes);
      }
      for (int c = 0; c < n_chunks; ++c) {
        const int nbb = min(2, (B - c * 256 + 127) / 128);
        // every compute warp is done with the previous (tile, chunk): W converted (c == 0), lse/targets staged, dh read
        tc::mbar_wait(&mbar[MB_WREADY], ph_ready);
        ph_ready ^= 1;
        if (k == 0 && c == 0) {
          tc::mbar_wait(&mbar[MB_H], ph_h);  // requested before the prologue barrier
          ph_h ^= 1;
        } else if (!resident) {  // h chunk (all MMAs that read the previous one are complete)
          if (lane == 0) {
            tc::mbar_expect_tx(&mbar[MB_H], (uint32_t)nbb * 2 * BLK);
            tc::bulk_g2s(h_blk, hpack + (size_t)c * 4 * BLK, (uint32_t)nbb * 2 * BLK, &mbar[MB_H]);
          }
          tc::mbar_wait(&mbar[MB_H], ph_h);
          ph_h ^= 1;
        }
        tc::tc_fence_after();
        if (lane == 0) {
          if (c == 0 && t + (int)gridDim.x < n_tiles) prefetch_w(t + gridDim.x);  // every compute warp has read the staging buffer
          for (int bb = 0; bb < nbb; ++bb) issue_logits(bb);
        }
        __syncwarp();
        for (int bb = 0; bb < nbb; ++bb) {
          tc::mbar_wait(&mbar[MB_DL], ph_dl);
          ph_dl ^= 1;
          if (c == 0 && bb == 0 && k > 0) tc::mbar_wait(&mbar[MB_DWFREE], (k - 1) & 1);  // the Adam warps have read dW/db of the previous tile
          tc::tc_fence_after();
          if (lane == 0) {
            issue_dW(bb, c == 0 && bb == 0);
            issue_db(c == 0 && bb == 0);
            issue_dh(bb, resident && k > 0);
            tc::mma_commit(&mbar[MB_G]);
            if (c == n_chunks - 1 && bb == nbb - 1) tc::mma_commit(&mbar[MB_GT]);
          }
          __syncwarp();
        }
      }
    }
  } else if (adam) {
    // ---- Adam warps: warp (aq, ch) owns vocabulary rows [32 aq, 32 aq + 32) x columns [32 ch, 32 ch + 32) of every tile.
    // Both 16-column halves leave TMEM right after the tile's gradients complete (the first half waits in registers,
    // the second in the warp's smem staging), so the next tile's dW GEMM is never held up by the update.
    const int aw = warp - (NT / 32 + 1);  // 0..7
    const int aq = warp & 3;              // TMEM lane quarter this warp may read (hardware: warp index % 4)
    const int c0 = (aw >> 2) * 32;
    float *xp = xpose + aw * (32 * 20);
    const int rsub = lane >> 2, csub = (lane & 3) * 4;  // staging read-out: 8 rows x 64 bytes per instruction
    int k = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++k) {
      const int v0 = t * 128 + aq * 32;
      const int nrows = min(32, Vloc - v0);  // may be <= 0 in the last tile
      float bp = 0.f, bm = 0.f, bv = 0.f;
      if (c0 == 0 && lane < nrows) { bp = hp.b[v0 + lane]; bm = hp.bm[v0 + lane]; bv = hp.bv[v0 + lane]; }
      tc::mbar_wait(&mbar[MB_GT], k & 1);
      tc::tc_fence_after();
      float4 G[4];
      {
        float g[16];
        tc::tmem_ld16(tmem + ((uint32_t)(aq * 32) << 16) + T_DW + (uint32_t)c0, g);
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4 *>(xp + lane * 20 + j) = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) G[i] = *reinterpret_cast<const float4 *>(xp + (i * 8 + rsub) * 20 + csub);
      __syncwarp();
      float dbias = 0.f;
      {
        float g[16];
        tc::tmem_ld16(tmem + ((uint32_t)(aq * 32) << 16) + T_DW + (uint32_t)(c0 + 16), g);
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4 *>(xp + lane * 20 + j) = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
        if (c0 == 0) {
          tc::tmem_ld16(tmem + ((uint32_t)(aq * 32) << 16) + T_DB, g);
          dbias = g[0];
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&mbar[MB_DWFREE]);  // this warp's share of dW/db is out of TMEM
#pragma unroll 1
      for (int rd = 0; rd < 2; ++rd) {
        if (rd == 1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) G[i] = *reinterpret_cast<const float4 *>(xp + (i * 8 + rsub) * 20 + csub);
        }
#pragma unroll
        for (int ip = 0; ip < 4; ip += 2) {  // two float4 per lane and array in flight (register budget: 72)
          float4 P[2], M[2], U[2];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int row = (ip + i) * 8 + rsub;
            const int64_t off = (int64_t)(v0 + max(0, min(row, nrows - 1))) * 64 + c0 + rd * 16 + csub;
            if (nrows > 0) {
              P[i] = *reinterpret_cast<const float4 *>(hp.w + off);
              M[i] = *reinterpret_cast<const float4 *>(hp.wm + off);
              U[i] = *reinterpret_cast<const float4 *>(hp.wv + off);
            }
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int row = (ip + i) * 8 + rsub;
            const float4 g = G[ip + i];
            adam_f(P[i].x, M[i].x, U[i].x, g.x, b1, b2, eps, step_size, inv_bc2_sqrt);
            adam_f(P[i].y, M[i].y, U[i].y, g.y, b1, b2, eps, step_size, inv_bc2_sqrt);
            adam_f(P[i].z, M[i].z, U[i].z, g.z, b1, b2, eps, step_size, inv_bc2_sqrt);
            adam_f(P[i].w, M[i].w, U[i].w, g.w, b1, b2, eps, step_size, inv_bc2_sqrt);
            if (row < nrows) {
              const int64_t off = (int64_t)(v0 + row) * 64 + c0 + rd * 16 + csub;
              *reinterpret_cast<float4 *>(hp.w + off) = P[i];
              *reinterpret_cast<float4 *>(hp.wm + off) = M[i];
              *reinterpret_cast<float4 *>(hp.wv + off) = U[i];
            }
          }
        }
      }
      if (c0 == 0 && lane < nrows) {
        adam_f(bp, bm, bv, dbias, b1, b2, eps, step_size, inv_bc2_sqrt);
        hp.b[v0 + lane] = bp; hp.bm[v0 + lane] = bm; hp.bv[v0 + lane] = bv;
      }
      __syncwarp();  // xp is rewritten by the next tile
    }
  } else {
#define TRACE2(tag) do { if (trace && blockIdx.x == 0 && threadIdx.x == 0 && tr_n < 120) { asm volatile("" ::: "memory"); trace[2 * tr_n] = (tag); trace[2 * tr_n + 1] = clock64(); ++tr_n; asm volatile("" ::: "memory"); } } while (0)
    auto publish = [&](int which) {
      tc::fence_async_smem();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&mbar[which]);
    };
    uint32_t phG = 0, phL[2] = {0, 0};
    float bias_p = 0.f;
    if (tid < 128 && (int)blockIdx.x < n_tiles && (int)blockIdx.x * 128 + tid < Vloc) bias_p = hp.b[blockIdx.x * 128 + tid];
    // Chunked mode (more than 256 sessions): dh of a (tile, chunk) leaves TMEM into this CTA's slice (first tile stores,
    // later tiles add).  The read-modify-write (64 KB each way, one 256-byte row per lane) used to sit between the
    // chunk's last gradient MMA and the next chunk's first logits MMA; it now runs BEHIND this warp's next
    // publish(MB_WREADY), i.e. while the issuer loads the next h chunk and issues its logits.  Safe: the next dh MMA
    // is issued only after every compute warp arrived at MB_DL, which each does after its own read-out.
    int pend_r0 = -1, pend_nbb = 0, pend_k = 0;
    float nx_lse = 0.f;  // log-sum-exp / target of this thread's row in the next chunk, loaded one chunk ahead
    int nx_tgt = -1;
    bool nx_valid = false;
    auto flush_dh = [&]() {
      if (resident || pend_r0 < 0) return;
      float *slice = dh_part + (int64_t)blockIdx.x * B * 64;
      for (int bb = 0; bb < pend_nbb; ++bb) {
        const int row_base = pend_r0 + bb * 128 + q * 32;  // the warp's 32 rows; this lane holds 16 columns of row_base + lane
        float g[16];
        tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + T_DH + (uint32_t)(bb * 64 + cq * 16), g);
        // lane = row means every lane of a load / store touches another 128-byte line (32 L1 wavefronts per instruction:
        // ~8 k LSU cycles per chunk).  Quad transpose through shuffles instead: instruction i covers rows 8i .. 8i + 7,
        // the four lanes of a quad write the four float4 of ONE row = 64 contiguous bytes (8 lines per instruction).
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int src = 8 * i + (lane >> 2);
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x = __shfl_sync(0xffffffffu, g[4 * j], src), y = __shfl_sync(0xffffffffu, g[4 * j + 1], src);
            const float z = __shfl_sync(0xffffffffu, g[4 * j + 2], src), w = __shfl_sync(0xffffffffu, g[4 * j + 3], src);
            if ((lane & 3) == j) o = make_float4(x, y, z, w);
          }
          const int row = row_base + src;
          if (row < B) {
            float4 *dst = reinterpret_cast<float4 *>(slice + (int64_t)row * 64 + cq * 16) + (lane & 3);
            if (pend_k > 0) { const float4 p = *dst; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
            *dst = o;
          }
        }
      }
      pend_r0 = -1;
    };
    int k = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++k) {
      const int v0 = t * 128;
      // ---- W tile: staged fp32 (TMA) -> bf16 hi/lo in smem ---------------------------------------------
      // (converting it into registers during the previous tile's last gradient GEMMs was measured SLOWER: 475 -> 518 us)
      TRACE2(1);
      tc::mbar_wait(&mbar[MB_W], k & 1);
      TRACE2(2);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = tid + NT * i, row = c >> 3, c8 = c & 7;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (v0 + row < Vloc) {
          const float4 *p = reinterpret_cast<const float4 *>(w_stage + row * 64 + c8 * 8);
          a = p[0];
          b = p[1];
        }
        tc::store_split8(w_hi, w_lo, row, c8, a, b);
      }
      if (tid < 128) bias_s[tid] = bias_p;
      for (int c = 0; c < n_chunks; ++c) {
        const int r0 = c * 256, nbb = min(2, (B - r0 + 127) / 128);
        float *lse_c = lse_s + (resident ? 0 : (c & 1) * 256);
        int *tgt_c = tgt_s + (resident ? 0 : (c & 1) * 256);
        if (!resident && tid < 256) {  // this chunk's rows (the other buffer may still be read by a slower warp)
          if (!nx_valid) {  // very first chunk of the CTA; afterwards the values were requested one chunk ahead
            const int row = r0 + tid;
            nx_lse = row < B ? row_stats[(int64_t)row * 8] : 0.f;
            nx_tgt = row < B ? (int)(target[row] - vocab_lo) : -1;
          }
          lse_c[tid] = nx_lse;
          tgt_c[tid] = nx_tgt;
        }
        publish(MB_WREADY);
        flush_dh();  // dh of the previous (tile, chunk) leaves TMEM while the issuer loads the next h chunk / issues its logits
        if (!resident && tid < 256) {  // rows of the NEXT (tile, chunk): in flight during this chunk's epilogues
          const int row = (c + 1 < n_chunks ? c + 1 : 0) * 256 + tid;
          nx_lse = row < B ? row_stats[(int64_t)row * 8] : 0.f;
          nx_tgt = row < B ? (int)(target[row] - vocab_lo) : -1;
          nx_valid = true;
        }
        if (c == 0 && tid < 128) {  // logits bias of the next tile (its Adam update belongs to a later tile: no hazard)
          const int nt = t + gridDim.x;
          bias_p = (nt < n_tiles && nt * 128 + tid < Vloc) ? hp.b[nt * 128 + tid] : 0.f;
        }
        TRACE2(3);
        for (int bb = 0; bb < nbb; ++bb) {
          TRACE2(4);
          tc::mbar_wait(&mbar[MB_L0 + bb], phL[bb]);
          phL[bb] ^= 1;
          tc::tc_fence_after();
          TRACE2(5);
          // ---- epilogue 1: dlogits of (row, 32 columns), computed in registers first ------------------
          const int r = q * 32 + lane, rl = bb * 128 + r;
          float l[32];
          tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + T_L + (uint32_t)(bb * 128 + cq * 32), l);
          TRACE2(51);
          {
            // dl = exp(l + b - lse) / B as ONE ex2 of an fma: (l + b) log2e + (log2(1/B) - lse log2e); the one-hot
            // target and the masking of columns/rows beyond the matrix are rare fix-ups outside the hot loop
            const bool rv = r0 + rl < B;
            const float cst = fmaf(-lse_c[rl], LOG2E_F, log2_inv_B);
            const int tj = tgt_c[rl] - v0 - cq * 32;
            const float4 *bg = reinterpret_cast<const float4 *>(bias_s + cq * 32);
            const int nvalid = rv ? min(32, Vloc - v0 - cq * 32) : 0;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 b4 = bg[j4];
              l[j4 * 4 + 0] = tc::ex2_ftz(fmaf(l[j4 * 4 + 0] + b4.x, LOG2E_F, cst));
              l[j4 * 4 + 1] = tc::ex2_ftz(fmaf(l[j4 * 4 + 1] + b4.y, LOG2E_F, cst));
              l[j4 * 4 + 2] = tc::ex2_ftz(fmaf(l[j4 * 4 + 2] + b4.z, LOG2E_F, cst));
              l[j4 * 4 + 3] = tc::ex2_ftz(fmaf(l[j4 * 4 + 3] + b4.w, LOG2E_F, cst));
            }
            if (tj >= 0 && tj < 32) {
              const float sub = inv_B - ((extra && rv) ? __ldg(extra + r0 + rl) : 0.f);
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j == tj) l[j] -= sub;
            }
            if (nvalid < 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j >= nvalid) l[j] = 0.f;
            }
          }
          TRACE2(6);
          if (bb > 0) {  // the dl buffer is still being read by the gradient MMAs of block bb-1
            tc::mbar_wait(&mbar[MB_G], phG);
            phG ^= 1;
          }
          TRACE2(7);
          {
            uint8_t *bh = dl_hi + (cq >> 1) * BLK, *bl = dl_lo + (cq >> 1) * BLK;
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8)
              tc::store_split8(bh, bl, r, (cq & 1) * 4 + c8, make_float4(l[c8 * 8], l[c8 * 8 + 1], l[c8 * 8 + 2], l[c8 * 8 + 3]),
                               make_float4(l[c8 * 8 + 4], l[c8 * 8 + 5], l[c8 * 8 + 6], l[c8 * 8 + 7]));
          }
          publish(MB_DL);
          TRACE2(8);
        }
        tc::mbar_wait(&mbar[MB_G], phG);  // gradient MMAs of the chunk's last block: W / dl / h operands may be overwritten
        phG ^= 1;
        tc::tc_fence_after();
        TRACE2(10);
        if (!resident) { pend_r0 = r0; pend_nbb = nbb; pend_k = k; }  // read out behind the next publish(MB_WREADY)
      }
    }
    flush_dh();
    // ---- resident dh of this CTA: TMEM -> its slice ---------------------------------------------------
    float *slice = dh_part + (int64_t)blockIdx.x * B * 64;
    for (int bb = 0; resident && bb < nbb0; ++bb) {
      const int row = bb * 128 + q * 32 + lane;
      float g[16];
      if (k > 0) {
        tc::tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + T_DH + (uint32_t)(bb * 64 + cq * 16), g);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) g[j] = 0.f;
      }
      if (row < B) {
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4 *>(slice + (int64_t)row * 64 + cq * 16 + j) = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

bool tc_bwd_supported(const rec_engine *e, int B) { (void)B; return tc_heads_supported(e); }

// The packed bf16 hi/lo image of the supervised states only needs the GRU forward: the Q step produces it on a side
// stream right after the forward pass instead of in front of the backward kernel (which sits on the critical path).
int launch_h_prepack_early(rec_engine *e, const float *h, int B) {
  if (!tc_heads_supported(e)) return REC_OK;
  h_prepack_kernel<<<cdiv(B, 128), 256, 0, e->stream>>>(h, B, e->hpack);
  REC_LAUNCH_CHECK(e);
  e->hpack_ready = true;
  return REC_OK;
}

int tc_bwd_slices(const rec_engine *e) {
  const int n_tiles = cdiv(e->Vloc, 128);
  int n_cta = e->sm_count < n_tiles ? e->sm_count : n_tiles;
  if (n_cta > e->n_dh_part - 1) n_cta = e->n_dh_part - 1;
  return n_cta;
}

// Dense (supervised) head on tensor cores; returns the number of dh slices it wrote.
int launch_head_bwd_adam_tc(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B, float step_size,
                            float bc2_sqrt, const rec_train_hparams *hp, float inv_B, int *n_slices) {
  const rec_net_params &p = e->nets[net_id].p;
  TcTrainPtrs t = {p.head_w[0], p.head_w_m[0], p.head_w_v[0], p.head_b[0], p.head_b_m[0], p.head_b_v[0]};
  const int n_tiles = cdiv(e->Vloc, 128);
  const int n_cta = tc_bwd_slices(e);
  // operands 12 x 16 KB, ones 4 KB, bias / double-buffered lse / targets 5 KB, Adam warps' transpose staging 20 KB
  const size_t smem = 1024 + 12 * (size_t)BLK + 4096 + 3072 + 2048 + 8 * 32 * 20 * 4;
  static bool attr_set[REC_MAX_DEVICES] = {};  // per device: the opt-in is a per-device function attribute
  if (!attr_set[e->dev]) {
    REC_CUDA(e, cudaFuncSetAttribute(head_bwd_adam_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    REC_CUDA(e, cudaFuncSetAttribute(head_bwd_adam_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[e->dev] = true;
  }
  if (!e->hpack_ready) {
    h_prepack_kernel<<<cdiv(B, 128), 256, 0, e->stream>>>(h, B, e->hpack);
    REC_LAUNCH_CHECK(e);
  }
  e->hpack_ready = false;
  auto kernel = B > 256 ? head_bwd_adam_tc2_kernel<true> : head_bwd_adam_tc2_kernel<false>;
  kernel<<<n_cta, BWD2_THREADS, smem, e->stream>>>(t, e->hpack, b->a, e->row_stats, B, e->Vloc, e->cfg.vocab_lo, n_tiles, inv_B,
                                                  e->dh_part, hp->beta1, hp->beta2, hp->eps, step_size, 1.f / bc2_sqrt,
                                                  trace_sel() == 0 ? e->trace : nullptr, e->d_sc, e->bwd_extra);
  REC_LAUNCH_CHECK(e);
  *n_slices = n_cta;
  return REC_OK;
}
