// Item-embedding backward + dense Adam (reference: nn.Embedding autograd + torch.optim.Adam over the
// whole [N+1,E] table, models/SQN/sqn_gru.py:50-62,173-181,248-252).
//
// The gradient of the table is never materialised densely.  Token positions p = b*L + t that hit the
// same row are combined by a position-ordered segmented sum (deterministic, no float atomics): the
// lowest position of each row is its "leader", adds the dx rows of all later duplicates in order and
// publishes slot_of_row[row] = p.  The Adam sweep then streams p/m/v of EVERY row once (24 B/param,
// the algorithmic minimum for dense-Adam semantics: rows with zero gradient still decay m, v and
// move) and picks the gradient row through slot_of_row (or 0).
#include "common.cuh"

__global__ void emb_keys_kernel(const int64_t *__restrict__ s, const int64_t *__restrict__ lens, int B, int L,
                                int N, int packed, int frozen_row, int32_t *__restrict__ keys) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= B * L) return;
  int b = p / L, t = p - b * L;
  int len = L;
  if (packed) { int64_t l = lens[b]; len = (int)(l < 1 ? 1 : (l > L ? L : l)); }
  int64_t it = s[p];
  it = it < 0 ? 0 : (it > N ? N : it);
  keys[p] = (t < len && (int)it != frozen_row) ? (int)it : -1;
}

// One warp per position.  keys are staged in shared memory when they fit.
// Positions are processed in chunks of `chunk` (a multiple of 8): duplicates are combined inside a chunk here;
// when the batch spans several chunks (large / multi-GPU global batches) emb_chunk_merge_kernel then folds
// the chunk leaders together in chunk order, which keeps the whole reduction O(P * chunk) and deterministic.
__global__ void __launch_bounds__(256) emb_segment_kernel(const int32_t *__restrict__ keys_all, int P_all, int E, int dirs,
                                                          const float *__restrict__ dx,
                                                          float *__restrict__ grad_rows,
                                                          int32_t *__restrict__ slot_of_row, int use_smem, int chunk,
                                                          uint8_t *__restrict__ leader_flag) {
  const int warps_per_chunk_blocks = chunk >> 3;                   // CTAs per chunk (8 warps = 8 positions per CTA)
  const int cidx = blockIdx.x / warps_per_chunk_blocks;            // chunk of this CTA
  const int c0 = cidx * chunk;
  const int P = min(chunk, P_all - c0);                            // positions of this chunk
  const int32_t *keys = keys_all + c0;
  dx += (int64_t)c0 * dirs * E;
  grad_rows += (int64_t)c0 * E;
  extern __shared__ int32_t skeys[];
  constexpr int LISTCAP = 160;
  __shared__ int plist[8][LISTCAP];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int32_t *kp = keys;
  if (use_smem) {
    for (int i = threadIdx.x; i < P; i += blockDim.x) skeys[i] = keys[i];
    __syncthreads();
    kp = skeys;
  }
  const int p = (blockIdx.x - cidx * warps_per_chunk_blocks) * (blockDim.x >> 5) + wid;  // position inside the chunk
  if (p >= P) return;
  const int row = kp[p];
  if (row < 0) return;
  // leader test: any earlier position with the same row?
  for (int q0 = 0; q0 < p; q0 += 32) {
    int q = q0 + lane;
    bool hit = (q < p) && (kp[q] == row);
    if (__ballot_sync(0xffffffffu, hit)) return;
  }
  // accumulate own dx (sum over directions) then later duplicates in position order;
  float acc[8];
  // E > 256 handled by looping chunks of 256 columns
  for (int e0 = 0; e0 < E; e0 += 256) {
    const int Ec = min(256, E - e0);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
    float2 acc2 = make_float2(0.f, 0.f);
    auto add_chunk = [&](int q) {
      for (int d = 0; d < dirs; ++d) {
        const float *src = dx + ((int64_t)q * dirs + d) * E + e0;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          int col = c * 32 + lane;
          if (col < Ec) acc[c] += src[col];
        }
      }
    };
    const bool fast = (E <= 64 && dirs == 1);
    int cnt = 0;
    auto flush_list = [&]() {
      for (int i0 = 0; i0 < cnt; i0 += 32) {
        float2 v[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          v[u] = make_float2(0.f, 0.f);
          if (i0 + u < cnt && 2 * lane < E)
            v[u] = *reinterpret_cast<const float2 *>(dx + (int64_t)plist[wid][i0 + u] * E + 2 * lane);
        }
#pragma unroll
        for (int u = 0; u < 32; ++u)
          if (i0 + u < cnt) { acc2.x += v[u].x; acc2.y += v[u].y; }
      }
      __syncwarp();
      cnt = 0;
    };
    if (fast) { if (2 * lane < E) acc2 = *reinterpret_cast<const float2 *>(dx + (int64_t)p * E + 2 * lane); }
    else add_chunk(p);
    for (int q0 = p + 1; q0 < P; q0 += 32) {
      int q = q0 + lane;
      unsigned m = __ballot_sync(0xffffffffu, (q < P) && (kp[q] == row));
      if (fast) {
        // Hot rows (Zipf head) have hundreds of duplicates.  Pass 1 only records matching positions
        // in a per-warp list; flush_list() then loads 16 rows at a time (independent loads, overlapped
        // latency) and adds them strictly in position order.
        if (m) {
          if (m & (1u << lane)) plist[wid][cnt + __popc(m & ((1u << lane) - 1u))] = q;
          cnt += __popc(m);
          __syncwarp();
          if (cnt > LISTCAP - 32) { flush_list(); }
        }
      } else {
        while (m) {
          int l = __ffs(m) - 1;
          m &= m - 1;
          add_chunk(q0 + l);
        }
      }
    }
    if (fast) {
      flush_list();
      if (2 * lane < E) *reinterpret_cast<float2 *>(grad_rows + (int64_t)p * E + 2 * lane) = acc2;
      continue;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      int col = c * 32 + lane;
      if (col < Ec) grad_rows[(int64_t)p * E + e0 + col] = acc[c];
    }
  }
  if (lane == 0) {
    if (leader_flag) leader_flag[c0 + p] = 1;   // several chunks: published by the merge pass
    else slot_of_row[row] = p;
  }
}

// Fold the leaders of chunk `cidx` into the global slot map: the first chunk that sees a row owns its
// accumulator, later chunks add their partial row to it.  Rows are unique inside a chunk, so no two warps of
// one launch touch the same accumulator; chunks are launched in order => fixed summation order.
__global__ void __launch_bounds__(256) emb_chunk_merge_kernel(const int32_t *__restrict__ keys, const uint8_t *__restrict__ leader_flag,
                                                              int c0, int c1, int E, float *__restrict__ grad_rows,
                                                              int32_t *__restrict__ slot_of_row) {
  const int lane = threadIdx.x & 31;
  const int p = c0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= c1 || !leader_flag[p]) return;
  const int row = keys[p];
  const int prev = slot_of_row[row];
  if (prev < 0) {
    if (lane == 0) slot_of_row[row] = p;
    return;
  }
  for (int c = lane; c < E; c += 32) grad_rows[(int64_t)prev * E + c] += grad_rows[(int64_t)p * E + c];
}

__device__ __forceinline__ void adam_update4(float4 &p, float4 &m, float4 &v, const float4 g, float b1, float b2,
                                             float eps, float step_size, float inv_bc2_sqrt) {
  adam_elem(p.x, m.x, v.x, g.x, b1, b2, eps, step_size, inv_bc2_sqrt);
  adam_elem(p.y, m.y, v.y, g.y, b1, b2, eps, step_size, inv_bc2_sqrt);
  adam_elem(p.z, m.z, v.z, g.z, b1, b2, eps, step_size, inv_bc2_sqrt);
  adam_elem(p.w, m.w, v.w, g.w, b1, b2, eps, step_size, inv_bc2_sqrt);
}

// Dense Adam sweep over a [rows, D] matrix whose gradient is row-sparse: row r has gradient
// grad_rows[slot_of_row[r] * grad_stride4 + c] when slot_of_row[r] >= 0, else 0.  HBM-bound streaming
// kernel (24 B/param): every thread keeps UNROLL independent (p, m, v) float4 triples in flight.
// Optional bias vector (one element per row) is updated by the thread that owns column chunk 0.
struct AdamStreamSet {  // up to 3 same-shaped tensors updated by one launch (blockIdx.y)
  float4 *p[3], *m[3], *v[3];
  const float4 *grad_rows[3];
  float *bp[3], *bm[3], *bv[3];
  const float *bgrad[3];
};

template <int UNROLL>
__global__ void __launch_bounds__(256) adam_stream_kernel(AdamStreamSet ts, const int32_t *__restrict__ slot_of_row,
                                                          int grad_stride4, uint32_t n4, int D4, int d4_shift, float b1,
                                                          float b2, float eps, float step_size, float inv_bc2_sqrt,
                                                          const float *__restrict__ sc) {
  if (sc) { step_size = sc[0]; inv_bc2_sqrt = sc[1]; }
  float4 *__restrict__ p = ts.p[blockIdx.y];
  float4 *__restrict__ m = ts.m[blockIdx.y];
  float4 *__restrict__ v = ts.v[blockIdx.y];
  const float4 *__restrict__ grad_rows = ts.grad_rows[blockIdx.y];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * UNROLL) {
    float4 pv[UNROLL], mv[UNROLL], vv[UNROLL];
    int slot[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint32_t i = i0 + u * stride;
      if (i < n4) {
        pv[u] = p[i]; mv[u] = m[i]; vv[u] = v[i];
        slot[u] = __ldg(slot_of_row + (d4_shift >= 0 ? (i >> d4_shift) : (i / (uint32_t)D4)));
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const uint32_t i = i0 + u * stride;
      if (i < n4) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (slot[u] >= 0) {
          const uint32_t c = d4_shift >= 0 ? (i & (uint32_t)(D4 - 1)) : (i % (uint32_t)D4);
          g = __ldg(grad_rows + (size_t)slot[u] * grad_stride4 + c);
        }
        adam_update4(pv[u], mv[u], vv[u], g, b1, b2, eps, step_size, inv_bc2_sqrt);
        p[i] = pv[u]; m[i] = mv[u]; v[i] = vv[u];
      }
    }
  }
}

// bias vectors of the same tensors: one element per row, gradient through the same slot map
__global__ void __launch_bounds__(256) adam_bias_kernel(AdamStreamSet ts, const int32_t *__restrict__ slot_of_row,
                                                        int bgrad_stride, int rows, float b1, float b2, float eps,
                                                        float step_size, float inv_bc2_sqrt, const float *__restrict__ sc) {
  if (sc) { step_size = sc[0]; inv_bc2_sqrt = sc[1]; }
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float *bp = ts.bp[blockIdx.y], *bm = ts.bm[blockIdx.y], *bv = ts.bv[blockIdx.y];
  int slot = slot_of_row[r];
  float g = slot >= 0 ? ts.bgrad[blockIdx.y][(size_t)slot * bgrad_stride] : 0.f;
  float p = bp[r], m = bm[r], v = bv[r];
  adam_elem(p, m, v, g, b1, b2, eps, step_size, inv_bc2_sqrt);
  bp[r] = p; bm[r] = m; bv[r] = v;
}

static int launch_adam_stream_set(rec_engine *e, const AdamStreamSet &ts, int n_tensors, bool with_bias, int64_t rows, int D,
                                  const int32_t *slot_of_row, int grad_stride, int bgrad_stride,
                                  const rec_train_hparams *hp, float step_size, float bc2_sqrt, int time_slot = -1) {
  // time_slot >= 0 (kernel-timing mode): CUDA events around the ONE adam_stream_kernel launch (rec_last_kernel_ms)
  const int64_t n4 = rows * (D / 4);
  if (n4 >= (int64_t)1 << 31) REC_FAIL(e, REC_EINVAL, "adam_stream: tensor too large (%lld float4)", (long long)n4);
  // The sweep is persistent (grid-stride) with a SMALL resident footprint: `ctas` CTAs per SM keep enough bytes in
  // flight to saturate the HBM (UNROLL x 48 B per thread) while leaving most warp slots and registers of every
  // SM to the latency-bound kernels of the other graph branches, which would otherwise starve behind it.
  static const int ctas = getenv("REC_SWEEP_CTAS") ? atoi(getenv("REC_SWEEP_CTAS")) : 2;
  static const int unroll = getenv("REC_SWEEP_UNROLL") ? atoi(getenv("REC_SWEEP_UNROLL")) : 2;
  const int D4 = D / 4;
  int shift = -1;
  if ((D4 & (D4 - 1)) == 0) { shift = 0; while ((1 << shift) < D4) ++shift; }
  int64_t want = cdiv64(n4, 256 * unroll);
  int per = (e->sm_count * ctas) / n_tensors;
  int blocks = (int)(want < (int64_t)per ? want : (int64_t)per);
  if (blocks < 1) blocks = 1;
  dim3 grid(blocks, n_tensors);
  if (e->timing && time_slot >= 0) cudaEventRecord(e->ev[2 * time_slot], e->stream);
  if (unroll == 4)
    adam_stream_kernel<4><<<grid, 256, 0, e->stream>>>(ts, slot_of_row, grad_stride / 4, (uint32_t)n4, D4, shift, hp->beta1,
                                                       hp->beta2, hp->eps, step_size, 1.f / bc2_sqrt, e->d_sc);
  else
    adam_stream_kernel<2><<<grid, 256, 0, e->stream>>>(ts, slot_of_row, grad_stride / 4, (uint32_t)n4, D4, shift, hp->beta1,
                                                       hp->beta2, hp->eps, step_size, 1.f / bc2_sqrt, e->d_sc);
  if (e->timing && time_slot >= 0) cudaEventRecord(e->ev[2 * time_slot + 1], e->stream);
  REC_LAUNCH_CHECK(e);
  if (with_bias) {
    dim3 g2(cdiv((int)rows, 256), n_tensors);
    adam_bias_kernel<<<g2, 256, 0, e->stream>>>(ts, slot_of_row, bgrad_stride, (int)rows, hp->beta1, hp->beta2, hp->eps, step_size,
                                               1.f / bc2_sqrt, e->d_sc);
    REC_LAUNCH_CHECK(e);
  }
  return REC_OK;
}

int launch_adam_stream(rec_engine *e, float *p, float *m, float *v, int64_t rows, int D, const int32_t *slot_of_row,
                       const float *grad_rows, int grad_stride, float *bp, float *bm, float *bv, const float *bgrad,
                       int bgrad_stride, const rec_train_hparams *hp, float step_size, float bc2_sqrt) {
  AdamStreamSet ts = {};
  ts.p[0] = (float4 *)p; ts.m[0] = (float4 *)m; ts.v[0] = (float4 *)v; ts.grad_rows[0] = (const float4 *)grad_rows;
  ts.bp[0] = bp; ts.bm[0] = bm; ts.bv[0] = bv; ts.bgrad[0] = bgrad;
  return launch_adam_stream_set(e, ts, 1, bp != nullptr, rows, D, slot_of_row, grad_stride, bgrad_stride, hp, step_size, bc2_sqrt);
}

// ---- row-sparse gradients of the Q heads --------------------------------------------------------
// dW_{1+j}[a_b] += dq[b,j] * h[b], db_{1+j}[a_b] += dq[b,j]: one warp per batch row; the lowest b of each
// distinct action is the leader and sums its duplicates in batch order (deterministic, no atomics).
// grad_rows[(b*n_q + j), 0:D], bgrad[b*n_q + j], slot_of_row[a_b - vocab_lo] = b.
__global__ void __launch_bounds__(256) q_grad_rows_kernel(const int64_t *__restrict__ a, const float *__restrict__ dq,
                                                          const float *__restrict__ h, int B, int D, int n_q, int Vloc,
                                                          int vocab_lo, float *__restrict__ grad_rows,
                                                          float *__restrict__ bgrad, int32_t *__restrict__ slot_of_row,
                                                          int dq_stride) {
  // the action ids of the whole batch are staged in shared memory: the two scans below are latency chains
  // (load -> ballot -> next) and ran at L2 latency per iteration when they read global memory
  extern __shared__ int32_t sa[];
  for (int i = threadIdx.x; i < B; i += blockDim.x) sa[i] = (int32_t)a[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // one warp per (batch row, Q head)
  if (w >= B * n_q) return;
  const int b = w / n_q, j = w - b * n_q;
  const int key = sa[b];
  const int loc = key - vocab_lo;
  if (loc < 0 || loc >= Vloc) return;
  for (int q0 = 0; q0 < b; q0 += 32) {
    int q = q0 + lane;
    if (__ballot_sync(0xffffffffu, q < b && sa[q] == key)) return;
  }
  // Pass 1 of each column block only records the matching rows in a per-warp list (shared-memory reads and
  // ballots, no global loads); flush() then fetches dq and the h rows of 16 list entries at a time -- independent
  // loads -- and accumulates strictly in batch order.  A popular action has hundreds of duplicates: chasing them
  // one 32-row window at a time cost two dependent global latencies per window.
  constexpr int CAP = 128;
  __shared__ int plist[8][CAP];
  const int wid = threadIdx.x >> 5;
  for (int d0 = 0; d0 < D; d0 += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float bsum = 0.f;
    const int col = d0 + lane * 4;
    int cnt = 0;
    auto flush = [&]() {
      for (int i0 = 0; i0 < cnt; i0 += 16) {
        float4 hv[16];
        float g[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          g[u] = 0.f;
          hv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i0 + u < cnt) {
            const int q = plist[wid][i0 + u];
            g[u] = dq[q * dq_stride + j];
            if (col < D) hv[u] = *reinterpret_cast<const float4 *>(h + (int64_t)q * D + col);
          }
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          if (i0 + u < cnt) {
            bsum += g[u];
            acc.x = fmaf(g[u], hv[u].x, acc.x); acc.y = fmaf(g[u], hv[u].y, acc.y);
            acc.z = fmaf(g[u], hv[u].z, acc.z); acc.w = fmaf(g[u], hv[u].w, acc.w);
          }
        }
      }
      __syncwarp();
      cnt = 0;
    };
    for (int q0 = b; q0 < B; q0 += 32) {
      const int q = q0 + lane;
      const bool hit = q < B && sa[q] == key;
      const unsigned mm = __ballot_sync(0xffffffffu, hit);
      if (!mm) continue;
      if (hit) plist[wid][cnt + __popc(mm & ((1u << lane) - 1u))] = q;
      cnt += __popc(mm);
      __syncwarp();
      if (cnt > CAP - 32) flush();
    }
    flush();
    if (col < D) *reinterpret_cast<float4 *>(grad_rows + ((int64_t)b * n_q + j) * D + col) = acc;
    if (d0 == 0 && lane == 0) bgrad[b * n_q + j] = bsum;
  }
  if (lane == 0 && j == 0) slot_of_row[loc] = b;
}

__global__ void q_slot_reset_kernel(const int64_t *__restrict__ a, int B, int Vloc, int vocab_lo,
                                    int32_t *__restrict__ slot_of_row) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int64_t loc = a[b] - vocab_lo;
  if (loc >= 0 && loc < Vloc) slot_of_row[loc] = -1;
}

// Adam on heads first_head .. first_head + n - 1 of net `net_id`: streaming sweep with row-sparse gradients
// (dL/dQ_j(s_b, a_b) = dq[b * dq_stride + j]).  At most 3 tensors share one sweep launch.
int launch_q_heads_adam_ex(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B, float step_size,
                           float bc2_sqrt, const rec_train_hparams *hp, int first_head, int n, const float *dq,
                           int dq_stride, int wait_mark) {
  const int D = e->D;
  const rec_net_params &p = e->nets[net_id].p;
  if ((size_t)B * sizeof(int32_t) > 48 * 1024) REC_FAIL(e, REC_EINVAL, "Q-head gradient rows: batch of %d sessions exceeds the 12288 supported", B);
  q_grad_rows_kernel<<<cdiv(B * n, 8), 256, (size_t)B * sizeof(int32_t), e->stream>>>(b->a, dq, h, B, D, n, e->Vloc, e->cfg.vocab_lo, e->q_grad_rows,
                                                       e->q_bgrad, e->q_slot, dq_stride);
  REC_LAUNCH_CHECK(e);
  if (wait_mark >= 0) side_wait_mark(e, wait_mark);
  for (int j0 = 0; j0 < n; j0 += 3) {
    const int nj = n - j0 < 3 ? n - j0 : 3;
    AdamStreamSet ts = {};
    for (int j = 0; j < nj; ++j) {
      const int hd = first_head + j0 + j;
      ts.p[j] = (float4 *)p.head_w[hd]; ts.m[j] = (float4 *)p.head_w_m[hd]; ts.v[j] = (float4 *)p.head_w_v[hd];
      ts.grad_rows[j] = (const float4 *)(e->q_grad_rows + (int64_t)(j0 + j) * D);
      ts.bp[j] = p.head_b[hd]; ts.bm[j] = p.head_b_m[hd]; ts.bv[j] = p.head_b_v[hd];
      ts.bgrad[j] = e->q_bgrad + j0 + j;
    }
    int rc = launch_adam_stream_set(e, ts, nj, true, e->Vloc, D, e->q_slot, n * D, n, hp, step_size, bc2_sqrt, j0 == 0 ? 3 : -1);
    if (rc) return rc;
  }
  q_slot_reset_kernel<<<cdiv(B, 256), 256, 0, e->stream>>>(b->a, B, e->Vloc, e->cfg.vocab_lo, e->q_slot);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

// Adam on every Q head (heads 1..n_q) of net `net_id` (SQN / SMORL).
int launch_q_heads_adam(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B, float step_size,
                        float bc2_sqrt, const rec_train_hparams *hp, int wait_mark) {
  return launch_q_heads_adam_ex(e, net_id, h, b, B, step_size, bc2_sqrt, hp, 1, e->cfg.n_heads - 1, e->dq, 3, wait_mark);
}

__global__ void emb_reset_kernel(const int32_t *__restrict__ keys, int P, int32_t *__restrict__ slot_of_row) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P && keys[p] >= 0) slot_of_row[keys[p]] = -1;
}

__global__ void fill_i32_kernel(int32_t *p, int64_t n, int32_t v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

int launch_fill_i32(rec_engine *e, int32_t *p, int64_t n, int32_t v) {
  fill_i32_kernel<<<(int)cdiv64(n, 256 * 8) < 1 ? 1 : (int)cdiv64(n, 256 * 8), 256, 0, e->stream>>>(p, n, v);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}


// ---- sort-based duplicate combining (E = 64, one direction, P <= 32768) ------------------------------------
// emb_rank_kernel orders the token positions by (row, position) (see below).
// emb_tilesum64_kernel: one warp per tile of 32 sorted entries loads its 32 dx rows with independent loads and
// walks them in sorted order; a row whose run continues from the previous tile leaves a carry that the last
// warp to finish folds into the run's accumulator in tile order.  The summation order depends only on the
// sorted order => deterministic, no float atomics.
// Sorting by brute-force ranking: P <= 8192 keys means <= 67 M comparisons (1 G at the 32768 cap), spread over every SM, which beats
// a single-CTA sorting network by far.  Every CTA stages the row keys of all positions in shared memory
// (invalid positions get INT_MAX), 8 warps share 32 elements and count the entries ordered before each:
// rank(i) = #{j : row_j < row_i} + #{j < i : row_j == row_i}.  Ranks are a permutation, so the scatter is
// conflict-free.
// Batches beyond EMB_CHUNK positions are ranked chunk by chunk (quadratic work only inside a chunk) and the
// sorted chunks are merged by emb_merge_rank_kernel.
constexpr int EMB_CHUNK = 4096;

__global__ void __launch_bounds__(256) emb_rank_kernel(const int32_t *__restrict__ keys, int P, int C,
                                                       int32_t *__restrict__ out_pos, int32_t *__restrict__ out_row,
                                                       int32_t *__restrict__ counts) {
  extern __shared__ int32_t srow[];
  const int tid = threadIdx.x;
  const int c = (blockIdx.x * 32) / C;        // chunk of this CTA (C is a multiple of 32)
  const int c0 = c * C;
  const int Pc = min(C, P - c0);
  const int Ppad = (Pc + 15) & ~15;
  int n_valid = 0;
  for (int i = tid; i < Ppad; i += 256) {   // keys were computed once by emb_keys_kernel (-1 = no gradient)
    const int k = i < Pc ? keys[c0 + i] : -1;
    n_valid += k >= 0;
    srow[i] = k >= 0 ? k : 0x7fffffff;
  }
  if (blockIdx.x * 32 == c0) {
    __shared__ int nv[8];
    for (int o = 16; o > 0; o >>= 1) n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
    if ((tid & 31) == 0) nv[tid >> 5] = n_valid;
    __syncthreads();
    if (tid == 0) counts[c] = nv[0] + nv[1] + nv[2] + nv[3] + nv[4] + nv[5] + nv[6] + nv[7];
  } else {
    __syncthreads();
  }
  // lane = element, warp = 1/8 of the scan range: every shared-memory read is a warp-wide broadcast and the
  // loop is branch-free (row_j is counted when it is below row_i + [j < i])
  __shared__ int part_cnt[8][32];
  const int lane = tid & 31, wid = tid >> 5;
  const int i = blockIdx.x * 32 + lane - c0;  // position inside the chunk
  const int ri = i < Pc ? srow[i] : 0x7ffffffe;
  const int Q = Ppad >> 3;                  // entries per warp (multiple of 2)
  int cnt = 0;
  const int2 *src = reinterpret_cast<const int2 *>(srow + wid * Q);
  const int jb = wid * Q;
#pragma unroll 4
  for (int g = 0; g < (Q >> 1); ++g) {
    const int2 r = src[g];
    const int j0 = jb + 2 * g;
    cnt += (r.x < ri + (j0 < i ? 1 : 0)) ? 1 : 0;
    cnt += (r.y < ri + (j0 + 1 < i ? 1 : 0)) ? 1 : 0;
  }
  part_cnt[wid][lane] = cnt;
  __syncthreads();
  if (wid == 0 && i < Pc && ri != 0x7fffffff) {
    int rank = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) rank += part_cnt[w][lane];
    out_pos[c0 + rank] = c0 + i;
    out_row[c0 + rank] = ri;
  }
}

// Global rank of entry r of sorted chunk c = r + sum over the other chunks of the entries ordered before it:
// chunks hold ascending position ranges, so an equal row of an EARLIER chunk comes first (upper bound) and of a
// LATER chunk comes after (lower bound).  All chunk lists are staged in shared memory; the searches are binary.
__global__ void __launch_bounds__(256) emb_merge_rank_kernel(const int32_t *__restrict__ c_pos, const int32_t *__restrict__ c_row,
                                                             const int32_t *__restrict__ counts, int C, int n_chunks,
                                                             int32_t *__restrict__ sorted_pos, int32_t *__restrict__ sorted_row,
                                                             int cap) {
  extern __shared__ int32_t srow[];
  __shared__ int cnt[16];
  const int tid = threadIdx.x;
  if (tid < n_chunks) cnt[tid] = counts[tid];
  __syncthreads();
  for (int i = tid; i < n_chunks * C; i += 256) {
    const int c = i / C, r = i - c * C;
    srow[i] = r < cnt[c] ? c_row[i] : 0x7fffffff;
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid == 0) {
    int tot = 0;
    for (int c = 0; c < n_chunks; ++c) tot += cnt[c];
    sorted_row[cap] = tot;
  }
  const int e = blockIdx.x * 256 + tid;
  if (e >= n_chunks * C) return;
  const int c = e / C, r = e - c * C;
  if (r >= cnt[c]) return;
  const int row = srow[e];
  int rank = r;
  for (int o = 0; o < n_chunks; ++o) {
    if (o == c) continue;
    const int32_t *list = srow + o * C;
    const int thr = row + (o < c ? 1 : 0);   // first index with list[idx] >= thr
    int lo = 0, hi = cnt[o];
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (list[mid] < thr) lo = mid + 1; else hi = mid;
    }
    rank += lo;
  }
  sorted_pos[rank] = c_pos[e];
  sorted_row[rank] = row;
}

__global__ void __launch_bounds__(256) emb_tilesum64_kernel(const int32_t *__restrict__ sorted_pos, int32_t *__restrict__ sorted_row,
                                                            int cap, const float *__restrict__ dx, float *__restrict__ grad_rows,
                                                            int32_t *__restrict__ slot_of_row, float *__restrict__ carry,
                                                            int32_t *__restrict__ tile_meta, int E, int dirs) {
  // blockIdx.y = 64-column chunk of the embedding row (E = 64: one chunk); dx rows are [position][direction][E] and
  // the directions of a position are added first (direction order), then the positions in sorted order
  const int lane = threadIdx.x & 31;
  const int col = blockIdx.y * 64 + 2 * lane;
  const int T = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_valid = sorted_row[cap];
  const int n_tiles = (n_valid + 31) >> 5;
  if (T >= n_tiles) return;
  const int idx = 32 * T + lane;
  const bool valid = idx < n_valid;
  const int pos = valid ? sorted_pos[idx] : -1;
  const int row = valid ? sorted_row[idx] : -2;
  int prev = __shfl_up_sync(0xffffffffu, row, 1);
  if (lane == 0) prev = idx > 0 ? sorted_row[idx - 1] : -1;
  const bool head = valid && row != prev;
  const unsigned hmask = __ballot_sync(0xffffffffu, head), vmask = __ballot_sync(0xffffffffu, valid);
  if (head) slot_of_row[row] = pos;
  float2 v[32];
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    const int pp = __shfl_sync(0xffffffffu, pos, u);
    v[u] = make_float2(0.f, 0.f);
    if (pp >= 0) {
      v[u] = *reinterpret_cast<const float2 *>(dx + (int64_t)pp * dirs * E + col);
      if (dirs == 2) {
        const float2 w = *reinterpret_cast<const float2 *>(dx + ((int64_t)pp * 2 + 1) * E + col);
        v[u].x += w.x; v[u].y += w.y;
      }
    }
  }
  float2 acc = make_float2(0.f, 0.f);
  int cur = -1;  // leader position of the open run; -1 = run continued from the previous tile (-> carry)
  auto flush = [&]() {
    float *dst = cur < 0 ? carry + (int64_t)T * E : grad_rows + (int64_t)cur * E;
    *reinterpret_cast<float2 *>(dst + col) = acc;
  };
#pragma unroll
  for (int u = 0; u < 32; ++u) {
    if ((vmask >> u) & 1u) {
      if ((hmask >> u) & 1u) {
        flush();
        cur = __shfl_sync(0xffffffffu, pos, u);
        acc = make_float2(0.f, 0.f);
      }
      acc.x += v[u].x;
      acc.y += v[u].y;
    }
  }
  flush();
  if (lane == 0) {
    tile_meta[2 * T] = (hmask & 1u) ? 0 : 1;
    tile_meta[2 * T + 1] = cur;
  }
}

// One warp per run that crosses a tile boundary (the warp of the run's first continuation tile): adds the
// carries of the following tiles to the run's accumulator in tile order.
__global__ void __launch_bounds__(256) emb_carry_kernel(const int32_t *__restrict__ sorted_row, int cap,
                                                        float *__restrict__ grad_rows, const float *__restrict__ carry,
                                                        const int32_t *__restrict__ tile_meta, int E) {
  const int lane = threadIdx.x & 31;
  const int col = blockIdx.y * 64 + 2 * lane;
  const int T = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_tiles = (sorted_row[cap] + 31) >> 5;
  if (T < 1 || T >= n_tiles) return;
  if (!tile_meta[2 * T]) return;                 // tile starts with a run head
  const int leader = tile_meta[2 * (T - 1) + 1];
  if (leader < 0) return;                        // previous tile is itself a continuation: not the first one
  float2 *dst = reinterpret_cast<float2 *>(grad_rows + (int64_t)leader * E + col);
  float2 acc = *dst;
  for (int T0 = T; T0 < n_tiles; T0 += 32) {
    // tiles T0.. belong to the run while they start inside it; the run ends in the first tile that has a head
    const int t = T0 + lane;
    const bool cont = t < n_tiles && tile_meta[2 * t] != 0;
    const bool ends = t < n_tiles && tile_meta[2 * t + 1] >= 0;
    const unsigned nc = __ballot_sync(0xffffffffu, !cont), en = __ballot_sync(0xffffffffu, ends);
    int n = nc ? __ffs(nc) - 1 : 32;             // tiles of this batch that continue the run
    if (en && __ffs(en) - 1 < n) n = __ffs(en);  // the tile where the run ends still carries its tail
    float2 v[32];
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      v[u] = make_float2(0.f, 0.f);
      if (u < n) v[u] = *reinterpret_cast<const float2 *>(carry + (int64_t)(T0 + u) * E + col);
    }
#pragma unroll
    for (int u = 0; u < 32; ++u) {
      acc.x += v[u].x;
      acc.y += v[u].y;
    }
    if (n < 32) break;
  }
  *dst = acc;
}

int launch_embedding_update(rec_engine *e, int net_id, const int64_t *s, const int64_t *lengths, int B,
                            float step_size, float bc2_sqrt, const rec_train_hparams *hp, int stages) {
  const rec_config &c = e->cfg;
  const int L = c.state_size, E = c.embedding_dim, P = B * L;
  NetBind &nb = e->nets[net_id];
  const bool sorted_path = (E % 64 == 0 && e->dirs <= 2 && P <= 32768);
  static bool rank_attr_set[REC_MAX_DEVICES] = {};
  if (sorted_path && !rank_attr_set[e->dev]) {
    REC_CUDA(e, cudaFuncSetAttribute(emb_merge_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * (int)sizeof(int32_t)));
    rank_attr_set[e->dev] = true;
  }
  if (stages & 1) {
    emb_keys_kernel<<<cdiv(P, 256), 256, 0, e->stream>>>(s, lengths, B, L, c.item_num, c.use_packed_seq,
                                                        c.frozen_pad_row, e->emb_keys);
    if (sorted_path) {
      REC_LAUNCH_CHECK(e);
      const int cap = c.max_batch * L;
      const int n_chunks = cdiv(P, EMB_CHUNK);
      const int Cs = n_chunks > 1 ? EMB_CHUNK : ((P + 31) & ~31);
      const size_t smem = (size_t)((Cs + 15) & ~15) * sizeof(int32_t);
      if (n_chunks == 1) {
        emb_rank_kernel<<<cdiv(P, 32), 256, smem, e->stream>>>(e->emb_keys, P, Cs, e->emb_sorted, e->emb_seg, e->emb_seg + cap);
      } else {
        emb_rank_kernel<<<cdiv(P, 32), 256, smem, e->stream>>>(e->emb_keys, P, Cs, e->emb_csort, e->emb_csort + e->emb_csort_n,
                                                               e->emb_ccount);
        REC_LAUNCH_CHECK(e);
        emb_merge_rank_kernel<<<cdiv(n_chunks * EMB_CHUNK, 256), 256, (size_t)n_chunks * EMB_CHUNK * sizeof(int32_t), e->stream>>>(
            e->emb_csort, e->emb_csort + e->emb_csort_n, e->emb_ccount, EMB_CHUNK, n_chunks, e->emb_sorted, e->emb_seg, cap);
      }
    }
    REC_LAUNCH_CHECK(e);
  }
  if ((stages & 2) && sorted_path) {
    const dim3 tg(cdiv(cdiv(P, 32), 8), E / 64);
    emb_tilesum64_kernel<<<tg, 256, 0, e->stream>>>(e->emb_sorted, e->emb_seg, c.max_batch * L, e->dx, e->emb_grad_rows, e->emb_slot,
                                                   e->emb_carry, e->emb_tmeta, E, e->dirs);
    REC_LAUNCH_CHECK(e);
    emb_carry_kernel<<<tg, 256, 0, e->stream>>>(e->emb_seg, c.max_batch * L, e->emb_grad_rows, e->emb_carry, e->emb_tmeta, E);
    REC_LAUNCH_CHECK(e);
  } else if (stages & 2) {
    const int chunk = 4096;
    const int n_chunks = cdiv(P, chunk);
    const int span = n_chunks > 1 ? chunk : P;
    int use_smem = 1;
    size_t smem = (size_t)span * sizeof(int32_t);
    static bool attr_set[REC_MAX_DEVICES] = {};  // per device: the opt-in is a per-device function attribute
    if (!attr_set[e->dev]) {
      REC_CUDA(e, cudaFuncSetAttribute(emb_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_set[e->dev] = true;
    }
    uint8_t *flags = n_chunks > 1 ? e->emb_leader : nullptr;
    if (flags) REC_CUDA(e, cudaMemsetAsync(flags, 0, (size_t)P, e->stream));
    emb_segment_kernel<<<n_chunks > 1 ? n_chunks * (chunk / 8) : cdiv(P, 8), 256, smem, e->stream>>>(
        e->emb_keys, P, E, e->dirs, e->dx, e->emb_grad_rows, e->emb_slot, use_smem, n_chunks > 1 ? chunk : ((P + 7) / 8) * 8, flags);
    REC_LAUNCH_CHECK(e);
    for (int ch = 0; ch < n_chunks && n_chunks > 1; ++ch) {
      const int c0 = ch * chunk, c1 = c0 + chunk < P ? c0 + chunk : P;
      emb_chunk_merge_kernel<<<cdiv(c1 - c0, 8), 256, 0, e->stream>>>(e->emb_keys, flags, c0, c1, E, e->emb_grad_rows, e->emb_slot);
      REC_LAUNCH_CHECK(e);
    }
  }
  if (stages & 4) {
    if (e->timing) cudaEventRecord(e->ev[4], e->stream);
    // row-sharded table (rec_set_embedding_shard): this rank sweeps the rows it owns; the other rows of its copy are
    // refreshed from their owners before they are read (rec_emb_rows_gather / _scatter)
    const int64_t lo = e->emb_row_hi > 0 ? e->emb_row_lo : 0, hi = e->emb_row_hi > 0 ? e->emb_row_hi : (int64_t)c.item_num + 1;
    int rc = launch_adam_stream(e, nb.p.emb + lo * E, nb.p.emb_m + lo * E, nb.p.emb_v + lo * E, hi - lo, E, e->emb_slot + lo,
                                e->emb_grad_rows, E, nullptr, nullptr, nullptr, nullptr, 0, hp, step_size, bc2_sqrt);
    if (rc) return rc;
    if (e->timing) cudaEventRecord(e->ev[5], e->stream);
    emb_reset_kernel<<<cdiv(P, 256), 256, 0, e->stream>>>(e->emb_keys, P, e->emb_slot);
    REC_LAUNCH_CHECK(e);
  }
  return REC_OK;
}


// ---- row-sharded embedding table -----------------------------------------------------
// Every rank keeps a full-size copy of the table but OWNS -- sweeps with Adam -- only rows [lo, hi).  Before a step
// reads token rows, their current values travel from the owners: every rank gathers the rows it owns into a
// [n, E] buffer (zeros elsewhere), ONE all-reduce(sum) of that buffer delivers each row from its single owner
// bit-exactly (x + 0 + ... + 0), and the rows a rank does not own are written into its copy.
__global__ void emb_rows_gather_kernel(const float4 *__restrict__ table, const int64_t *__restrict__ ids, int64_t n, int E4,
                                       int64_t lo, int64_t hi, float4 *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * E4) return;
  const int64_t r = i / E4;
  const int c = (int)(i - r * E4);
  const int64_t id = ids[r];
  out[i] = (id >= lo && id < hi) ? table[id * E4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
}
__global__ void emb_rows_scatter_kernel(float4 *__restrict__ table, const int64_t *__restrict__ ids, int64_t n, int E4,
                                        int64_t lo, int64_t hi, int64_t n_rows, const float4 *__restrict__ rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * E4) return;
  const int64_t r = i / E4;
  const int c = (int)(i - r * E4);
  const int64_t id = ids[r];
  if (id < 0 || id >= n_rows || (id >= lo && id < hi)) return;  // owned rows are already current
  table[id * E4 + c] = rows[i];  // a token that occurs twice writes the same bits twice
}

int launch_emb_rows(rec_engine *e, int net_id, const int64_t *ids, int64_t n, float *rows, bool scatter) {
  const rec_config &c = e->cfg;
  const int E4 = c.embedding_dim / 4;
  const int64_t n_rows = (int64_t)c.item_num + 1;
  const int64_t lo = e->emb_row_hi > 0 ? e->emb_row_lo : 0, hi = e->emb_row_hi > 0 ? e->emb_row_hi : n_rows;
  if (n == 0) return REC_OK;
  const unsigned blocks = (unsigned)((n * E4 + 255) / 256);
  if (scatter)
    emb_rows_scatter_kernel<<<blocks, 256, 0, e->stream>>>((float4 *)e->nets[net_id].p.emb, ids, n, E4, lo, hi, n_rows,
                                                          (const float4 *)rows);
  else
    emb_rows_gather_kernel<<<blocks, 256, 0, e->stream>>>((const float4 *)e->nets[net_id].p.emb, ids, n, E4, lo, hi, (float4 *)rows);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}
