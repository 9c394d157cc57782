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
__global__ void __launch_bounds__(256) emb_segment_kernel(const int32_t *__restrict__ keys, int P, int E, int dirs,
                                                          const float *__restrict__ dx,
                                                          float *__restrict__ grad_rows,
                                                          int32_t *__restrict__ slot_of_row, int use_smem) {
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
  const int p = blockIdx.x * (blockDim.x >> 5) + wid;
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
      for (int i0 = 0; i0 < cnt; i0 += 16) {
        float2 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          v[u] = make_float2(0.f, 0.f);
          if (i0 + u < cnt && 2 * lane < E)
            v[u] = *reinterpret_cast<const float2 *>(dx + (int64_t)plist[wid][i0 + u] * E + 2 * lane);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u)
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
  if (lane == 0) slot_of_row[row] = p;
}

__device__ __forceinline__ void adam_update4(float4 &p, float4 &m, float4 &v, const float4 g, float b1, float b2,
                                             float eps, float step_size, float bc2_sqrt) {
#define REC_AD1(c)                                         \
  m.c = m.c + (g.c - m.c) * (1.f - b1);                    \
  v.c = v.c * b2 + ((1.f - b2) * g.c) * g.c;               \
  p.c = p.c + (-step_size * m.c) / (sqrtf(v.c) / bc2_sqrt + eps);
  REC_AD1(x) REC_AD1(y) REC_AD1(z) REC_AD1(w)
#undef REC_AD1
}

// Dense Adam sweep over the table; one float4 per thread, grid-stride.
__global__ void __launch_bounds__(256) emb_adam_kernel(float4 *__restrict__ p, float4 *__restrict__ m,
                                                       float4 *__restrict__ v,
                                                       const int32_t *__restrict__ slot_of_row,
                                                       const float4 *__restrict__ grad_rows, int64_t n4, int E4,
                                                       float b1, float b2, float eps, float step_size,
                                                       float bc2_sqrt) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / E4;
    int c = (int)(i - row * E4);
    int slot = __ldg(slot_of_row + row);
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (slot >= 0) g = __ldg(grad_rows + (int64_t)slot * E4 + c);
    float4 pv = p[i], mv = m[i], vv = v[i];
    adam_update4(pv, mv, vv, g, b1, b2, eps, step_size, bc2_sqrt);
    p[i] = pv; m[i] = mv; v[i] = vv;
  }
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

int launch_embedding_update(rec_engine *e, int net_id, const int64_t *s, const int64_t *lengths, int B,
                            float step_size, float bc2_sqrt, const rec_train_hparams *hp) {
  const rec_config &c = e->cfg;
  const int L = c.state_size, E = c.embedding_dim, P = B * L;
  NetBind &nb = e->nets[net_id];
  emb_keys_kernel<<<cdiv(P, 256), 256, 0, e->stream>>>(s, lengths, B, L, c.item_num, c.use_packed_seq,
                                                      c.frozen_pad_row, e->emb_keys);
  REC_LAUNCH_CHECK(e);
  int use_smem = (size_t)P * sizeof(int32_t) <= 96 * 1024;
  size_t smem = use_smem ? (size_t)P * sizeof(int32_t) : 0;
  static bool attr_set = false;
  if (!attr_set) {
    REC_CUDA(e, cudaFuncSetAttribute(emb_segment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  emb_segment_kernel<<<cdiv(P, 8), 256, smem, e->stream>>>(e->emb_keys, P, E, e->dirs, e->dx, e->emb_grad_rows,
                                                          e->emb_slot, use_smem);
  REC_LAUNCH_CHECK(e);
  const int64_t n4 = (int64_t)(c.item_num + 1) * (E / 4);
  int blocks = (int)(cdiv64(n4, 256 * 4) < (int64_t)e->sm_count * 16 ? cdiv64(n4, 256 * 4) : (int64_t)e->sm_count * 16);
  if (blocks < 1) blocks = 1;
  if (e->timing) cudaEventRecord(e->ev[4], e->stream);
  emb_adam_kernel<<<blocks, 256, 0, e->stream>>>((float4 *)nb.p.emb, (float4 *)nb.p.emb_m, (float4 *)nb.p.emb_v,
                                                e->emb_slot, (const float4 *)e->emb_grad_rows, n4, E / 4, hp->beta1,
                                                hp->beta2, hp->eps, step_size, bc2_sqrt);
  REC_LAUNCH_CHECK(e);
  if (e->timing) cudaEventRecord(e->ev[5], e->stream);
  emb_reset_kernel<<<cdiv(P, 256), 256, 0, e->stream>>>(e->emb_keys, P, e->emb_slot);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}
