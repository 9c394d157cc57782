// GRU trunk: fused embedding gather + input projection + recurrence (forward), BPTT, weight
// gradients and the GRU-parameter Adam update.
//
// Semantics follow torch.nn.GRU as used by the reference (models/GRU4Rec/model.py:49-55,74-77;
// models/SQN/sqn_gru.py:69-75,94-104; models/BidirGRU4Rec/model.py:51-58,83-90):
//   r = sig(W_ir x + b_ir + W_hr h + b_hr);  z = sig(W_iz x + b_iz + W_hz h + b_hz)
//   n = tanh(W_in x + b_in + r * (W_hn h + b_hn));  h' = (1-z) n + z h;  h0 = 0
// pack_padded_sequence(enforce_sorted=False) is folded in as a per-row length mask: the forward
// direction consumes tokens 0..len-1, the reverse direction len-1..0, and the final state is the
// state after exactly len tokens.  Only layer 0 is ever consumed by the heads (quirk q3).
#include "common.cuh"

struct GruWeights {
  const float *wiT[2], *whT[2];  // transposed [E][3H], [H][3H]
  const float *wi[2], *wh[2];    // original   [3H][E], [3H][H]
  const float *bi[2], *bh[2];
};

__device__ __forceinline__ int eff_len(const int64_t *lens, int b, int L, int packed) {
  if (!packed) return L;
  int64_t l = lens[b];
  return (int)(l < 1 ? 1 : (l > L ? L : l));
}

// ------------------------------------------------------------------------------------------------
// Forward.  grid = (ceil(B/R), dirs), block = 256.  One CTA carries R sessions through all steps.
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) gru_fwd_kernel(const float *__restrict__ emb, GruWeights w,
                                                      const int64_t *__restrict__ s,
                                                      const int64_t *__restrict__ lens, int B, int L, int E,
                                                      int H, int N, int packed, float *__restrict__ h_out,
                                                      float *__restrict__ gates_save,
                                                      float *__restrict__ hprev_save, int save) {
  extern __shared__ float smem[];
  const int dir = blockIdx.y, dirs = gridDim.y;
  const int b0 = blockIdx.x * R;
  const int G = 3 * H;
  float *xs = smem;                  // [R][E]
  float *hs = xs + R * E;            // [R][H]
  float *pre_i = hs + R * H;         // [R][3H]
  float *pre_h = pre_i + R * G;      // [R][3H]
  int *len_s = (int *)(pre_h + R * G);  // [R]
  int *tok_s = len_s + R;               // [R][L] item ids
  const int tid = threadIdx.x, NT = blockDim.x;

  if (tid < R) len_s[tid] = (b0 + tid < B) ? eff_len(lens, b0 + tid, L, packed) : 0;
  for (int i = tid; i < R * L; i += NT) {
    int r = i / L, t = i % L;
    int64_t it = (b0 + r < B) ? s[(int64_t)(b0 + r) * L + t] : 0;
    tok_s[i] = (int)(it < 0 ? 0 : (it > N ? N : it));
  }
  for (int i = tid; i < R * H; i += NT) hs[i] = 0.f;
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) maxlen = max(maxlen, len_s[r]);

  const float *__restrict__ wiT = w.wiT[dir];
  const float *__restrict__ whT = w.whT[dir];
  const float *__restrict__ bi = w.bi[dir];
  const float *__restrict__ bh = w.bh[dir];
  const int E4 = E >> 2;

  for (int i = 0; i < maxlen; ++i) {
    // gather x_t rows
    for (int idx = tid; idx < R * E4; idx += NT) {
      int r = idx / E4, c = idx - r * E4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < len_s[r]) {
        int tok = dir ? (len_s[r] - 1 - i) : i;
        v = __ldg(reinterpret_cast<const float4 *>(emb + (int64_t)tok_s[r * L + tok] * E) + c);
      }
      reinterpret_cast<float4 *>(xs)[idx] = v;
    }
    __syncthreads();
    // gate pre-activations: one gate row per thread, R sessions at once
    for (int j = tid; j < G; j += NT) {
      float ai[R], ah[R];
      const float bij = bi[j], bhj = bh[j];
#pragma unroll
      for (int r = 0; r < R; ++r) { ai[r] = bij; ah[r] = bhj; }
      for (int k = 0; k < E; k += 4) {
        float w0 = __ldg(wiT + (int64_t)(k + 0) * G + j), w1 = __ldg(wiT + (int64_t)(k + 1) * G + j);
        float w2 = __ldg(wiT + (int64_t)(k + 2) * G + j), w3 = __ldg(wiT + (int64_t)(k + 3) * G + j);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float4 x = *reinterpret_cast<const float4 *>(xs + r * E + k);
          ai[r] = fmaf(w0, x.x, ai[r]); ai[r] = fmaf(w1, x.y, ai[r]);
          ai[r] = fmaf(w2, x.z, ai[r]); ai[r] = fmaf(w3, x.w, ai[r]);
        }
      }
      for (int k = 0; k < H; k += 4) {
        float w0 = __ldg(whT + (int64_t)(k + 0) * G + j), w1 = __ldg(whT + (int64_t)(k + 1) * G + j);
        float w2 = __ldg(whT + (int64_t)(k + 2) * G + j), w3 = __ldg(whT + (int64_t)(k + 3) * G + j);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float4 x = *reinterpret_cast<const float4 *>(hs + r * H + k);
          ah[r] = fmaf(w0, x.x, ah[r]); ah[r] = fmaf(w1, x.y, ah[r]);
          ah[r] = fmaf(w2, x.z, ah[r]); ah[r] = fmaf(w3, x.w, ah[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) { pre_i[r * G + j] = ai[r]; pre_h[r * G + j] = ah[r]; }
    }
    __syncthreads();
    // combine gates, update state
    for (int idx = tid; idx < R * H; idx += NT) {
      int r = idx / H, u = idx - r * H;
      if (i < len_s[r]) {
        const float *pi = pre_i + r * G, *ph = pre_h + r * G;
        float rg = sigmoidf_(pi[u] + ph[u]);
        float zg = sigmoidf_(pi[H + u] + ph[H + u]);
        float phn = ph[2 * H + u];
        float ng = tanhf(pi[2 * H + u] + rg * phn);
        float hold = hs[idx];
        float hnew = (1.f - zg) * ng + zg * hold;
        if (save) {
          int tok = dir ? (len_s[r] - 1 - i) : i;
          int64_t base = ((int64_t)(b0 + r) * L + tok) * dirs + dir;
          float *g = gates_save + base * 4 * H;
          g[u] = rg; g[H + u] = zg; g[2 * H + u] = ng; g[3 * H + u] = phn;
          hprev_save[base * H + u] = hold;
        }
        hs[idx] = hnew;
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < R * H; idx += NT) {
    int r = idx / H, u = idx - r * H;
    if (b0 + r < B) h_out[(int64_t)(b0 + r) * (dirs * H) + dir * H + u] = hs[idx];
  }
}

// ------------------------------------------------------------------------------------------------
// BPTT.  grid = (ceil(B/R), dirs), block = 256.  Writes d(pre-activations) per token and dx.
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) gru_bwd_kernel(GruWeights w, const int64_t *__restrict__ lens, int B,
                                                      int L, int E, int H, int packed,
                                                      const float *__restrict__ dh_in,
                                                      const float *__restrict__ gates_save,
                                                      const float *__restrict__ hprev_save,
                                                      float *__restrict__ dgi, float *__restrict__ dgh,
                                                      float *__restrict__ dx) {
  extern __shared__ float smem[];
  const int dir = blockIdx.y, dirs = gridDim.y;
  const int b0 = blockIdx.x * R;
  const int G = 3 * H;
  float *dhs = smem;              // [R][H] running dL/dh
  float *dhd = dhs + R * H;       // [R][H] direct (z-gated) part
  float *dai = dhd + R * H;       // [R][3H]
  float *dah = dai + R * G;       // [R][3H]
  int *len_s = (int *)(dah + R * G);
  const int tid = threadIdx.x, NT = blockDim.x;
  if (tid < R) len_s[tid] = (b0 + tid < B) ? eff_len(lens, b0 + tid, L, packed) : 0;
  for (int idx = tid; idx < R * H; idx += NT) {
    int r = idx / H, u = idx - r * H;
    dhs[idx] = (b0 + r < B) ? dh_in[(int64_t)(b0 + r) * (dirs * H) + dir * H + u] : 0.f;
  }
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) maxlen = max(maxlen, len_s[r]);
  const float *__restrict__ wi = w.wi[dir];
  const float *__restrict__ wh = w.wh[dir];

  for (int i = maxlen - 1; i >= 0; --i) {
    for (int idx = tid; idx < R * H; idx += NT) {
      int r = idx / H, u = idx - r * H;
      float a_r = 0.f, a_z = 0.f, a_n = 0.f, h_n = 0.f, direct = dhs[idx];
      if (i < len_s[r]) {
        int tok = dir ? (len_s[r] - 1 - i) : i;
        int64_t base = ((int64_t)(b0 + r) * L + tok) * dirs + dir;
        const float *g = gates_save + base * 4 * H;
        float rg = g[u], zg = g[H + u], ng = g[2 * H + u], phn = g[3 * H + u];
        float hp = hprev_save[base * H + u];
        float dhv = dhs[idx];
        float dn = dhv * (1.f - zg);
        float dz = dhv * (hp - ng);
        a_n = dn * (1.f - ng * ng);
        float dr = a_n * phn;
        a_r = dr * rg * (1.f - rg);
        a_z = dz * zg * (1.f - zg);
        h_n = a_n * rg;
        direct = dhv * zg;
        float *gi = dgi + base * G, *gh = dgh + base * G;
        gi[u] = a_r; gi[H + u] = a_z; gi[2 * H + u] = a_n;
        gh[u] = a_r; gh[H + u] = a_z; gh[2 * H + u] = h_n;
      }
      dai[r * G + u] = a_r; dai[r * G + H + u] = a_z; dai[r * G + 2 * H + u] = a_n;
      dah[r * G + u] = a_r; dah[r * G + H + u] = a_z; dah[r * G + 2 * H + u] = h_n;
      dhd[idx] = direct;
    }
    __syncthreads();
    for (int c = tid; c < H + E; c += NT) {
      float acc[R];
      if (c < H) {
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = dhd[r * H + c];
        for (int j = 0; j < G; ++j) {
          float wv = __ldg(wh + (int64_t)j * H + c);
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = fmaf(dah[r * G + j], wv, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) dhs[r * H + c] = acc[r];  // inactive rows: dah = 0 -> unchanged
      } else {
        int ce = c - H;
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.f;
        for (int j = 0; j < G; ++j) {
          float wv = __ldg(wi + (int64_t)j * E + ce);
#pragma unroll
          for (int r = 0; r < R; ++r) acc[r] = fmaf(dai[r * G + j], wv, acc[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (i < len_s[r]) {
            int tok = dir ? (len_s[r] - 1 - i) : i;
            dx[(((int64_t)(b0 + r) * L + tok) * dirs + dir) * E + ce] = acc[r];
          }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Weight gradients: dW_ih = sum_p dgi[p]^T x[p], dW_hh = sum_p dgh[p]^T hprev[p], biases = column
// sums.  Split over token positions; partials reduced in a fixed order by gru_adam_kernel.
// grid = (ceil(3H/64), ceil(max(E,H)/64), dirs*2*splits), block = 256 (16x16, 4x4 micro tile).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gru_wgrad_kernel(const float *__restrict__ emb,
                                                        const int64_t *__restrict__ s,
                                                        const int64_t *__restrict__ lens, int B, int L, int E,
                                                        int H, int N, int packed, int dirs, int splits,
                                                        const float *__restrict__ dgi,
                                                        const float *__restrict__ dgh,
                                                        const float *__restrict__ hprev_save,
                                                        float *__restrict__ part, int KS) {
  __shared__ float As[16][64];
  __shared__ float Bs[16][64];
  const int G = 3 * H;
  int z = blockIdx.z;
  const int split = z % splits; z /= splits;
  const int which = z & 1;  // 0: ih (x), 1: hh (hprev)
  const int dir = z >> 1;
  const int K = which ? H : E;
  const int j0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  if (k0 >= K) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int P = B * L;
  const int per = (P + splits - 1) / splits;
  const int p_lo = split * per, p_hi = min(P, p_lo + per);
  float acc[4][4];
  float bacc[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) { bacc[a] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f; }
  const float *A = which ? dgh : dgi;
  for (int p0 = p_lo; p0 < p_hi; p0 += 16) {
    // stage 16 positions x 64 columns of A and of the right operand
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int e = tid + q * 256;
      int pp = e >> 6, col = e & 63;
      int p = p0 + pp;
      float av = 0.f, bv = 0.f;
      if (p < p_hi) {
        int b = p / L, t = p - b * L;
        if (t < eff_len(lens, b, L, packed)) {
          int64_t base = (int64_t)p * dirs + dir;
          if (j0 + col < G) av = A[base * G + j0 + col];
          if (k0 + col < K) {
            if (which) bv = hprev_save[base * H + k0 + col];
            else {
              int64_t it = s[p];
              it = it < 0 ? 0 : (it > N ? N : it);
              bv = __ldg(emb + it * E + k0 + col);
            }
          }
        }
      }
      As[pp][col] = av;
      Bs[pp][col] = bv;
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < 16; ++pp) {
      float4 a4 = *reinterpret_cast<const float4 *>(&As[pp][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4 *>(&Bs[pp][tx * 4]);
      float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) { bacc[a] += av[a];
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]); }
    }
    __syncthreads();
  }
  // part layout: [split][dir][which][3H][KS], bias at column KS-1
  float *out = part + ((((int64_t)split * dirs + dir) * 2 + which) * G) * KS;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int j = j0 + ty * 4 + a;
    if (j >= G) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      int k = k0 + tx * 4 + c;
      if (k < K) out[(int64_t)j * KS + k] = acc[a][c];
    }
    if (blockIdx.y == 0 && tx == 0) out[(int64_t)j * KS + KS - 1] = bacc[a];
  }
}

__device__ __forceinline__ void adam_update(float &p, float &m, float &v, float g, float b1, float b2,
                                            float eps, float step_size, float bc2_sqrt) {
  m = m + (g - m) * (1.f - b1);
  v = v * b2 + ((1.f - b2) * g) * g;
  float denom = sqrtf(v) / bc2_sqrt + eps;
  p = p + (-step_size * m) / denom;
}

struct GruAdamPtrs {
  float *p[2][4], *m[2][4], *v[2][4];  // [dir][w_ih, w_hh, b_ih, b_hh]
  float *wiT[2], *whT[2];
};

// One thread per GRU parameter: reduce the split partials in order, Adam, refresh transposes.
__global__ void gru_adam_kernel(GruAdamPtrs q, const float *__restrict__ part, int splits, int dirs, int E,
                                int H, int KS, float b1, float b2, float eps, float step_size,
                                float bc2_sqrt) {
  const int G = 3 * H;
  const int per_dir = G * E + G * H + 2 * G;
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= per_dir * dirs) return;
  int dir = gid / per_dir, o = gid - dir * per_dir;
  int which, j, k, t;
  if (o < G * E) { t = 0; which = 0; j = o / E; k = o - j * E; }
  else if (o < G * E + G * H) { o -= G * E; t = 1; which = 1; j = o / H; k = o - j * H; }
  else if (o < G * E + G * H + G) { o -= G * E + G * H; t = 2; which = 0; j = o; k = KS - 1; }
  else { o -= G * E + G * H + G; t = 3; which = 1; j = o; k = KS - 1; }
  float g = 0.f;
  for (int sidx = 0; sidx < splits; ++sidx)
    g += part[((((int64_t)sidx * dirs + dir) * 2 + which) * G + j) * KS + k];
  int64_t off = (t == 0) ? (int64_t)j * E + k : (t == 1) ? (int64_t)j * H + k : j;
  float p = q.p[dir][t][off], m = q.m[dir][t][off], v = q.v[dir][t][off];
  adam_update(p, m, v, g, b1, b2, eps, step_size, bc2_sqrt);
  q.p[dir][t][off] = p; q.m[dir][t][off] = m; q.v[dir][t][off] = v;
  if (t == 0) q.wiT[dir][(int64_t)k * G + j] = p;
  if (t == 1) q.whT[dir][(int64_t)k * G + j] = p;
}

__global__ void transpose_kernel(const float *__restrict__ in, float *__restrict__ out, int rows, int cols) {
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= rows * cols) return;
  int r = gid / cols, c = gid - r * cols;
  out[(int64_t)c * rows + r] = in[gid];
}

// ------------------------------------------------------------------------------------------------
static GruWeights gru_weights(const rec_engine *e, int net_id) {
  GruWeights w;
  const NetBind &nb = e->nets[net_id];
  for (int d = 0; d < 2; ++d) {
    w.wiT[d] = nb.w_ihT[d]; w.whT[d] = nb.w_hhT[d];
    w.wi[d] = nb.p.w_ih[d]; w.wh[d] = nb.p.w_hh[d];
    w.bi[d] = nb.p.b_ih[d]; w.bh[d] = nb.p.b_hh[d];
  }
  return w;
}

static const int GRU_R = 4;

int launch_gru_transpose(rec_engine *e, int net_id) {
  const rec_config &c = e->cfg;
  NetBind &nb = e->nets[net_id];
  const int G = 3 * c.hidden_dim;
  for (int d = 0; d < e->dirs; ++d) {
    int n1 = G * c.embedding_dim, n2 = G * c.hidden_dim;
    transpose_kernel<<<cdiv(n1, 256), 256, 0, e->stream>>>(nb.p.w_ih[d], nb.w_ihT[d], G, c.embedding_dim);
    REC_LAUNCH_CHECK(e);
    transpose_kernel<<<cdiv(n2, 256), 256, 0, e->stream>>>(nb.p.w_hh[d], nb.w_hhT[d], G, c.hidden_dim);
    REC_LAUNCH_CHECK(e);
  }
  return REC_OK;
}

int launch_gru_forward(rec_engine *e, int net_id, const int64_t *s, const int64_t *lengths, int B,
                       float *h_out, bool save) {
  const rec_config &c = e->cfg;
  const int E = c.embedding_dim, H = c.hidden_dim, L = c.state_size;
  size_t smem = (size_t)GRU_R * (E + H + 6 * H) * sizeof(float) + (size_t)GRU_R * (1 + L) * sizeof(int);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    REC_CUDA(e, cudaFuncSetAttribute(gru_fwd_kernel<GRU_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  if (smem > 200 * 1024) REC_FAIL(e, REC_EINVAL, "GRU forward needs %zu B of shared memory (E=%d,H=%d)", smem, E, H);
  dim3 grid(cdiv(B, GRU_R), e->dirs);
  gru_fwd_kernel<GRU_R><<<grid, 256, smem, e->stream>>>(e->nets[net_id].p.emb, gru_weights(e, net_id), s, lengths,
                                                       B, L, E, H, c.item_num, c.use_packed_seq, h_out,
                                                       e->gates_save, e->hprev_save, save ? 1 : 0);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

int launch_gru_backward(rec_engine *e, int net_id, const int64_t *s, const int64_t *lengths, int B,
                        const float *dh, float step_size, float bc2_sqrt, const rec_train_hparams *hp) {
  const rec_config &c = e->cfg;
  const int E = c.embedding_dim, H = c.hidden_dim, L = c.state_size, G = 3 * H;
  size_t smem = (size_t)GRU_R * (2 * H + 2 * G) * sizeof(float) + GRU_R * sizeof(int);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    REC_CUDA(e, cudaFuncSetAttribute(gru_bwd_kernel<GRU_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid(cdiv(B, GRU_R), e->dirs);
  gru_bwd_kernel<GRU_R><<<grid, 256, smem, e->stream>>>(gru_weights(e, net_id), lengths, B, L, E, H,
                                                       c.use_packed_seq, dh, e->gates_save, e->hprev_save,
                                                       e->dgi, e->dgh, e->dx);
  REC_LAUNCH_CHECK(e);
  // weight gradients (split over token positions) + Adam on the GRU parameters
  const int KS = (E > H ? E : H) + 1;
  const int splits = e->wgrad_splits;
  dim3 g2(cdiv(G, 64), cdiv(E > H ? E : H, 64), e->dirs * 2 * splits);
  gru_wgrad_kernel<<<g2, 256, 0, e->stream>>>(e->nets[net_id].p.emb, s, lengths, B, L, E, H, c.item_num,
                                             c.use_packed_seq, e->dirs, splits, e->dgi, e->dgh, e->hprev_save,
                                             e->wgrad_part, KS);
  REC_LAUNCH_CHECK(e);
  NetBind &nb = e->nets[net_id];
  GruAdamPtrs q;
  for (int d = 0; d < 2; ++d) {
    q.p[d][0] = nb.p.w_ih[d]; q.m[d][0] = nb.p.w_ih_m[d]; q.v[d][0] = nb.p.w_ih_v[d];
    q.p[d][1] = nb.p.w_hh[d]; q.m[d][1] = nb.p.w_hh_m[d]; q.v[d][1] = nb.p.w_hh_v[d];
    q.p[d][2] = nb.p.b_ih[d]; q.m[d][2] = nb.p.b_ih_m[d]; q.v[d][2] = nb.p.b_ih_v[d];
    q.p[d][3] = nb.p.b_hh[d]; q.m[d][3] = nb.p.b_hh_m[d]; q.v[d][3] = nb.p.b_hh_v[d];
    q.wiT[d] = nb.w_ihT[d]; q.whT[d] = nb.w_hhT[d];
  }
  int total = (G * E + G * H + 2 * G) * e->dirs;
  gru_adam_kernel<<<cdiv(total, 256), 256, 0, e->stream>>>(q, e->wgrad_part, splits, e->dirs, E, H, KS, hp->beta1,
                                                          hp->beta2, hp->eps, step_size, bc2_sqrt);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}
