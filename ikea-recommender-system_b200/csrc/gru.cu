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
// Fast forward path for E = H = 64 (every headline config).  192 threads = one gate row each; the
// thread keeps its rows of W_ih and W_hh (2 x 64 floats) in REGISTERS for the whole sequence, x_t and
// h live in shared memory and are read as broadcast float4.  Up to 3 independent passes (main(s),
// main(s'), boot(s')) run as one launch: grid = (ceil(B/2), dirs, n_pass).
// ------------------------------------------------------------------------------------------------
struct GruPass {
  const float *emb;
  GruWeights w;
  const int64_t *s, *lens;
  float *h_out;
  int save;
};
struct GruPasses { GruPass p[3]; };

#define FR 3   // sessions per CTA in the forward fast path (FR * 64 == 192 threads in the gate-combine phase)
#define BR 2   // sessions per CTA in the BPTT fast path
__global__ void __launch_bounds__(192, 2) gru_fwd64_kernel(GruPasses ps, int B, int L, int N, int packed,
                                                           float *__restrict__ gates_save,
                                                           float *__restrict__ hprev_save) {
  constexpr int H = 64, E = 64, G = 192;
  const GruPass &P = ps.p[blockIdx.z];
  const int dir = blockIdx.y, dirs = gridDim.y;
  const int b0 = blockIdx.x * FR;
  const int tid = threadIdx.x;
  __shared__ __align__(16) float xs[FR][E];
  __shared__ __align__(16) float hs[FR][H];
  __shared__ float pre_i[FR][G];
  __shared__ float pre_h[FR][G];
  __shared__ int len_s[FR];
  __shared__ int tok_s[FR][64];  // L <= 64 on this path
  __shared__ __align__(16) float xall[FR][16][E];  // L <= 16: every token row of the CTA's sessions (12 KB)

  float wi[E], wh[H];
  {
    const float4 *ri = reinterpret_cast<const float4 *>(P.w.wi[dir] + (int64_t)tid * E);
    const float4 *rh = reinterpret_cast<const float4 *>(P.w.wh[dir] + (int64_t)tid * H);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float4 a = __ldg(ri + k), b = __ldg(rh + k);
      wi[4 * k] = a.x; wi[4 * k + 1] = a.y; wi[4 * k + 2] = a.z; wi[4 * k + 3] = a.w;
      wh[4 * k] = b.x; wh[4 * k + 1] = b.y; wh[4 * k + 2] = b.z; wh[4 * k + 3] = b.w;
    }
  }
  const float bij = P.w.bi[dir][tid], bhj = P.w.bh[dir][tid];
  if (tid < FR) len_s[tid] = (b0 + tid < B) ? eff_len(P.lens, b0 + tid, L, packed) : 0;
  for (int i = tid; i < FR * L; i += 192) {
    int r = i / L, t = i - r * L;
    int64_t it = (b0 + r < B) ? P.s[(int64_t)(b0 + r) * L + t] : 0;
    tok_s[r][t] = (int)(it < 0 ? 0 : (it > N ? N : it));
  }
  if (tid < FR * H) hs[tid / H][tid % H] = 0.f;
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int r = 0; r < FR; ++r) maxlen = max(maxlen, len_s[r]);
  // x loader: threads [0, FR*16) own one float4 of one row
  const int xr = tid >> 4, xc = tid & 15;
  auto load_x = [&](int i) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < FR * 16 && i < len_s[xr]) {
      int tok = dir ? (len_s[xr] - 1 - i) : i;
      v = __ldg(reinterpret_cast<const float4 *>(P.emb + (int64_t)tok_s[xr][tok] * E) + xc);
    }
    return v;
  };
  // L <= 16 (every headline config: L = 10): ALL token rows of the CTA's sessions are requested up front (one memory
  // round trip, overlapped with the weight loads above) instead of one dependent gather per time step
  const bool pre = L <= 16;
  if (pre) {
    for (int i = tid; i < FR * L * 16; i += 192) {
      const int r = i / (L * 16), rem = i - r * L * 16, t = rem >> 4, c = rem & 15;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < len_s[r]) v = __ldg(reinterpret_cast<const float4 *>(P.emb + (int64_t)tok_s[r][t] * E) + c);
      reinterpret_cast<float4 *>(&xall[r][t][0])[c] = v;
    }
  } else if (tid < FR * 16) {
    reinterpret_cast<float4 *>(&xs[xr][0])[xc] = load_x(0);
  }
  __syncthreads();
  for (int i = 0; i < maxlen; ++i) {
    float4 xnext = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!pre) xnext = load_x(i + 1);  // prefetch, consumed after the gate phase
    const float4 *xrow[FR];
#pragma unroll
    for (int r = 0; r < FR; ++r)
      xrow[r] = pre ? reinterpret_cast<const float4 *>(&xall[r][dir ? max(len_s[r] - 1 - i, 0) : i][0])
                    : reinterpret_cast<const float4 *>(&xs[r][0]);
    float ai[FR][2], ah[FR][2];
#pragma unroll
    for (int r = 0; r < FR; ++r) { ai[r][0] = bij; ai[r][1] = 0.f; ah[r][0] = bhj; ah[r][1] = 0.f; }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
#pragma unroll
      for (int r = 0; r < FR; ++r) {
        float4 x = xrow[r][k];
        float4 hv = reinterpret_cast<const float4 *>(&hs[r][0])[k];
        ai[r][0] = fmaf(wi[4 * k], x.x, ai[r][0]); ai[r][1] = fmaf(wi[4 * k + 1], x.y, ai[r][1]);
        ai[r][0] = fmaf(wi[4 * k + 2], x.z, ai[r][0]); ai[r][1] = fmaf(wi[4 * k + 3], x.w, ai[r][1]);
        ah[r][0] = fmaf(wh[4 * k], hv.x, ah[r][0]); ah[r][1] = fmaf(wh[4 * k + 1], hv.y, ah[r][1]);
        ah[r][0] = fmaf(wh[4 * k + 2], hv.z, ah[r][0]); ah[r][1] = fmaf(wh[4 * k + 3], hv.w, ah[r][1]);
      }
    }
#pragma unroll
    for (int r = 0; r < FR; ++r) { pre_i[r][tid] = ai[r][0] + ai[r][1]; pre_h[r][tid] = ah[r][0] + ah[r][1]; }
    __syncthreads();
    if (tid < FR * H) {
      const int r = tid >> 6, u = tid & 63;
      if (i < len_s[r]) {
        float rg = sigmoidf_(pre_i[r][u] + pre_h[r][u]);
        float zg = sigmoidf_(pre_i[r][H + u] + pre_h[r][H + u]);
        float phn = pre_h[r][2 * H + u];
        float ng = tanhf(pre_i[r][2 * H + u] + rg * phn);
        float hold = hs[r][u];
        if (P.save) {
          int tok = dir ? (len_s[r] - 1 - i) : i;
          int64_t base = ((int64_t)(b0 + r) * L + tok) * dirs + dir;
          float *g = gates_save + base * 4 * H;
          g[u] = rg; g[H + u] = zg; g[2 * H + u] = ng; g[3 * H + u] = phn;
          hprev_save[base * H + u] = hold;
        }
        hs[r][u] = (1.f - zg) * ng + zg * hold;
      }
    }
    if (!pre && tid < FR * 16) reinterpret_cast<float4 *>(&xs[xr][0])[xc] = xnext;
    __syncthreads();
  }
  if (tid < FR * H) {
    const int r = tid >> 6, u = tid & 63;
    if (b0 + r < B) P.h_out[(int64_t)(b0 + r) * (dirs * H) + dir * H + u] = hs[r][u];
  }
}

// Fast BPTT for E = H = 64: 128 threads; thread c < 64 owns COLUMN c of W_hh (192 registers) and
// produces dh_prev[c]; thread 64 + c owns column c of W_ih and produces dx[c].  grid = (ceil(B/2), dirs).
__global__ void __launch_bounds__(128, 2) gru_bwd64_kernel(GruWeights w, const int64_t *__restrict__ lens, int B,
                                                           int L, int packed, const float *__restrict__ dh_in,
                                                           const float *__restrict__ gates_save,
                                                           const float *__restrict__ hprev_save,
                                                           float *__restrict__ dgi, float *__restrict__ dgh,
                                                           float *__restrict__ dx) {
  constexpr int H = 64, E = 64, G = 192;
  const int dir = blockIdx.y, dirs = gridDim.y;
  const int b0 = blockIdx.x * BR;
  const int tid = threadIdx.x;
  __shared__ float dhs[BR][H];
  __shared__ float dhd[BR][H];
  __shared__ __align__(16) float dai[BR][G];
  __shared__ __align__(16) float dah[BR][G];
  __shared__ int len_s[BR];
  const bool is_h = tid < 64;
  const int c = tid & 63;
  float wc[G];
  {
    const float *src = is_h ? (w.wh[dir] + c) : (w.wi[dir] + c);
#pragma unroll
    for (int j = 0; j < G; ++j) wc[j] = __ldg(src + (int64_t)j * 64);
  }
  if (tid < BR) len_s[tid] = (b0 + tid < B) ? eff_len(lens, b0 + tid, L, packed) : 0;
  {
    const int r = tid >> 6, u = tid & 63;
    dhs[r][u] = (b0 + r < B) ? dh_in[(int64_t)(b0 + r) * (dirs * H) + dir * H + u] : 0.f;
  }
  __syncthreads();
  const int maxlen = max(len_s[0], len_s[1]);
  for (int i = maxlen - 1; i >= 0; --i) {
    {
      const int r = tid >> 6, u = tid & 63;
      float a_r = 0.f, a_z = 0.f, a_n = 0.f, h_n = 0.f, direct = dhs[r][u];
      if (i < len_s[r]) {
        int tok = dir ? (len_s[r] - 1 - i) : i;
        int64_t base = ((int64_t)(b0 + r) * L + tok) * dirs + dir;
        const float *g = gates_save + base * 4 * H;
        float rg = g[u], zg = g[H + u], ng = g[2 * H + u], phn = g[3 * H + u];
        float hp = hprev_save[base * H + u];
        float dhv = direct;
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
      dai[r][u] = a_r; dai[r][H + u] = a_z; dai[r][2 * H + u] = a_n;
      dah[r][u] = a_r; dah[r][H + u] = a_z; dah[r][2 * H + u] = h_n;
      dhd[r][u] = direct;
    }
    __syncthreads();
    {
      float acc[BR][2];
#pragma unroll
      for (int r = 0; r < BR; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }
#pragma unroll
      for (int j4 = 0; j4 < G / 4; ++j4) {
#pragma unroll
        for (int r = 0; r < BR; ++r) {
          float4 d = is_h ? reinterpret_cast<const float4 *>(&dah[r][0])[j4]
                          : reinterpret_cast<const float4 *>(&dai[r][0])[j4];
          acc[r][0] = fmaf(d.x, wc[4 * j4], acc[r][0]); acc[r][1] = fmaf(d.y, wc[4 * j4 + 1], acc[r][1]);
          acc[r][0] = fmaf(d.z, wc[4 * j4 + 2], acc[r][0]); acc[r][1] = fmaf(d.w, wc[4 * j4 + 3], acc[r][1]);
        }
      }
      // dhs is only read in the elementwise phase above (already past the barrier): safe to update
#pragma unroll
      for (int r = 0; r < BR; ++r) {
        float v = acc[r][0] + acc[r][1];
        if (is_h) {
          dhs[r][c] = dhd[r][c] + v;  // inactive rows: dah = 0 -> dhs unchanged
        } else if (i < len_s[r]) {
          int tok = dir ? (len_s[r] - 1 - i) : i;
          dx[(((int64_t)(b0 + r) * L + tok) * dirs + dir) * E + c] = v;
        }
      }
    }
    __syncthreads();
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
  adam_elem(p, m, v, g, b1, b2, eps, step_size, 1.f / bc2_sqrt);
}

struct GruAdamPtrs {
  float *p[2][4], *m[2][4], *v[2][4];  // [dir][w_ih, w_hh, b_ih, b_hh]
  float *wiT[2], *whT[2];
};

// 64 GRU parameters per block x 4 split groups: partials are summed in a fixed order (group g takes splits
// g, g+4, ...; groups are combined in order), then Adam + refresh of the transposed copies.
__global__ void __launch_bounds__(256) gru_adam_kernel(GruAdamPtrs q, const float *__restrict__ part, int splits, int dirs,
                                                       int E, int H, int KS, float b1, float b2, float eps,
                                                       float step_size, float bc2_sqrt, const float *__restrict__ sc) {
  if (sc) { step_size = sc[0]; bc2_sqrt = 1.f / sc[1]; }
  __shared__ float sh[4][64];
  const int G = 3 * H;
  const int per_dir = G * E + G * H + 2 * G;
  const int o = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int gid = blockIdx.x * 64 + o;
  const bool valid = gid < per_dir * dirs;
  int dir = 0, which = 0, j = 0, k = 0, t = 0;
  if (valid) {
    dir = gid / per_dir;
    int oo = gid - dir * per_dir;
    if (oo < G * E) { t = 0; which = 0; j = oo / E; k = oo - j * E; }
    else if (oo < G * E + G * H) { oo -= G * E; t = 1; which = 1; j = oo / H; k = oo - j * H; }
    else if (oo < G * E + G * H + G) { oo -= G * E + G * H; t = 2; which = 0; j = oo; k = KS - 1; }
    else { oo -= G * E + G * H + G; t = 3; which = 1; j = oo; k = KS - 1; }
  }
  float acc = 0.f;
  if (valid)
    for (int sidx = grp; sidx < splits; sidx += 4)
      acc += part[((((int64_t)sidx * dirs + dir) * 2 + which) * G + j) * KS + k];
  sh[grp][o] = acc;
  __syncthreads();
  if (grp != 0 || !valid) return;
  const float g = ((sh[0][o] + sh[1][o]) + sh[2][o]) + sh[3][o];
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

static GruPass make_pass(rec_engine *e, int net_id, const int64_t *s, const int64_t *lengths, float *h_out, bool save) {
  GruPass p;
  p.emb = e->nets[net_id].p.emb; p.w = gru_weights(e, net_id); p.s = s; p.lens = lengths; p.h_out = h_out;
  p.save = save ? 1 : 0;
  return p;
}

static bool gru_fast_path(const rec_engine *e) {
  return e->cfg.embedding_dim == 64 && e->cfg.hidden_dim == 64 && e->cfg.state_size <= 64;
}

// Up to three independent passes in one launch (fast path) or back-to-back launches (generic path).
int launch_gru_forward_multi(rec_engine *e, int n_pass, const int *net_ids, const int64_t *const *s,
                             const int64_t *const *lengths, float *const *h_out, const bool *save, int B) {
  const rec_config &c = e->cfg;
  if (gru_fast_path(e)) {
    GruPasses ps;
    for (int i = 0; i < 3; ++i) ps.p[i] = make_pass(e, net_ids[i < n_pass ? i : 0], s[i < n_pass ? i : 0],
                                                    lengths[i < n_pass ? i : 0], h_out[i < n_pass ? i : 0],
                                                    save[i < n_pass ? i : 0]);
    dim3 grid(cdiv(B, FR), e->dirs, n_pass);
    gru_fwd64_kernel<<<grid, 192, 0, e->stream>>>(ps, B, c.state_size, c.item_num, c.use_packed_seq, e->gates_save,
                                                 e->hprev_save);
    REC_LAUNCH_CHECK(e);
    return REC_OK;
  }
  if (gru_tc_supported(e)) return launch_gru_forward_tc(e, n_pass, net_ids, s, lengths, h_out, save, B);
  for (int i = 0; i < n_pass; ++i) {
    int rc = launch_gru_forward(e, net_ids[i], s[i], lengths[i], B, h_out[i], save[i]);
    if (rc) return rc;
  }
  return REC_OK;
}

int launch_gru_forward(rec_engine *e, int net_id, const int64_t *s, const int64_t *lengths, int B,
                       float *h_out, bool save) {
  const rec_config &c = e->cfg;
  if (gru_fast_path(e) || gru_tc_supported(e)) {
    const int64_t *sa[1] = {s}, *la[1] = {lengths};
    float *ha[1] = {h_out};
    return launch_gru_forward_multi(e, 1, &net_id, sa, la, ha, &save, B);
  }
  const int E = c.embedding_dim, H = c.hidden_dim, L = c.state_size;
  size_t smem = (size_t)GRU_R * (E + H + 6 * H) * sizeof(float) + (size_t)GRU_R * (1 + L) * sizeof(int);
  static bool attr_set[REC_MAX_DEVICES] = {};  // per device: the opt-in is a per-device function attribute
  if (!attr_set[e->dev] && smem > 48 * 1024) {
    REC_CUDA(e, cudaFuncSetAttribute(gru_fwd_kernel<GRU_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set[e->dev] = true;
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
                        const float *dh, float step_size, float bc2_sqrt, const rec_train_hparams *hp, int stages) {
  const rec_config &c = e->cfg;
  const int E = c.embedding_dim, H = c.hidden_dim, L = c.state_size, G = 3 * H;
  size_t smem = (size_t)GRU_R * (2 * H + 2 * G) * sizeof(float) + GRU_R * sizeof(int);
  static bool attr_set[REC_MAX_DEVICES] = {};  // per device: the opt-in is a per-device function attribute
  if (!attr_set[e->dev] && smem > 48 * 1024) {
    REC_CUDA(e, cudaFuncSetAttribute(gru_bwd_kernel<GRU_R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set[e->dev] = true;
  }
  if (!gru_fast_path(e) && gru_tc_supported(e)) {
    if (stages & 3) {
      int rc = launch_gru_backward_tc(e, net_id, s, lengths, B, dh, stages & 3);
      if (rc) return rc;
    }
    stages &= 4;
  }
  if (stages & 1) {
    if (gru_fast_path(e)) {
      dim3 grid(cdiv(B, BR), e->dirs);
      gru_bwd64_kernel<<<grid, 128, 0, e->stream>>>(gru_weights(e, net_id), lengths, B, L, c.use_packed_seq, dh,
                                                   e->gates_save, e->hprev_save, e->dgi, e->dgh, e->dx);
    } else {
      dim3 grid(cdiv(B, GRU_R), e->dirs);
      gru_bwd_kernel<GRU_R><<<grid, 256, smem, e->stream>>>(gru_weights(e, net_id), lengths, B, L, E, H,
                                                           c.use_packed_seq, dh, e->gates_save, e->hprev_save,
                                                           e->dgi, e->dgh, e->dx);
    }
    REC_LAUNCH_CHECK(e);
  }
  // weight gradients (split over token positions) + Adam on the GRU parameters
  const int KS = (E > H ? E : H) + 1;
  if (stages & 2) e->wgrad_used = e->wgrad_splits;
  const int splits = e->wgrad_splits;
  if (stages & 2) {
    dim3 g2(cdiv(G, 64), cdiv(E > H ? E : H, 64), e->dirs * 2 * splits);
    gru_wgrad_kernel<<<g2, 256, 0, e->stream>>>(e->nets[net_id].p.emb, s, lengths, B, L, E, H, c.item_num,
                                               c.use_packed_seq, e->dirs, splits, e->dgi, e->dgh, e->hprev_save,
                                               e->wgrad_part, KS);
    REC_LAUNCH_CHECK(e);
  }
  if (!(stages & 4)) return REC_OK;
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
  gru_adam_kernel<<<cdiv(total, 64), 256, 0, e->stream>>>(q, e->wgrad_part, e->wgrad_used, e->dirs, E, H, KS, hp->beta1,
                                                          hp->beta2, hp->eps, step_size, bc2_sqrt, e->d_sc);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}
