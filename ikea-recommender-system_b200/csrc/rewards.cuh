// Reward helpers shared by the TD kernels (metrics.cu) and the fused per-row Q kernel (heads.cu).
// Reference: evaluate/diversity.py:15-73 (cosine diversity, eps 1e-6), models/SMORL/smorl_gru.py:291-325.
#pragma once
#include "common.cuh"

__device__ __forceinline__ int last_action_of(const int64_t *s, const int64_t *lens, int b, int L, int N,
                                              int pad_pos_end) {
  int64_t it;
  if (pad_pos_end) {
    int64_t l = lens[b];
    l = l < 1 ? 1 : (l > L ? L : l);
    it = s[(int64_t)b * L + (l - 1)];
  } else {
    it = s[(int64_t)b * L + (L - 1)];
  }
  return (int)(it < 0 ? 0 : (it > N ? N : it));
}

// 1 - mean_j cos(E[last], E[map(id_j)]), j < k   (CosineSimilarity(dim=2, eps=1e-6)); warp-cooperative
__device__ __forceinline__ float diversity_reward_warp(const float *__restrict__ E, int dim, int last,
                                                       const int32_t *ids, int k, const int64_t *out_to_in,
                                                       int N, int lane, int V = 0x7fffffff) {
  const float eps = 1e-6f;
  const float *x = E + (int64_t)last * dim;
  float nx = 0.f;
  for (int d = lane; d < dim; d += 32) nx = fmaf(x[d], x[d], nx);
  nx = fmaxf(sqrtf(warp_sum(nx)), eps);
  float sim_sum = 0.f;
  for (int j = 0; j < k; ++j) {
    int64_t id = ids[j];
    if (id < 0 || id >= V) continue;  // exhausted top-k slot (k > V): contributes cos = 0
    if (out_to_in) id = out_to_in[id];
    id = id < 0 ? 0 : (id > N ? N : id);
    const float *y = E + id * dim;
    float ny = 0.f, dot = 0.f;
    for (int d = lane; d < dim; d += 32) { ny = fmaf(y[d], y[d], ny); dot = fmaf(x[d], y[d], dot); }
    ny = fmaxf(sqrtf(warp_sum(ny)), eps);
    dot = warp_sum(dot);
    sim_sum += dot / (nx * ny);
  }
  return 1.f - sim_sum / (float)k;
}

