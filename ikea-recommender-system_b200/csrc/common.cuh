// Shared device helpers + the engine structure behind the opaque rec_engine handle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/recsys_b200.h"

#define REC_NEG_INF (-3.402823466e+38f)
#define REC_MAX_DEVICES 64

struct NetBind {
  rec_net_params p;
  bool bound;
  int64_t adam_step;
  // engine-owned transposed copies of the GRU weights: WiT[E][3H], WhT[H][3H] per direction
  float *w_ihT[2], *w_hhT[2];
};

// Per-pass GRU buffers: which (net, sequence, lengths) produced which final state.
struct rec_engine {
  rec_config cfg;
  int dev;               // CUDA device the engine lives on (every entry point switches to it: DevGuard)
  cudaStream_t stream;
  char err[512];
  NetBind nets[REC_MAX_NETS];
  int D, dirs, Vloc;
  int sm_count;
  int64_t launches;

  // ---- workspace (device) ----
  float *h_state[3];     // [maxB, D] final states: main(s), main(s'), boot(s')
  float *gates_save;     // [maxB, L, dirs, 4, H]  r,z,n,pre_h_n of main(s)
  float *hprev_save;     // [maxB, L, dirs, H]
  float *dgi, *dgh;      // [maxB, L, dirs, 3H]
  float *dx;             // [maxB, L, dirs, E]
  float *dh;             // [maxB, D]
  float *dh_part;        // [n_part, maxB, D]
  int n_dh_part;
  float *wgrad_part;     // [splits][dirs][3H][E+H+2]
  int wgrad_splits;
  int wgrad_used;        // split-K slices the most recent weight-gradient launch wrote (<= wgrad_splits)
  int32_t *emb_keys;     // [maxB*L] row id or -1
  int32_t *emb_slot;     // [N+1] slot of the leader position or -1
  float *emb_grad_rows;  // [maxB*L, E]
  uint8_t *emb_leader;   // [maxB*L] chunk-leader flags (batches spanning several dedup chunks)
  int32_t *emb_sorted;   // [maxB*L] positions sorted by (row, position)  (sort-based dedup, E = 64)
  int32_t *emb_seg;      // [maxB*L + 2] row of each sorted entry; [cap] = valid entries, [cap+1] = finished-tile counter
  int32_t *emb_csort;    // [2][emb_csort_n] per-chunk sorted (position | row) lists of batches beyond one chunk
  int emb_csort_n;
  int32_t *emb_ccount;   // [16] valid entries per chunk
  float *emb_carry;      // [maxB*L/32 + 1, 64] partial sums of runs continued from the previous tile
  int32_t *emb_tmeta;    // [maxB*L/32 + 1, 2] (tile starts inside a run, leader of the tile's last run)
  // head statistics partials: [n_split][maxB][PART_STRIDE]
  float *part;
  int part_stride, n_split_max;
  float *row_stats;      // [maxB][ROW_STRIDE] merged per-row results
  int32_t *row_ids;      // [maxB][REC_MAX_TOPK] merged top-k ids / argmax
  float *row_topv;       // [maxB][REC_MAX_TOPK]
  float *q_sa;           // [maxB][3]
  float *q_boot;         // [maxB][3]
  float *dq;             // [maxB][3]
  float *rewards;        // [maxB][3]
  float *loss_buf;       // [8]
  int32_t *astar;        // [maxB]
  uint8_t *drop_mask;    // [maxB, D] keep mask of the last supervised step with dropout
  // saved call context for the phase-split API
  rec_batch cur_batch;
  rec_batch dp_local;    // data-parallel trunk: this rank's own sessions (engine-owned copy)
  bool dp_active;
  rec_train_hparams cur_hp;
  int cur_main, cur_topk, cur_phase;
  float cur_step_size, cur_bc2_sqrt;
  float *q_grad_rows;    // [maxB][n_q][D] row-sparse Q-head gradients
  float *q_bgrad;        // [maxB][n_q]
  int32_t *q_slot;       // [Vloc] leader batch row of each action or -1
  uint8_t *hpack;        // [ceil(maxB/128)][hi|lo][16 KB] packed bf16 image of h for the tensor-core backward
  bool hpack_ready;      // the packed image of the current supervised states was already produced (early, off the critical path)
  float *summary;        // [maxB][part_stride] per-row record of this shard
  float *qpack;          // [2][maxB][3] Q(s,a) | Q_boot(s',a*) contributions of this shard
  bool timing;
  float *h_sc, *d_sc;    // d_sc: Adam scalars {lr/(1-b1^t), 1/sqrt(1-b2^t)} written by adam_step_kernel, read by every Adam kernel
  long long *d_step;     // [REC_MAX_NETS] device-side Adam step counters
  // CUDA-graph replay of the single-GPU train step (fixed engine-owned input buffers)
  bool use_graph;
  cudaStream_t cap_stream;  // private stream used only while capturing (the caller's stream may be the legacy default)
  // Independent branches of the step (Q-head Adam sweep next to the supervised-head kernel, embedding chain
  // next to the GRU weight update) run on a second stream; under graph capture they become parallel branches.
  cudaStream_t side[3];
  cudaEvent_t ev_fork[3], ev_join[3], ev_mark[4];
  bool overlap;        // REC_NO_OVERLAP=1 serialises everything on the caller's stream
  bool tl_on;          // REC_TIMELINE=1
  int tl_n, tl_steps;
  struct TlEntry { cudaEvent_t ev; const char *file; int line; int stream; } tl[400];
  bool side_dirty[3];  // work was issued on side[i] since its last join
  rec_batch own;         // engine-owned copy of the caller's batch (pointers into own_block)
  uint8_t *own_block, *h_own;  // device block and its pinned host mirror (host entry points)
  size_t own_bytes;
  float *h_loss;         // pinned: losses of the last host-entry step
  struct GraphEntry { uint64_t key; void *exec; int launches; int seen; } graphs[16];
  int n_graphs;
  long long *trace;      // optional device buffer for clock64 phase traces (debug)
  bool use_tc;           // tensor-core (tcgen05) head kernels when D == 64
  // what the most recent head-statistics launch published per (split,row) record (consumed by the merge that follows)
  int st_kpub;           // top-k candidates per record (>= topk when the scores are approximate: see rec_kpub)
  int st_apub;           // greedy-action candidates per record in the top-k slots (0: only the (value,id) pair at [3],[4])
  bool st_approx;        // scores are bf16x3 tensor-core products: the merge re-scores its candidates in fp32
  cudaEvent_t ev[12];
  float last_ms[3];
  // operand images of the K-loop tensor-core head kernels for D >= 128 (heads_tck.cu); allocated at first use
  uint8_t *k_wimg[2];    // [0] statistics / backward head, [1] greedy-action heads (pre-combined)
  uint8_t *k_himg[2];    // states scored by [0] / [1]
  uint8_t *k_hT;         // h^T image (dW GEMM)
  uint8_t *k_dlT;        // dlogits^T image [V, B]
  float *k_db;           // bias-gradient partials [session blocks][V]
  float *k_bias;         // combined bias of the greedy-action heads
  int k_sup_net, k_sup_head;  // which (net, head) k_wimg[0] / k_himg[0] currently hold (-1: none)
  float *k_cmax;         // chunk maxima [B][V/32] of the evaluation top-k (HeadTopk<.., CM>)
  bool k_hold;           // rec_eval_hold_params: parameters are frozen by the caller, images may be reused
  int64_t param_epoch, k_img_epoch;  // bumped by every entry point that may write parameters / epoch of the cached image
  // rec_set_embedding_shard: rows [emb_row_lo, emb_row_hi) of every net's embedding table are OWNED by this rank (the
  // Adam sweep touches no other row); emb_row_hi == 0: all rows (unsharded, the default)
  int64_t emb_row_lo, emb_row_hi;
  int k_img_net, k_img_head;
  uint8_t *k_bblk;       // bias operand blocks of HeadCmaxPair [tiles + 1][4 KB]
  float *k_cmax2;        // level-2 maxima [B][2 * ceil(tiles / 8)] (HeadCmaxPair)
  int *k_chosen;         // [B][KC] the chunks with the largest maxima per row
  // SARM (5 Q heads, rec_train_step_sarm)
  float *sarm_qmax;      // [5][maxB] max_a Q_i(s', a), exact (re-scored)
  float *sarm_dq;        // [maxB][5] dL/dQ_i(s_b, a_b)
  float *sarm_extra;     // [maxB] head 0's Q gradient, added to its dlogits at the target column
  const float *bwd_extra;  // non-null while a supervised-head backward must add a per-row gradient at the target column
  bool k_fresh[2];       // k_wimg[i] was packed ahead of its consumer in this step (tck_prepack_heads)
  // tensor-core GRU trunk for E, H >= 128 (gru_tc.cu); allocated at first use
  uint8_t *g_wimg;       // [net][dir][W_ih | W_hh | W_hh regrouped] weight images
  uint8_t *g_ximg[2], *g_ximg2;  // gathered embedding rows of s / s' (main table) / s' (bootstrap table)
  float *g_gi[3];        // [B L, dirs, 3H] input projections per pass
  uint8_t *g_himg[2];    // ping/pong bf16 images of h_t: [pass][dir][session block][H/64]
  uint8_t *g_hprev_img;  // [dir][B L][H] image of h_{t-1} per position (dW_hh GEMM)
  uint8_t *g_dstep[2];   // ping/pong images of the gate gradients of one step: [dir][session block][3H/64]
  uint8_t *g_dgi_img, *g_dgh_img;  // [dir][B L][3H] images of the gate gradients (dx / weight-gradient GEMMs)
  float *g_dhw;          // [B, D] running dL/dh of the BPTT
};

#define REC_FAIL(e, code, ...)                          \
  do {                                                  \
    if (e) snprintf((e)->err, sizeof((e)->err), __VA_ARGS__); \
    return (code);                                      \
  } while (0)

#define REC_CUDA(e, call)                                                                 \
  do {                                                                                    \
    cudaError_t _st = (call);                                                             \
    if (_st != cudaSuccess) {                                                             \
      if (e) snprintf((e)->err, sizeof((e)->err), "%s failed: %s (%s:%d)", #call,         \
                      cudaGetErrorString(_st), __FILE__, __LINE__);                       \
      return REC_ECUDA;                                                                   \
    }                                                                                     \
  } while (0)

// REC_TIMELINE=1 (debug, eager mode): an event after every launch; rec_timeline_dump prints when each kernel
// finished relative to the start of the step and on which stream.
void rec_timeline_record(rec_engine *e, const char *file, int line);
#define REC_LAUNCH_CHECK(e)                                                               \
  do {                                                                                    \
    (e)->launches++;                                                                      \
    if ((e)->tl_on) rec_timeline_record((e), __FILE__, __LINE__);                         \
    cudaError_t _st = cudaGetLastError();                                                 \
    if (_st != cudaSuccess) {                                                             \
      snprintf((e)->err, sizeof((e)->err), "kernel launch failed: %s (%s:%d)",            \
               cudaGetErrorString(_st), __FILE__, __LINE__);                              \
      return REC_ECUDA;                                                                   \
    }                                                                                     \
  } while (0)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// torch.optim.Adam element update (lerp / addcmul / addcdiv form).  sqrt and the final division use the
// hardware approximations (<= 2 ulp): the update is HBM-bound streaming work and the IEEE sequences cost
// ~10x the instructions; the difference is far inside the 1e-3 parity tolerance.
__device__ __forceinline__ float fast_sqrtf(float x) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void adam_elem(float &p, float &m, float &v, float g, float b1, float b2, float eps,
                                          float step_size, float inv_bc2_sqrt) {
  m = fmaf(g - m, 1.f - b1, m);
  v = fmaf((1.f - b2) * g, g, v * b2);
  p = p - __fdividef(step_size * m, fmaf(fast_sqrtf(v), inv_bc2_sqrt, eps));
}

// (score desc, id asc): is candidate (v, i) strictly better than (w, j)?
__device__ __forceinline__ bool better(float v, int i, float w, int j) {
  return (v > w) || (v == w && i < j);
}

// Candidate margins of the fp32 re-score (SURVEY 7 hard part 1): the tensor-core statistics kernels rank by bf16x3
// scores (~1e-5 relative), so every stage keeps MORE candidates than the k it is asked for, and the merge that
// ends the chain re-scores its top-(k + margin) candidates with a fixed-order fp32 FFMA dot product and orders them
// by (exact score desc, id asc) -- torch.topk / argmax on fp32 logits (eval_protocol.py:75, sqn_gru.py:229,
// tensor_operations.py:73-84).  `cs` = threads that share a row inside the CTA (each keeps a private top-k list).
__host__ __device__ inline int rec_kpub(int topk, int cs) {
  int k = topk + (topk <= 2 ? 2 : 8);
  if (k > cs * topk) k = cs * topk;
  return k > REC_MAX_TOPK ? REC_MAX_TOPK : k;
}
__host__ __device__ inline int rec_kt(int topk) {  // candidates the merge re-scores
  const int k = topk + (topk <= 2 ? 3 : 8);
  return k > REC_MAX_TOPK ? REC_MAX_TOPK : k;
}
#define REC_ARG_CAND 4   // greedy-action candidates the merge re-scores

// Every ABI entry point runs with the engine's device current and restores the caller's device afterwards
// (an engine on cuda:1 called while cuda:0 is current would otherwise launch on a stream of another device).
struct DevGuard {
  int prev;
  explicit DevGuard(const rec_engine *e) : prev(-1) {
    if (!e) return;
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != e->dev) { prev = cur; cudaSetDevice(e->dev); }
  }
  ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- launchers implemented in the .cu files -------------------------------------------------
// gru.cu
int launch_gru_forward(rec_engine *e, int net_id, const int64_t *s, const int64_t *lengths, int B,
                       float *h_out, bool save);
int launch_gru_forward_multi(rec_engine *e, int n_pass, const int *net_ids, const int64_t *const *s,
                             const int64_t *const *lengths, float *const *h_out, const bool *save, int B);
// stages: 1 = BPTT (dgi, dgh, dx), 2 = weight gradients, 4 = Adam on the GRU parameters
int launch_gru_backward(rec_engine *e, int net_id, const int64_t *s, const int64_t *lengths, int B,
                        const float *dh, float step_size, float bc2_sqrt, const rec_train_hparams *hp, int stages = 7);
int launch_gru_transpose(rec_engine *e, int net_id);
// embed.cu
// stages: 1 = order the token positions (needs only the batch), 2 = combine duplicate rows (needs dx),
// 4 = dense Adam sweep + slot reset (must run after the GRU weight gradients, which read the table)
int launch_embedding_update(rec_engine *e, int net_id, const int64_t *s, const int64_t *lengths, int B,
                            float step_size, float bc2_sqrt, const rec_train_hparams *hp, int stages = 7);

// wait_mark >= 0: the streaming sweep (not the gradient-row kernel) additionally waits for side_mark(e, wait_mark)
int launch_q_heads_adam(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B, float step_size,
                        float bc2_sqrt, const rec_train_hparams *hp, int wait_mark = -1);
// generalised: heads first_head .. first_head + n - 1 with dq[b * dq_stride + j] (SARM: heads 1..4, stride 5)
int launch_q_heads_adam_ex(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B, float step_size,
                           float bc2_sqrt, const rec_train_hparams *hp, int first_head, int n, const float *dq,
                           int dq_stride, int wait_mark = -1);

// Launches issued while a SideScope is alive go to side stream `idx`, ordered after everything issued so far on
// the current stream -- or, with `mark` >= 0, after the point remembered by side_mark(e, mark).  Scopes nest
// (the current stream may itself be a side stream).  side_join() makes the current stream wait for side `idx`.
// With overlap off the scope is a no-op and everything runs in program order on one stream.
static inline bool side_enabled(const rec_engine *e) { return e->overlap && !e->timing && !e->trace; }
static inline void side_mark(rec_engine *e, int k) {
  if (side_enabled(e)) cudaEventRecord(e->ev_mark[k], e->stream);
}
struct SideScope {
  rec_engine *e;
  cudaStream_t main;
  SideScope(rec_engine *e_, int idx, int mark = -1) : e(e_), main(e_->stream) {
    if (side_enabled(e)) {
      cudaEvent_t ev = mark >= 0 ? e->ev_mark[mark] : e->ev_fork[idx];
      if (mark < 0) cudaEventRecord(ev, main);
      cudaStreamWaitEvent(e->side[idx], ev, 0);
      e->stream = e->side[idx];
      e->side_dirty[idx] = true;
    }
  }
  ~SideScope() { e->stream = main; }
};
static inline void side_wait_mark(rec_engine *e, int k) {
  if (side_enabled(e)) cudaStreamWaitEvent(e->stream, e->ev_mark[k], 0);
}
static inline void side_join(rec_engine *e, int idx) {
  if (!e->side_dirty[idx]) return;
  cudaEventRecord(e->ev_join[idx], e->side[idx]);
  cudaStreamWaitEvent(e->stream, e->ev_join[idx], 0);
  e->side_dirty[idx] = false;
}
// heads.cu
struct HeadStatsArgs {
  int net_id;
  const float *h;         // [B, D]
  int B;
  int do_stats;           // 1: online (max,sumexp) + target logit of head `stats_head`
  int stats_head;         // head whose logits feed stats / top-k
  const int64_t *target;  // [B] global action ids (target logit)
  int topk;               // >0: running top-k of `stats_head` (score desc, id asc)
  int n_arg;              // >0: argmax over sum_j w[j] * Q_{1+arg_shift+j}, j < n_arg
  float w[3];
  int arg_shift;          // 0: the Q heads follow the supervised head (SQN / SMORL); SARM scores head i with arg_shift = i - 1
};
int launch_head_stats(rec_engine *e, const HeadStatsArgs &a, int *n_split_out);
bool tc_heads_supported(const rec_engine *e);
int launch_head_stats_tc(rec_engine *e, const HeadStatsArgs &a, int *n_split_out);
bool tc_bwd_supported(const rec_engine *e, int B);
// heads_tck.cu: D = 128, 256, ... (K-loop pipelines over packed operand images)
bool tck_heads_supported(const rec_engine *e);
bool tck_topk_supported(const rec_engine *e, const HeadStatsArgs &a);
bool tck_chunk_topk_supported(const rec_engine *e, const HeadStatsArgs &a);
int launch_head_topk_chunks(rec_engine *e, const HeadStatsArgs &a, int *n_split_out, float *summary);
int launch_head_stats_tck(rec_engine *e, const HeadStatsArgs &a, int *n_split_out);
int tck_bwd_slices(const rec_engine *e);
int launch_head_bwd_adam_tck(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B, float step_size,
                             float bc2_sqrt, const rec_train_hparams *hp, float inv_B);
void tck_free(rec_engine *e);
int tck_prepack_heads(rec_engine *e, int net_id, int n_arg, const float *w);
// gru_tc.cu: GRU trunk for E, H multiples of 128
bool gru_tc_supported(const rec_engine *e);
int launch_gru_forward_tc(rec_engine *e, int n_pass, const int *net_ids, const int64_t *const *s,
                          const int64_t *const *lengths, float *const *h_out, const bool *save, int B);
int launch_gru_backward_tc(rec_engine *e, int net_id, const int64_t *s, const int64_t *lengths, int B, const float *dh,
                           int stages);
void gtc_free(rec_engine *e);
int launch_head_bwd_adam_tc(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B, float step_size,
                            float bc2_sqrt, const rec_train_hparams *hp, float inv_B, int *n_slices);
int launch_h_prepack_early(rec_engine *e, const float *h, int B);
// `src`: the statistics launch that produced `part` (its records may hold approximate scores and extra candidates:
// e->st_*); nullptr: `part` holds exact per-shard summaries written by a previous merge (cross-GPU merge).
int launch_head_merge(rec_engine *e, const float *part, int n_split, int B, int topk, bool has_stats,
                      bool has_argmax, float *summary = nullptr, const HeadStatsArgs *src = nullptr);
int launch_head_logits(rec_engine *e, int net_id, int head, const float *h, int B, float *logits, int64_t ld);
int launch_row_dots(rec_engine *e, int net_id, const float *h, const int64_t *ids, const int32_t *ids32,
                    int B, int first_head, int n, float *out);
int launch_head_backward_adam(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B,
                              float step_size, float bc2_sqrt, const rec_train_hparams *hp, float inv_B);
int head_bwd_dense_slices(const rec_engine *e, int B);
int tc_bwd_slices(const rec_engine *e);
int launch_q_dh(rec_engine *e, int net_id, const rec_batch *b, int B);
int launch_sup_head_bwd(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B, float step_size,
                        float bc2_sqrt, const rec_train_hparams *hp, float inv_B);
int launch_q_heads_update(rec_engine *e, int net_id, const float *h, const rec_batch *b, int B, float step_size,
                          float bc2_sqrt, const rec_train_hparams *hp, int wait_mark = -1);
int launch_dh_reduce(rec_engine *e, int B);
int launch_emb_rows(rec_engine *e, int net_id, const int64_t *ids, int64_t n, float *rows, bool scatter);
int launch_dropout(rec_engine *e, int net_id, float *h, float *dh, int B, const rec_train_hparams *hp, bool backward,
                   int mask_row0 = 0);
int launch_q_rows_fused(rec_engine *e, int main_net, const rec_batch *b, const rec_train_hparams *hp, int n_split,
                        float alpha_eff, float *q_loss_rows, const HeadStatsArgs *src);
