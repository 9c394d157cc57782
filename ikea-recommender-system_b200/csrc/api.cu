// C ABI (include/recsys_b200.h): engine lifetime, parameter binding, train steps, evaluation.
#include <math.h>
#include <stdlib.h>
#include <new>
#include "common.cuh"

// launchers from the other translation units
int launch_fill_i32(rec_engine *e, int32_t *p, int64_t n, int32_t v);
int launch_td(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp, int n_q, float alpha_eff,
              float *q_loss_rows);
int launch_loss_reduce(rec_engine *e, int B, const float *q_loss_rows, float *out);
int launch_eval_metrics(rec_engine *e, const rec_batch *b, const rec_eval_opts *o, int kmax,
                        const rec_eval_accum *acc, double *rowm, int32_t *topk_ids, float *topk_scores);
size_t head_bwd_smem_bytes(int D);
int launch_pack_batch(rec_engine *e, const rec_batch *b, uint8_t *out);
int launch_gather_batch(rec_engine *e, const rec_batch *columns, int64_t n_rows, const int64_t *idx, int B, const rec_batch *out);
int launch_build_rows(rec_engine *e, const int64_t *off, int64_t n_sessions, const int64_t *items, const float *rewards,
                      int64_t n, int64_t pad_id, int pad_end, const rec_batch *out);
int launch_unpack_batch(rec_engine *e, const uint8_t *gathered, int G, int Bl, size_t stride, const rec_batch *out);

static char g_err[512] = "";

static int head_stats_dispatch(rec_engine *e, const HeadStatsArgs &a, int *n_split) {
  if (tck_topk_supported(e, a)) return launch_head_stats_tck(e, a, n_split);  // evaluation-shaped top-k (any D % 64 == 0)
  if (tc_heads_supported(e)) return launch_head_stats_tc(e, a, n_split);
  if (tck_heads_supported(e) && a.topk == 0) return launch_head_stats_tck(e, a, n_split);
  return launch_head_stats(e, a, n_split);
}

struct EngineExtra {
  float *q_loss_rows;
  double *rowm;
};
static EngineExtra &extra(rec_engine *e) { return *reinterpret_cast<EngineExtra *>(e + 1); }

#define ALLOC(e, ptr, type, count)                                                      \
  do {                                                                                  \
    cudaError_t _st = cudaMalloc((void **)&(ptr), sizeof(type) * (size_t)(count));      \
    if (_st != cudaSuccess) {                                                           \
      snprintf(g_err, sizeof(g_err), "cudaMalloc(%s, %zu B) failed: %s", #ptr,          \
               sizeof(type) * (size_t)(count), cudaGetErrorString(_st));                \
      rec_destroy(e);                                                                   \
      return REC_ENOMEM;                                                                \
    }                                                                                   \
  } while (0)

extern "C" int rec_abi_version(void) { return REC_ABI_VERSION; }

extern "C" const char *rec_last_error(const rec_engine *e) {
  DevGuard dev_guard(e); return e ? e->err : g_err; }

extern "C" int rec_create(const rec_config *cfg, void *stream, rec_engine **out) {
  if (!cfg || !out) { snprintf(g_err, sizeof(g_err), "rec_create: null argument"); return REC_EINVAL; }
  *out = nullptr;
  int dev = 0;
  cudaError_t st = cudaGetDevice(&dev);
  cudaDeviceProp prop;
  if (st == cudaSuccess) st = cudaGetDeviceProperties(&prop, dev);
  if (st != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "rec_create: no CUDA device (%s); this library has no CPU fallback",
             cudaGetErrorString(st));
    return REC_ENODEV;
  }
  if (prop.major != 10) {
    snprintf(g_err, sizeof(g_err), "rec_create: device %s is sm_%d%d; this library is built for sm_100a (B200) only",
             prop.name, prop.major, prop.minor);
    return REC_ENODEV;
  }
  const rec_config &c = *cfg;
  if (c.embedding_dim % 4 || c.hidden_dim % 4 || c.embedding_dim <= 0 || c.hidden_dim <= 0) {
    snprintf(g_err, sizeof(g_err), "rec_create: embedding_dim (%d) and hidden_dim (%d) must be positive multiples of 4",
             c.embedding_dim, c.hidden_dim);
    return REC_EINVAL;
  }
  if (dev >= REC_MAX_DEVICES) { snprintf(g_err, sizeof(g_err), "rec_create: device ordinal %d not supported", dev); return REC_EINVAL; }
  if (c.n_heads < 1 || c.n_heads > REC_MAX_HEADS || c.n_heads == 3 || c.n_nets < 1 || c.n_nets > REC_MAX_NETS ||
      c.max_batch < 1 || c.state_size < 1 || c.item_num < 1 || c.action_dim < 1 || c.vocab_lo < 0 ||
      c.vocab_hi > c.action_dim || c.vocab_lo >= c.vocab_hi || c.max_topk < 1 || c.max_topk > REC_MAX_TOPK) {
    snprintf(g_err, sizeof(g_err), "rec_create: invalid configuration");
    return REC_EINVAL;
  }
  void *mem = calloc(1, sizeof(rec_engine) + sizeof(EngineExtra));
  if (!mem) return REC_ENOMEM;
  rec_engine *e = new (mem) rec_engine;
  memset((void *)e, 0, sizeof(rec_engine) + sizeof(EngineExtra));
  e->cfg = c;
  e->dev = dev;
  e->stream = (cudaStream_t)stream;
  e->dirs = c.bidirectional ? 2 : 1;
  e->D = c.hidden_dim * e->dirs;
  e->Vloc = c.vocab_hi - c.vocab_lo;
  e->sm_count = prop.multiProcessorCount;
  e->use_tc = true;
  e->k_sup_net = -1; e->k_sup_head = -1;
  if (head_bwd_smem_bytes(e->D) > 220 * 1024) {
    snprintf(g_err, sizeof(g_err), "rec_create: head width D=%d exceeds the shared-memory budget of the backward kernel", e->D);
    free(mem);
    return REC_EINVAL;
  }
  const int64_t mb = c.max_batch, L = c.state_size, H = c.hidden_dim, E = c.embedding_dim, G = 3 * H;
  const int dirs = e->dirs, D = e->D;
  for (int i = 0; i < 3; ++i) ALLOC(e, e->h_state[i], float, mb * D);
  ALLOC(e, e->gates_save, float, mb * L * dirs * 4 * H);
  ALLOC(e, e->hprev_save, float, mb * L * dirs * H);
  ALLOC(e, e->dgi, float, mb * L * dirs * G);
  ALLOC(e, e->dgh, float, mb * L * dirs * G);
  ALLOC(e, e->dx, float, mb * L * dirs * E);
  ALLOC(e, e->dh, float, mb * D);
  const int n_tiles = (e->Vloc + 63) / 64;
  e->n_dh_part = (n_tiles < 2 * e->sm_count ? n_tiles : 2 * e->sm_count) + 1;
  ALLOC(e, e->dh_part, float, (int64_t)e->n_dh_part * mb * D);
  e->wgrad_splits = 64;
  e->wgrad_used = 64;
  const int64_t KS = (E > H ? E : H) + 1;
  ALLOC(e, e->wgrad_part, float, (int64_t)e->wgrad_splits * dirs * 2 * G * KS);
  ALLOC(e, e->emb_keys, int32_t, mb * L);
  ALLOC(e, e->emb_slot, int32_t, (int64_t)c.item_num + 1);
  ALLOC(e, e->emb_grad_rows, float, mb * L * E);
  ALLOC(e, e->emb_leader, uint8_t, mb * L);
  ALLOC(e, e->emb_sorted, int32_t, mb * L);
  ALLOC(e, e->emb_seg, int32_t, mb * L + 2);
  cudaMemsetAsync(e->emb_seg, 0, sizeof(int32_t) * (mb * L + 2), e->stream);
  e->emb_csort_n = (int)((mb * L + 4095) / 4096 * 4096);
  ALLOC(e, e->emb_csort, int32_t, 2 * (int64_t)e->emb_csort_n);
  ALLOC(e, e->emb_ccount, int32_t, 16);
  ALLOC(e, e->emb_carry, float, (mb * L / 32 + 1) * (int64_t)(E > 64 ? E : 64));
  ALLOC(e, e->emb_tmeta, int32_t, (mb * L / 32 + 1) * 2);
  e->part_stride = 72;
  e->n_split_max = 2 * e->sm_count;
  ALLOC(e, e->part, float, ((int64_t)e->sm_count * 4 * 128 + 4 * mb) * e->part_stride);
  ALLOC(e, e->row_stats, float, mb * 8);
  ALLOC(e, e->row_ids, int32_t, mb * REC_MAX_TOPK);
  ALLOC(e, e->row_topv, float, mb * REC_MAX_TOPK);
  ALLOC(e, e->q_sa, float, mb * 3);
  ALLOC(e, e->q_boot, float, mb * 3);
  ALLOC(e, e->dq, float, mb * 3);
  ALLOC(e, e->rewards, float, mb * 3);
  ALLOC(e, e->loss_buf, float, 8);
  ALLOC(e, e->astar, int32_t, mb);
  ALLOC(e, e->drop_mask, uint8_t, mb * D);
  ALLOC(e, extra(e).q_loss_rows, float, mb);
  ALLOC(e, e->summary, float, mb * e->part_stride);
  ALLOC(e, e->d_sc, float, 4);
  ALLOC(e, e->d_step, long long, REC_MAX_NETS);
  cudaMemsetAsync(e->d_step, 0, sizeof(long long) * REC_MAX_NETS, e->stream);
  if (cudaMallocHost((void **)&e->h_sc, 4 * sizeof(float)) != cudaSuccess) { snprintf(g_err, sizeof(g_err), "cudaMallocHost failed"); rec_destroy(e); return REC_ENOMEM; }
  {
    // engine-owned batch: ONE block [int64 s | s_next | a | true_len | true_next_len][float r][uint8 is_end] so
    // that the host entry points fill it with a single H2D copy from the pinned mirror h_own
    e->own_bytes = (size_t)mb * (2 * L + 3) * 8 + (size_t)mb * 4 + (size_t)mb;
    ALLOC(e, e->own_block, uint8_t, (int64_t)e->own_bytes);
    int64_t *p64 = (int64_t *)e->own_block;
    e->own.s = p64; e->own.s_next = p64 + mb * L; e->own.a = p64 + 2 * mb * L; e->own.true_len = p64 + 2 * mb * L + mb;
    e->own.true_next_len = p64 + 2 * mb * L + 2 * mb;
    e->own.r = (float *)(p64 + mb * (2 * L + 3)); e->own.is_end = (uint8_t *)(e->own.r + mb);
    if (cudaMallocHost((void **)&e->h_own, e->own_bytes) != cudaSuccess ||
        cudaMallocHost((void **)&e->h_loss, 4 * sizeof(float)) != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "cudaMallocHost failed"); rec_destroy(e); return REC_ENOMEM;
    }
  }
  e->use_graph = getenv("REC_NO_GRAPH") == nullptr;
  e->overlap = getenv("REC_NO_OVERLAP") == nullptr;
  e->tl_on = getenv("REC_TIMELINE") != nullptr;
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  {
    // priorities (honoured between graph branches): main/capture stream > side 0/1 (supervised-head backward,
    // embedding chain) > side 2 (HBM-bound Q-head sweep, which must not push critical-path CTAs off the SMs)
    const int prio_mid = (prio_least + prio_greatest) / 2;
    const int prio[3] = {prio_mid, prio_mid, prio_least};
    for (int i = 0; i < 3; ++i) {
      e->side_dirty[i] = false;
      if (cudaStreamCreateWithPriority(&e->side[i], cudaStreamNonBlocking, prio[i]) != cudaSuccess ||
          cudaEventCreateWithFlags(&e->ev_fork[i], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming) != cudaSuccess)
        e->overlap = false;
    }
    for (int i = 0; i < 4; ++i)
      if (cudaEventCreateWithFlags(&e->ev_mark[i], cudaEventDisableTiming) != cudaSuccess) e->overlap = false;
  }
  // the capture stream (main branch of the graph) outranks the side branch: the one-CTA-per-SM tensor-core
  // kernels must get their SM slots before the streaming kernels fill the machine
  if (cudaStreamCreateWithPriority(&e->cap_stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) e->use_graph = false;
  ALLOC(e, e->hpack, uint8_t, (size_t)((mb + 255) / 256) * 4 * 16384);
  ALLOC(e, e->q_grad_rows, float, mb * 4 * D);  // up to 4 row-sparse heads (SARM: heads 1..4)
  ALLOC(e, e->q_bgrad, float, mb * 4);
  if (c.n_heads == 5) {
    ALLOC(e, e->sarm_qmax, float, 5 * mb);
    ALLOC(e, e->sarm_dq, float, mb * 5);
    ALLOC(e, e->sarm_extra, float, mb);
  }
  ALLOC(e, e->q_slot, int32_t, e->Vloc);
  ALLOC(e, e->qpack, float, 2 * mb * 3);
  ALLOC(e, extra(e).rowm, double, mb * (3 * REC_MAX_KLIST + 3));
  for (int n = 0; n < c.n_nets; ++n)
    for (int d = 0; d < dirs; ++d) {
      ALLOC(e, e->nets[n].w_ihT[d], float, G * E);
      ALLOC(e, e->nets[n].w_hhT[d], float, G * H);
    }
  for (int i = 0; i < 12; ++i) cudaEventCreate(&e->ev[i]);
  if (launch_fill_i32(e, e->emb_slot, (int64_t)c.item_num + 1, -1) != REC_OK ||
      launch_fill_i32(e, e->q_slot, e->Vloc, -1) != REC_OK) {
    snprintf(g_err, sizeof(g_err), "%s", e->err);
    rec_destroy(e);
    return REC_ECUDA;
  }
  cudaMemsetAsync(e->q_sa, 0, sizeof(float) * mb * 3, e->stream);
  cudaMemsetAsync(e->q_boot, 0, sizeof(float) * mb * 3, e->stream);
  cudaMemsetAsync(e->dq, 0, sizeof(float) * mb * 3, e->stream);
  e->err[0] = 0;
  *out = e;
  return REC_OK;
}

extern "C" void rec_destroy(rec_engine *e) {
  DevGuard dev_guard(e);
  if (!e) return;
  cudaStreamSynchronize(e->stream);
  void *ptrs[] = {e->h_state[0], e->h_state[1], e->h_state[2], e->gates_save, e->hprev_save, e->dgi, e->dgh, e->dx,
                  e->dh, e->dh_part, e->wgrad_part, e->emb_keys, e->emb_slot, e->emb_grad_rows, e->part, e->row_stats,
                  e->row_ids, e->row_topv, e->q_sa, e->q_boot, e->dq, e->rewards, e->loss_buf, e->astar, e->drop_mask,
                  extra(e).q_loss_rows, extra(e).rowm, e->summary, e->qpack, e->q_grad_rows, e->q_bgrad, e->q_slot, e->hpack, e->emb_leader, e->emb_sorted, e->emb_seg, e->emb_csort, e->emb_ccount, e->emb_carry, e->emb_tmeta, e->d_sc, e->d_step, (void *)e->own_block, e->sarm_qmax, e->sarm_dq, e->sarm_extra};
  for (void *p : ptrs) if (p) cudaFree(p);
  tck_free(e);
  gtc_free(e);
  for (int n = 0; n < REC_MAX_NETS; ++n)
    for (int d = 0; d < 2; ++d) {
      if (e->nets[n].w_ihT[d]) cudaFree(e->nets[n].w_ihT[d]);
      if (e->nets[n].w_hhT[d]) cudaFree(e->nets[n].w_hhT[d]);
    }
  for (int i = 0; i < 12; ++i) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
  for (int i = 0; i < e->n_graphs; ++i) if (e->graphs[i].exec) cudaGraphExecDestroy((cudaGraphExec_t)e->graphs[i].exec);
  if (e->h_sc) cudaFreeHost(e->h_sc);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  if (e->h_own) cudaFreeHost(e->h_own);
  if (e->h_loss) cudaFreeHost(e->h_loss);
  for (int i = 0; i < 3; ++i) {
    if (e->side[i]) { cudaStreamSynchronize(e->side[i]); cudaStreamDestroy(e->side[i]); }
    if (e->ev_fork[i]) cudaEventDestroy(e->ev_fork[i]);
    if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
  }
  for (int i = 0; i < 4; ++i) if (e->ev_mark[i]) cudaEventDestroy(e->ev_mark[i]);
  free(e);
}

static int check_net(rec_engine *e, int net_id, bool need_opt) {
  if (!e) return REC_EINVAL;
  if (net_id < 0 || net_id >= e->cfg.n_nets) REC_FAIL(e, REC_EINVAL, "net_id %d out of range", net_id);
  if (!e->nets[net_id].bound) REC_FAIL(e, REC_EINVAL, "net %d has no bound parameters (call rec_bind_params)", net_id);
  if (need_opt && !e->nets[net_id].p.emb_m) REC_FAIL(e, REC_EINVAL, "net %d was bound without Adam state", net_id);
  return REC_OK;
}

extern "C" int rec_bind_params(rec_engine *e, int net_id, const rec_net_params *p) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  if (!e || !p) return REC_EINVAL;
  if (net_id < 0 || net_id >= e->cfg.n_nets) REC_FAIL(e, REC_EINVAL, "net_id %d out of range", net_id);
  if (!p->emb) REC_FAIL(e, REC_EINVAL, "rec_bind_params: emb is null");
  for (int d = 0; d < e->dirs; ++d)
    if (!p->w_ih[d] || !p->w_hh[d] || !p->b_ih[d] || !p->b_hh[d]) REC_FAIL(e, REC_EINVAL, "rec_bind_params: GRU pointer null (dir %d)", d);
  for (int h = 0; h < e->cfg.n_heads; ++h)
    if (!p->head_w[h] || !p->head_b[h]) REC_FAIL(e, REC_EINVAL, "rec_bind_params: head %d pointer null", h);
  NetBind &nb = e->nets[net_id];
  nb.p = *p;
  nb.bound = true;
  return launch_gru_transpose(e, net_id);
}

__global__ void set_step_kernel(long long *d_step, long long v) { *d_step = v; }

extern "C" int rec_set_adam_step(rec_engine *e, int net_id, int64_t step) {
  DevGuard dev_guard(e);
  if (!e || net_id < 0 || net_id >= e->cfg.n_nets) return REC_EINVAL;
  e->nets[net_id].adam_step = step;
  set_step_kernel<<<1, 1, 0, e->stream>>>(e->d_step + net_id, (long long)step);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}
extern "C" int64_t rec_get_adam_step(const rec_engine *e, int net_id) {
  DevGuard dev_guard(e);
  if (!e || net_id < 0 || net_id >= e->cfg.n_nets) return -1;
  // the device counter is authoritative (steps replayed from a caller-captured graph never pass through the host)
  long long t = 0;
  if (cudaStreamSynchronize(e->stream) != cudaSuccess ||
      cudaMemcpy(&t, e->d_step + net_id, sizeof(t), cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  const_cast<rec_engine *>(e)->nets[net_id].adam_step = t;
  return t;
}

extern "C" int rec_set_stream(rec_engine *e, void *stream) {
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  e->stream = (cudaStream_t)stream;
  return REC_OK;
}

// ---- row-sharded embedding table -----------------------------------------------------------------------------------
static void drop_step_graphs(rec_engine *e) {
  for (int i = 0; i < e->n_graphs; ++i)
    if (e->graphs[i].exec) cudaGraphExecDestroy((cudaGraphExec_t)e->graphs[i].exec);
  e->n_graphs = 0;
}

extern "C" int rec_set_embedding_shard(rec_engine *e, int64_t row_lo, int64_t row_hi) {
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  const int64_t n_rows = (int64_t)e->cfg.item_num + 1;
  if (row_lo < 0 || row_hi > n_rows || row_lo >= row_hi) REC_FAIL(e, REC_EINVAL, "rec_set_embedding_shard: bad row range");
  if (e->cfg.embedding_dim % 4) REC_FAIL(e, REC_EINVAL, "rec_set_embedding_shard: embedding_dim must be a multiple of 4");
  const bool all = row_lo == 0 && row_hi == n_rows;
  e->emb_row_lo = all ? 0 : row_lo;
  e->emb_row_hi = all ? 0 : row_hi;
  drop_step_graphs(e);  // a captured step bakes the swept row range in
  return REC_OK;
}

static int emb_rows_call(rec_engine *e, int net_id, const int64_t *ids, int64_t n, float *rows, bool scatter) {
  DevGuard dev_guard(e);
  int rc = check_net(e, net_id, false);
  if (rc) return rc;
  if (n < 0 || (n > 0 && (!ids || !rows))) REC_FAIL(e, REC_EINVAL, "rec_emb_rows_%s: bad argument", scatter ? "scatter" : "gather");
  return launch_emb_rows(e, net_id, ids, n, rows, scatter);
}
extern "C" int rec_emb_rows_gather(rec_engine *e, int net_id, const int64_t *ids, int64_t n, float *rows_out) {
  return emb_rows_call(e, net_id, ids, n, rows_out, false);
}
extern "C" int rec_emb_rows_scatter(rec_engine *e, int net_id, const int64_t *ids, int64_t n, const float *rows) {
  if (e) e->param_epoch++;  // writes parameters: derived operand images are stale (rec_eval_hold_params)
  return emb_rows_call(e, net_id, ids, n, const_cast<float *>(rows), true);
}

extern "C" int rec_set_cuda_graphs(rec_engine *e, int on) {
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  e->use_graph = on != 0;
  return REC_OK;
}
extern "C" int rec_set_tensor_cores(rec_engine *e, int on) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  e->use_tc = on != 0;
  return REC_OK;
}
extern "C" int rec_debug_set_trace(rec_engine *e, long long *dev_buf) {
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  e->trace = dev_buf;
  return REC_OK;
}
extern "C" int rec_debug_copy_astar(rec_engine *e, int32_t *out, int B) {
  DevGuard dev_guard(e);
  if (!e || !out || B < 1 || B > e->cfg.max_batch) return REC_EINVAL;
  REC_CUDA(e, cudaMemcpyAsync(out, e->astar, sizeof(int32_t) * (size_t)B, cudaMemcpyDeviceToDevice, e->stream));
  return REC_OK;
}
extern "C" int64_t rec_launch_count(const rec_engine *e) {
  DevGuard dev_guard(e); return e ? e->launches : -1; }
extern "C" int rec_enable_kernel_timing(rec_engine *e, int on) {
  DevGuard dev_guard(e); if (!e) return REC_EINVAL; e->timing = on != 0; return REC_OK; }
extern "C" float rec_last_kernel_ms(rec_engine *e, int which) {
  DevGuard dev_guard(e);
  if (!e || which < 0 || which > 5) return -1.f;
  float ms = -1.f;
  if (cudaEventSynchronize(e->ev[2 * which + 1]) != cudaSuccess ||
      cudaEventElapsedTime(&ms, e->ev[2 * which], e->ev[2 * which + 1]) != cudaSuccess) {
    cudaGetLastError();  // events never recorded (kernel did not run in timing mode): not an engine error
    return -1.f;
  }
  return ms;
}

static int check_batch(rec_engine *e, const rec_batch *b, bool q) {
  if (!b || !b->s || !b->a || !b->true_len) REC_FAIL(e, REC_EINVAL, "batch: s/a/true_len must be non-null");
  if (b->B < 1 || b->B > e->cfg.max_batch) REC_FAIL(e, REC_EINVAL, "batch size %d outside [1, max_batch=%d]", b->B, e->cfg.max_batch);
  if (q && (!b->r || !b->s_next || !b->true_next_len || !b->is_end)) REC_FAIL(e, REC_EINVAL, "batch: r/s_next/true_next_len/is_end must be non-null");
  return REC_OK;
}

extern "C" int rec_forward_state(rec_engine *e, int net_id, const int64_t *s, const int64_t *lengths, int B, float *h_out) {
  DevGuard dev_guard(e);
  int rc = check_net(e, net_id, false);
  if (rc) return rc;
  if (!s || !lengths || !h_out || B < 1) REC_FAIL(e, REC_EINVAL, "rec_forward_state: bad argument");
  return launch_gru_forward(e, net_id, s, lengths, B, h_out, false);
}

extern "C" int rec_head_logits(rec_engine *e, int net_id, int head, const float *h, int B, float *logits, int64_t ld) {
  DevGuard dev_guard(e);
  int rc = check_net(e, net_id, false);
  if (rc) return rc;
  if (head < 0 || head >= e->cfg.n_heads || !h || !logits || ld < e->Vloc) REC_FAIL(e, REC_EINVAL, "rec_head_logits: bad argument");
  return launch_head_logits(e, net_id, head, h, B, logits, ld);
}

// torch.optim.Adam (single-tensor/foreach, capturable=False) computes its bias corrections as python doubles:
//   step_size = lr / (1 - beta1^t),  denom = sqrt(v) / sqrt(1 - beta2^t) + eps.
// The step counter lives in DEVICE memory and the scalars are produced by a one-thread kernel that is the first
// node of every step: a replayed graph stays valid across steps, and steps may be queued asynchronously without a
// host buffer being overwritten before the GPU has read it.  The host mirrors the counter for rec_get_adam_step.
__global__ void adam_step_kernel(long long *__restrict__ d_step, float lr, float beta1, float beta2, float *__restrict__ sc) {
  const long long t = ++(*d_step);
  const double bc1 = 1.0 - pow((double)beta1, (double)t);
  const double bc2 = 1.0 - pow((double)beta2, (double)t);
  sc[0] = (float)((double)lr / bc1);
  sc[1] = __fdiv_rn(1.f, (float)sqrt(bc2));
}

static void adam_scalars(rec_engine *e, int net_id, const rec_train_hparams *hp, float *step_size, float *bc2_sqrt) {
  int64_t t = ++e->nets[net_id].adam_step;   // host mirror; same values as the kernel (by-value fallbacks only)
  double bc1 = 1.0 - pow((double)hp->beta1, (double)t);
  double bc2 = 1.0 - pow((double)hp->beta2, (double)t);
  *step_size = (float)((double)hp->lr / bc1);
  *bc2_sqrt = (float)sqrt(bc2);
}

static int upload_adam_scalars(rec_engine *e, int net_id, const rec_train_hparams *hp) {
  adam_step_kernel<<<1, 1, 0, e->stream>>>(e->d_step + net_id, hp->lr, hp->beta1, hp->beta2, e->d_sc);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

// ---- CUDA-graph replay of a whole train step --------------------------------------------------------------
__global__ void copy_batch_kernel(rec_batch src, rec_batch dst, int L) {
  const int B = src.B;
  const int n = B * (2 * L + 5);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i < B * L) ((int64_t *)dst.s)[i] = src.s[i];
    else if (i < 2 * B * L) { if (src.s_next) ((int64_t *)dst.s_next)[i - B * L] = src.s_next[i - B * L]; }
    else if (i < 2 * B * L + B) ((int64_t *)dst.a)[i - 2 * B * L] = src.a[i - 2 * B * L];
    else if (i < 2 * B * L + 2 * B) ((int64_t *)dst.true_len)[i - 2 * B * L - B] = src.true_len[i - 2 * B * L - B];
    else if (i < 2 * B * L + 3 * B) { if (src.true_next_len) ((int64_t *)dst.true_next_len)[i - 2 * B * L - 2 * B] = src.true_next_len[i - 2 * B * L - 2 * B]; }
    else if (i < 2 * B * L + 4 * B) { if (src.r) ((float *)dst.r)[i - 2 * B * L - 3 * B] = src.r[i - 2 * B * L - 3 * B]; }
    else { if (src.is_end) ((uint8_t *)dst.is_end)[i - 2 * B * L - 4 * B] = src.is_end[i - 2 * B * L - 4 * B]; }
  }
}

void rec_timeline_record(rec_engine *e, const char *file, int line) {
  if (e->tl_n >= 400) return;
  rec_engine::TlEntry &t = e->tl[e->tl_n++];
  if (!t.ev) cudaEventCreate(&t.ev);
  t.file = file; t.line = line;
  t.stream = e->stream == e->side[0] ? 1 : e->stream == e->side[1] ? 2 : 0;
  cudaEventRecord(t.ev, e->stream);
}

static void timeline_begin(rec_engine *e) {
  if (!e->tl_on) return;
  e->tl_n = 0;
  rec_timeline_record(e, "step-start", 0);
}

static void timeline_dump(rec_engine *e) {
  if (!e->tl_on) return;
  cudaDeviceSynchronize();
  if (++e->tl_steps % 50 != 0) return;
  fprintf(stderr, "[timeline] step %d\n", e->tl_steps);
  for (int i = 1; i < e->tl_n; ++i) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e->tl[0].ev, e->tl[i].ev);
    const char *f = strrchr(e->tl[i].file, '/');
    fprintf(stderr, "[timeline]   end %7.1f us  stream %d  %s:%d\n", ms * 1e3, e->tl[i].stream, f ? f + 1 : e->tl[i].file, e->tl[i].line);
  }
}

static uint64_t fnv(uint64_t h, const void *p, size_t n) {
  const unsigned char *c = (const unsigned char *)p;
  for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; }
  return h;
}

// Runs `body(own_batch)` either directly or as a replay of its captured graph.  Everything that changes from
// step to step lives in memory (engine-owned batch copy, Adam scalars), so one graph per (kind, main net, B,
// hyper-parameters, output pointer) serves every later step.
template <class Body>
// `host`: b holds HOST pointers.  The batch is packed into the pinned mirror, the step starts with one H2D copy
// of the block and ends with a D2H copy of `n_out` floats from `out` into h_loss -- both inside the graph.
static int run_step_graphed(rec_engine *e, int kind, int main_net, const rec_batch *b, const rec_train_hparams *hp,
                            float *out, Body body, bool host = false, int n_out = 0) {
  rec_batch own = e->own;
  own.B = b->B;
  if (!b->r) { own.r = nullptr; own.s_next = nullptr; own.true_next_len = nullptr; own.is_end = nullptr; }
  if (host) {
    const size_t B = (size_t)b->B, L = (size_t)e->cfg.state_size, mb = (size_t)e->cfg.max_batch;
    int64_t *p64 = (int64_t *)e->h_own;
    memcpy(p64, b->s, 8 * B * L);
    memcpy(p64 + 2 * mb * L, b->a, 8 * B);
    memcpy(p64 + 2 * mb * L + mb, b->true_len, 8 * B);
    if (b->r) {
      memcpy(p64 + mb * L, b->s_next, 8 * B * L);
      memcpy(p64 + 2 * mb * L + 2 * mb, b->true_next_len, 8 * B);
      float *pr = (float *)(p64 + mb * (2 * L + 3));
      memcpy(pr, b->r, 4 * B);
      memcpy((uint8_t *)(pr + mb), b->is_end, B);
    }
    kind |= 16;
  }
  auto run = [&](const rec_batch *bb) -> int {
    int rc = upload_adam_scalars(e, main_net, hp);
    if (rc) return rc;
    if (host) REC_CUDA(e, cudaMemcpyAsync(e->own_block, e->h_own, e->own_bytes, cudaMemcpyHostToDevice, e->stream));
    if ((rc = body(bb))) return rc;
    if (host) REC_CUDA(e, cudaMemcpyAsync(e->h_loss, out, sizeof(float) * n_out, cudaMemcpyDeviceToHost, e->stream));
    return REC_OK;
  };
  if (!e->use_graph || e->timing || e->trace || e->tl_on) {
    timeline_begin(e);
    int rc = run(host ? &own : b);
    timeline_dump(e);
    return rc;
  }
  if (!host) {
    copy_batch_kernel<<<cdiv(b->B * (2 * e->cfg.state_size + 5), 256), 256, 0, e->stream>>>(*b, own, e->cfg.state_size);
    REC_LAUNCH_CHECK(e);
  }
  uint64_t key = 1469598103934665603ull;
  key = fnv(key, &kind, sizeof(kind)); key = fnv(key, &main_net, sizeof(main_net)); key = fnv(key, &b->B, sizeof(int));
  {  // field by field: the struct's padding bytes are indeterminate in a caller's stack object
    const float f[] = {hp->lr, hp->beta1, hp->beta2, hp->eps, hp->gamma, hp->alpha, hp->q_weights[0], hp->q_weights[1],
                       hp->q_weights[2], hp->nov_reward, hp->dropout_p};
    const int32_t i[] = {hp->div_dim, hp->topk_div, hp->topk_nov, hp->pad_pos_end, e->use_tc ? 1 : 0, e->overlap ? 1 : 0};
    const void *ptrs[] = {hp->div_emb, hp->unpopular, hp->out_to_in, hp->dropout_mask, out};
    key = fnv(key, f, sizeof(f)); key = fnv(key, i, sizeof(i)); key = fnv(key, ptrs, sizeof(ptrs));
    key = fnv(key, &hp->dropout_seed, sizeof(hp->dropout_seed));
  }
  for (int n = 0; n < e->cfg.n_nets; ++n) key = fnv(key, &e->nets[n].p, sizeof(rec_net_params));
  rec_engine::GraphEntry *g = nullptr;
  for (int i = 0; i < e->n_graphs; ++i) if (e->graphs[i].key == key) g = &e->graphs[i];
  if (!g) {
    if (e->n_graphs == 16) {  // recycle the oldest entry
      if (e->graphs[0].exec) cudaGraphExecDestroy((cudaGraphExec_t)e->graphs[0].exec);
      for (int i = 1; i < 16; ++i) e->graphs[i - 1] = e->graphs[i];
      e->n_graphs = 15;
    }
    g = &e->graphs[e->n_graphs++];
    g->key = key; g->exec = nullptr; g->launches = 0; g->seen = 0;
  }
  if (!g->exec && g->seen < 1) {  // first sighting: run eagerly (also sets kernel attributes outside any capture)
    g->seen++;
    return run(&own);
  }
  if (!g->exec) {
    const int64_t l0 = e->launches;
    cudaGraph_t graph = nullptr;
    // capture on the private stream (nothing executes); the instantiated graph is replayed on the caller's stream
    cudaStream_t user_stream = e->stream;
    REC_CUDA(e, cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    e->stream = e->cap_stream;
    int rc = run(&own);
    cudaError_t st = cudaStreamEndCapture(e->cap_stream, &graph);
    e->stream = user_stream;
    if (rc || st != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      if (!rc) REC_FAIL(e, REC_ECUDA, "CUDA graph capture failed: %s", cudaGetErrorString(st));
      return rc;
    }
    cudaGraphExec_t exec = nullptr;
    st = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (st != cudaSuccess) REC_FAIL(e, REC_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(st));
    g->exec = exec;
    g->launches = (int)(e->launches - l0);
    e->launches = l0;  // nothing ran during capture
  }
  REC_CUDA(e, cudaGraphLaunch((cudaGraphExec_t)g->exec, e->stream));
  e->launches += g->launches;
  return REC_OK;
}

// GRU backward + embedding update with the two independent chains overlapped:
//   main: BPTT -> weight gradients -> GRU Adam        side: duplicate sums (after BPTT) -> table Adam (after wgrad)
// `ranked`: stage 1 of the embedding update was already issued (it needs only the batch).
static int trunk_backward(rec_engine *e, int net, const int64_t *s, const int64_t *lens, int B, const float *dh,
                          float step_size, float bc2_sqrt, const rec_train_hparams *hp, bool ranked) {
  int rc;
  if (!ranked) {
    SideScope side(e, 1);
    if ((rc = launch_embedding_update(e, net, s, lens, B, step_size, bc2_sqrt, hp, 1))) return rc;
  }
  if ((rc = launch_gru_backward(e, net, s, lens, B, dh, step_size, bc2_sqrt, hp, 1))) return rc;
  {
    SideScope side(e, 1);
    if ((rc = launch_embedding_update(e, net, s, lens, B, step_size, bc2_sqrt, hp, 2))) return rc;
  }
  if ((rc = launch_gru_backward(e, net, s, lens, B, dh, step_size, bc2_sqrt, hp, 2))) return rc;
  {
    SideScope side(e, 1);  // the weight gradients read the table: its Adam sweep waits for them
    if ((rc = launch_embedding_update(e, net, s, lens, B, step_size, bc2_sqrt, hp, 4))) return rc;
  }
  if ((rc = launch_gru_backward(e, net, s, lens, B, dh, step_size, bc2_sqrt, hp, 4))) return rc;
  side_join(e, 1);
  return REC_OK;
}

static int supervised_body(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp, float *loss_out,
                           float step_size, float bc2_sqrt) {
  int rc;
  const int B = b->B;
  const bool drop = hp->dropout_p > 0.f;
  e->hpack_ready = false;
  if (tck_heads_supported(e)) {  // weight image of the head: packed next to the GRU forward
    SideScope side(e, 0);
    if ((rc = tck_prepack_heads(e, 0, 0, nullptr))) return rc;
  }
  if ((rc = launch_gru_forward(e, 0, b->s, b->true_len, B, e->h_state[0], true))) return rc;
  if (tck_heads_supported(e)) side_join(e, 0);
  if (drop && (rc = launch_dropout(e, 0, e->h_state[0], nullptr, B, hp, false))) return rc;  // heads see the dropped state
  HeadStatsArgs a = {};
  a.net_id = 0; a.h = e->h_state[0]; a.B = B; a.do_stats = 1; a.stats_head = 0; a.target = b->a;
  int n_split = 0;
  if ((rc = head_stats_dispatch(e, a, &n_split))) return rc;
  if ((rc = launch_head_merge(e, e->part, n_split, B, 0, true, false))) return rc;
  if ((rc = launch_loss_reduce(e, B, nullptr, e->loss_buf))) return rc;
  REC_CUDA(e, cudaMemcpyAsync(loss_out, e->loss_buf, sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
  if ((rc = launch_head_backward_adam(e, 0, e->h_state[0], b, B, step_size, bc2_sqrt, hp, 1.f / (float)B))) return rc;
  if (drop && (rc = launch_dropout(e, 0, nullptr, e->dh, B, hp, true))) return rc;           // dL/dh through the mask
  return trunk_backward(e, 0, b->s, b->true_len, B, e->dh, step_size, bc2_sqrt, hp, false);
}

extern "C" int rec_train_step_supervised(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp, float *loss_out) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  int rc = check_net(e, 0, true);
  if (rc) return rc;
  if ((rc = check_batch(e, b, false))) return rc;
  if (!hp || !loss_out) REC_FAIL(e, REC_EINVAL, "rec_train_step_supervised: null argument");
  if (hp->dropout_p < 0.f || hp->dropout_p >= 1.f) REC_FAIL(e, REC_EINVAL, "dropout_p must be in [0, 1)");
  if (e->Vloc != e->cfg.action_dim) REC_FAIL(e, REC_EINVAL, "sharded engine: use the rec_train_phase_* entry points");
  float step_size, bc2_sqrt;
  adam_scalars(e, 0, hp, &step_size, &bc2_sqrt);
  return run_step_graphed(e, 0, 0, b, hp, loss_out, [&](const rec_batch *bb) {
    return supervised_body(e, bb, hp, loss_out, step_size, bc2_sqrt);
  });
}

static int finish_host_step(rec_engine *e, int rc, float *out, int n) {
  if (rc) return rc;
  REC_CUDA(e, cudaStreamSynchronize(e->stream));
  for (int i = 0; i < n; ++i) out[i] = e->h_loss[i];
  return REC_OK;
}

extern "C" int rec_train_step_supervised_host(rec_engine *e, const rec_batch *host_b, const rec_train_hparams *hp, float *loss_host) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  int rc = check_net(e, 0, true);
  if (rc) return rc;
  if ((rc = check_batch(e, host_b, false))) return rc;
  if (!hp || !loss_host) REC_FAIL(e, REC_EINVAL, "rec_train_step_supervised_host: null argument");
  if (hp->dropout_p < 0.f || hp->dropout_p >= 1.f) REC_FAIL(e, REC_EINVAL, "dropout_p must be in [0, 1)");
  if (e->Vloc != e->cfg.action_dim) REC_FAIL(e, REC_EINVAL, "sharded engine: use the rec_train_phase_* entry points");
  float step_size, bc2_sqrt;
  adam_scalars(e, 0, hp, &step_size, &bc2_sqrt);
  float *out = e->loss_buf + 4;
  rec_batch hb = *host_b;
  hb.r = nullptr;  // supervised step: only s / a / true_len are read
  rc = run_step_graphed(e, 0, 0, &hb, hp, out, [&](const rec_batch *bb) {
    return supervised_body(e, bb, hp, out, step_size, bc2_sqrt);
  }, true, 1);
  return finish_host_step(e, rc, loss_host, 1);
}

// One fused SQN / SMORL step.  Branch structure (each branch is a stream; parallel branches under graph capture):
//   main   : GRU fwd -> supervised stats -> [mark 0] -> greedy-action stats -> fused per-row Q kernel -> [mark 1]
//            -> (after mark 2) dh reduce -> GRU backward / weight update            (critical path, top priority)
//   side 0 : (after mark 0) supervised-head backward + Adam (tensor cores, latency-bound) -> [mark 2]
//   side 1 : token ordering (start of step), duplicate sums, embedding-table Adam    (see trunk_backward)
//   side 2 : (after mark 1) losses, Q-head gradient rows -> (after mark 2) streaming Adam sweep of the Q heads.
//            The sweep saturates the HBM: it starts when the supervised-head kernel is done and shares the machine
//            with the small GRU-backward kernels, at the lowest priority.
static int q_step_body(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp, int main_net, float *losses_out,
                       float step_size, float bc2_sqrt) {
  int rc;
  const int B = b->B, boot = 1 - main_net, n_q = e->cfg.n_heads - 1;
  e->hpack_ready = false;
  {  // ordering of the token positions for the embedding backward: needs only the batch
    SideScope side(e, 1);
    if ((rc = launch_embedding_update(e, main_net, b->s, b->true_len, B, step_size, bc2_sqrt, hp, 1))) return rc;
  }
  if (tck_heads_supported(e)) {  // weight images of the heads (D >= 128): packed next to the GRU forward
    SideScope side(e, 0);
    const float wq[3] = {n_q == 3 ? hp->q_weights[0] : 1.f, hp->q_weights[1], hp->q_weights[2]};
    if ((rc = tck_prepack_heads(e, main_net, n_q, wq))) return rc;
  }
  // three GRU passes: main(s, len) [saved], main(s', len'), boot(s', len)  -- (q1) boot sees true_len
  {
    const int nets[3] = {main_net, main_net, boot};
    const int64_t *ss[3] = {b->s, b->s_next, b->s_next};
    const int64_t *ll[3] = {b->true_len, b->true_next_len, b->true_len};
    float *hh[3] = {e->h_state[0], e->h_state[1], e->h_state[2]};
    const bool sv[3] = {true, false, false};
    if ((rc = launch_gru_forward_multi(e, 3, nets, ss, ll, hh, sv, B))) return rc;
  }
  if (tck_heads_supported(e)) side_join(e, 0);
  if (tc_bwd_supported(e, B)) {  // operand image of the supervised-head backward: only needs the forward pass
    SideScope side(e, 0);
    if ((rc = launch_h_prepack_early(e, e->h_state[0], B))) return rc;
  }
  // supervised head statistics (+ top-k of the supervised logits for the SMORL rewards)
  HeadStatsArgs a = {};
  a.net_id = main_net; a.h = e->h_state[0]; a.B = B; a.do_stats = 1; a.stats_head = 0; a.target = b->a;
  a.topk = (n_q == 3) ? (hp->topk_div > hp->topk_nov ? hp->topk_div : hp->topk_nov) : 0;
  int n_split = 0;
  if (e->timing) cudaEventRecord(e->ev[8], e->stream);
  if ((rc = head_stats_dispatch(e, a, &n_split))) return rc;
  if (e->timing) cudaEventRecord(e->ev[9], e->stream);
  if ((rc = launch_head_merge(e, e->part, n_split, B, a.topk, true, false, nullptr, &a))) return rc;
  side_mark(e, 0);  // log-sum-exp of the supervised logits is final: its backward may start
  // greedy action a* = argmax_a sum_h w_h Q_h(s', a) on the main net
  HeadStatsArgs g = {};
  g.net_id = main_net; g.h = e->h_state[1]; g.B = B; g.n_arg = n_q;
  g.w[0] = n_q == 3 ? hp->q_weights[0] : 1.f; g.w[1] = hp->q_weights[1]; g.w[2] = hp->q_weights[2];
  if (e->timing) cudaEventRecord(e->ev[10], e->stream);
  if ((rc = head_stats_dispatch(e, g, &n_split))) return rc;
  if (e->timing) cudaEventRecord(e->ev[11], e->stream);
  {  // issued after the greedy-action statistics so that those get the SMs first
    SideScope side(e, 0, 0);
    if ((rc = launch_sup_head_bwd(e, main_net, e->h_state[0], b, B, step_size, bc2_sqrt, hp, 1.f / (float)B))) return rc;
    side_mark(e, 2);  // supervised head done: dh slices final, HBM free for the Q-head sweep
  }
  // a*, Q(s,a), Q_boot(s',a*), rewards, TD target, dq, Q-head dh slice: one launch
  const float alpha_eff = (n_q == 3) ? hp->alpha : 1.f;
  if ((rc = launch_q_rows_fused(e, main_net, b, hp, n_split, alpha_eff, extra(e).q_loss_rows, &g))) return rc;
  side_mark(e, 1);
  {
    SideScope side(e, 2, 1);
    if ((rc = launch_loss_reduce(e, B, extra(e).q_loss_rows, e->loss_buf))) return rc;
    REC_CUDA(e, cudaMemcpyAsync(losses_out, e->loss_buf, 2 * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
    static const int sweep_mark = getenv("REC_SWEEP_EARLY") ? -1 : 2;
    // (wide heads, D >= 128: the sweep also waits for the supervised head's dW + Adam kernel -- started together, the two
    // HBM-bound kernels fill every SM's register file and the GRU backward, which needs most of an SM per CTA, queues
    // behind both: 4.6 vs 3.8 ms per step at cfg3)
    if ((rc = launch_q_heads_update(e, main_net, e->h_state[0], b, B, step_size, bc2_sqrt, hp,
                                    tck_heads_supported(e) && sweep_mark >= 0 && getenv("REC_SWEEP_WITH_DW") ? 3 : sweep_mark))) return rc;
  }
  side_wait_mark(e, tck_heads_supported(e) ? 3 : 2);
  if ((rc = launch_dh_reduce(e, B))) return rc;
  if ((rc = trunk_backward(e, main_net, b->s, b->true_len, B, e->dh, step_size, bc2_sqrt, hp, true))) return rc;
  side_join(e, 0);
  side_join(e, 2);
  return REC_OK;
}

extern "C" int rec_train_step_q(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp, int main_net, float *losses_out) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  if (e->cfg.n_nets != 2 || e->cfg.n_heads < 2) REC_FAIL(e, REC_EINVAL, "rec_train_step_q needs a twin-net engine with Q heads");
  if (main_net != 0 && main_net != 1) REC_FAIL(e, REC_EINVAL, "main_net must be 0 or 1");
  int rc = check_net(e, main_net, true);
  if (rc) return rc;
  if ((rc = check_net(e, 1 - main_net, false))) return rc;
  if ((rc = check_batch(e, b, true))) return rc;
  if (!hp || !losses_out) REC_FAIL(e, REC_EINVAL, "rec_train_step_q: null argument");
  if (e->Vloc != e->cfg.action_dim) REC_FAIL(e, REC_EINVAL, "sharded engine: use the rec_train_phase_* entry points");
  const int n_q = e->cfg.n_heads - 1;
  if (n_q == 3) {
    if (!hp->div_emb || !hp->unpopular || hp->topk_div < 1 || hp->topk_nov < 1 || hp->div_dim < 1 ||
        hp->topk_div > e->cfg.max_topk || hp->topk_nov > e->cfg.max_topk)
      REC_FAIL(e, REC_EINVAL, "SMORL step needs div_emb, unpopular and 1 <= topk_div/topk_nov <= max_topk");
  }
  float step_size, bc2_sqrt;
  adam_scalars(e, main_net, hp, &step_size, &bc2_sqrt);
  return run_step_graphed(e, 1, main_net, b, hp, losses_out, [&](const rec_batch *bb) {
    return q_step_body(e, bb, hp, main_net, losses_out, step_size, bc2_sqrt);
  });
}

extern "C" int rec_train_step_q_host(rec_engine *e, const rec_batch *host_b, const rec_train_hparams *hp, int main_net,
                                     float *losses_host) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  if (e->cfg.n_nets != 2 || e->cfg.n_heads < 2) REC_FAIL(e, REC_EINVAL, "rec_train_step_q_host needs a twin-net engine with Q heads");
  if (main_net != 0 && main_net != 1) REC_FAIL(e, REC_EINVAL, "main_net must be 0 or 1");
  int rc = check_net(e, main_net, true);
  if (rc) return rc;
  if ((rc = check_net(e, 1 - main_net, false))) return rc;
  if ((rc = check_batch(e, host_b, true))) return rc;
  if (!hp || !losses_host) REC_FAIL(e, REC_EINVAL, "rec_train_step_q_host: null argument");
  if (e->Vloc != e->cfg.action_dim) REC_FAIL(e, REC_EINVAL, "sharded engine: use the rec_train_phase_* entry points");
  const int n_q = e->cfg.n_heads - 1;
  if (n_q == 3) {
    if (!hp->div_emb || !hp->unpopular || hp->topk_div < 1 || hp->topk_nov < 1 || hp->div_dim < 1 ||
        hp->topk_div > e->cfg.max_topk || hp->topk_nov > e->cfg.max_topk)
      REC_FAIL(e, REC_EINVAL, "SMORL step needs div_emb, unpopular and 1 <= topk_div/topk_nov <= max_topk");
  }
  float step_size, bc2_sqrt;
  adam_scalars(e, main_net, hp, &step_size, &bc2_sqrt);
  float *out = e->loss_buf + 4;
  rc = run_step_graphed(e, 1, main_net, host_b, hp, out, [&](const rec_batch *bb) {
    return q_step_body(e, bb, hp, main_net, out, step_size, bc2_sqrt);
  }, true, 2);
  return finish_host_step(e, rc, losses_host, 2);
}

// ---- SARM (models/SARM/sarm.py:117-149): one net, five Q heads, head 0 doubles as the supervised head ----------------
__global__ void copy_stat_kernel(const float *__restrict__ row_stats, int col, int B, float *__restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) out[b] = row_stats[(int64_t)b * 8 + col];
}

// One warp per session: Q_i(s, a) for the five heads (fp32 row dots), TD targets y_i = r + gamma * max_a Q_i(s', a)
// (the reference masks nothing with is_end here: sarm.py:133-135), per-row loss sum_i (y_i - Q_i)^2 / 5, the gradients
// dQ_i = (2/5) (Q_i - y_i) / B, and the dh contribution of heads 1..4 (head 0's travels through its dense backward:
// `extra`).  Runs before any Adam kernel touches the head weights.
struct SarmPtrs { const float *w[5], *b[5]; };
__global__ void __launch_bounds__(256) sarm_rows_entry(SarmPtrs P, const float *__restrict__ h, const int64_t *__restrict__ a,
                                                       const float *__restrict__ r, const float *__restrict__ qmax, int qmax_ld,
                                                       int B, int D, int Vloc, int vocab_lo, float gamma, float *__restrict__ dq,
                                                       float *__restrict__ extra, float *__restrict__ q_loss_rows,
                                                       float *__restrict__ dh_slice);

static int sarm_step_body(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp, float *losses_out, float step_size,
                          float bc2_sqrt) {
  int rc;
  const int B = b->B;
  e->hpack_ready = false;
  {  // token ordering of the embedding backward: needs only the batch
    SideScope side(e, 1);
    if ((rc = launch_embedding_update(e, 0, b->s, b->true_len, B, step_size, bc2_sqrt, hp, 1))) return rc;
  }
  {  // two passes of the same net: s (saved for the backward) and s' (values only)
    const int nets[2] = {0, 0};
    const int64_t *ss[2] = {b->s, b->s_next};
    const int64_t *ll[2] = {b->true_len, b->true_next_len};
    float *hh[2] = {e->h_state[0], e->h_state[1]};
    const bool sv[2] = {true, false};
    if ((rc = launch_gru_forward_multi(e, 2, nets, ss, ll, hh, sv, B))) return rc;
  }
  int n_split = 0;
  // max_a Q_i(s', a) for the five heads: greedy-action pass per head, merge with the fp32 re-score -> exact maxima
  for (int i = 0; i < 5; ++i) {
    HeadStatsArgs g = {};
    g.net_id = 0; g.h = e->h_state[1]; g.B = B; g.n_arg = 1; g.w[0] = 1.f; g.arg_shift = i - 1;
    if ((rc = head_stats_dispatch(e, g, &n_split))) return rc;
    if ((rc = launch_head_merge(e, e->part, n_split, B, 0, false, true, nullptr, &g))) return rc;
    copy_stat_kernel<<<cdiv(B, 256), 256, 0, e->stream>>>(e->row_stats, 2, B, e->sarm_qmax + (int64_t)i * e->cfg.max_batch);
    REC_LAUNCH_CHECK(e);
  }
  // cross-entropy statistics of head 0 on s
  HeadStatsArgs a = {};
  a.net_id = 0; a.h = e->h_state[0]; a.B = B; a.do_stats = 1; a.stats_head = 0; a.target = b->a;
  if ((rc = head_stats_dispatch(e, a, &n_split))) return rc;
  if ((rc = launch_head_merge(e, e->part, n_split, B, 0, true, false, nullptr, &a))) return rc;
  // per-row Q terms
  {
    const rec_net_params &p = e->nets[0].p;
    SarmPtrs P;
    for (int i = 0; i < 5; ++i) { P.w[i] = p.head_w[i]; P.b[i] = p.head_b[i]; }
    float *q_slice = e->dh_part + (int64_t)head_bwd_dense_slices(e, B) * B * e->D;
    sarm_rows_entry<<<cdiv(B, 8), 256, 0, e->stream>>>(P, e->h_state[0], b->a, b->r, e->sarm_qmax, e->cfg.max_batch, B, e->D, e->Vloc,
                                                       e->cfg.vocab_lo, hp->gamma, e->sarm_dq, e->sarm_extra, extra(e).q_loss_rows, q_slice);
    REC_LAUNCH_CHECK(e);
  }
  if ((rc = launch_loss_reduce(e, B, extra(e).q_loss_rows, e->loss_buf))) return rc;
  REC_CUDA(e, cudaMemcpyAsync(losses_out, e->loss_buf, 2 * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
  // head 0: dense backward + Adam with its Q gradient added at the target column; heads 1..4: row-sparse sweep
  e->bwd_extra = e->sarm_extra;
  rc = launch_sup_head_bwd(e, 0, e->h_state[0], b, B, step_size, bc2_sqrt, hp, 1.f / (float)B);
  e->bwd_extra = nullptr;
  if (rc) return rc;
  if ((rc = launch_q_heads_adam_ex(e, 0, e->h_state[0], b, B, step_size, bc2_sqrt, hp, 1, 4, e->sarm_dq + 1, 5))) return rc;
  if ((rc = launch_dh_reduce(e, B))) return rc;
  return trunk_backward(e, 0, b->s, b->true_len, B, e->dh, step_size, bc2_sqrt, hp, true);
}

__global__ void __launch_bounds__(256) sarm_rows_entry(SarmPtrs P, const float *__restrict__ h, const int64_t *__restrict__ a,
                                                       const float *__restrict__ r, const float *__restrict__ qmax, int qmax_ld,
                                                       int B, int D, int Vloc, int vocab_lo, float gamma, float *__restrict__ dq,
                                                       float *__restrict__ extra, float *__restrict__ q_loss_rows,
                                                       float *__restrict__ dh_slice) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int64_t loc = a[b] - vocab_lo;
  const bool here = loc >= 0 && loc < Vloc;
  const float *hr = h + (int64_t)b * D;
  float g[5];
  float loss = 0.f;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    float q = 0.f;
    if (here) {
      const float *wr = P.w[i] + loc * D;
      float acc = 0.f;
      for (int k = lane; k < D; k += 32) acc = fmaf(hr[k], __ldg(wr + k), acc);
      q = warp_sum(acc) + __ldg(P.b[i] + loc);
    }
    const float y = r[b] + gamma * qmax[(int64_t)i * qmax_ld + b];
    const float diff = y - q;
    loss += diff * diff;
    g[i] = here ? 0.2f * 2.f * (-diff) / (float)B : 0.f;
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 5; ++i) dq[b * 5 + i] = g[i];
    extra[b] = g[0];
    q_loss_rows[b] = 0.2f * loss;
  }
  for (int k = lane; k < D; k += 32) {
    float acc = 0.f;
    if (here) {
#pragma unroll
      for (int i = 1; i < 5; ++i) acc = fmaf(g[i], __ldg(P.w[i] + loc * D + k), acc);
    }
    dh_slice[(int64_t)b * D + k] = acc;
  }
}

extern "C" int rec_train_step_sarm(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp, float *losses_out) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  if (e->cfg.n_heads != 5 || e->cfg.n_nets != 1) REC_FAIL(e, REC_EINVAL, "rec_train_step_sarm needs a single-net engine with 5 heads");
  int rc = check_net(e, 0, true);
  if (rc) return rc;
  if ((rc = check_batch(e, b, true))) return rc;
  if (!hp || !losses_out) REC_FAIL(e, REC_EINVAL, "rec_train_step_sarm: null argument");
  if (e->Vloc != e->cfg.action_dim) REC_FAIL(e, REC_EINVAL, "rec_train_step_sarm: vocabulary-sharded engines are not supported");
  float step_size, bc2_sqrt;
  adam_scalars(e, 0, hp, &step_size, &bc2_sqrt);
  return run_step_graphed(e, 2, 0, b, hp, losses_out, [&](const rec_batch *bb) {
    return sarm_step_body(e, bb, hp, losses_out, step_size, bc2_sqrt);
  });
}

static int eval_kmax(const rec_eval_opts *o) {
  int k = 1;
  for (int i = 0; i < o->n_k; ++i) k = o->ks[i] > k ? o->ks[i] : k;
  for (int i = 0; i < o->n_cov; ++i) k = o->cov_ks[i] > k ? o->cov_ks[i] : k;
  if (o->div_emb && o->topk_div > k) k = o->topk_div;
  if (o->unpopular && o->topk_nov > k) k = o->topk_nov;
  return k;
}

extern "C" int rec_eval_hold_params(rec_engine *e, int on) {
  if (!e) return REC_EINVAL;
  e->k_hold = on != 0;
  e->k_img_epoch = -1;  // the first evaluation call after a change of state packs afresh
  return REC_OK;
}

extern "C" int rec_eval_batch(rec_engine *e, int net_id, const rec_batch *b, const rec_eval_opts *o,
                              const rec_eval_accum *acc, int32_t *topk_ids, float *topk_scores) {
  DevGuard dev_guard(e);
  int rc = check_net(e, net_id, false);
  if (rc) return rc;
  if ((rc = check_batch(e, b, false))) return rc;
  if (!o || !acc || !acc->hits || !acc->ndcg || !acc->reps || !acc->div_sum || !acc->nov_sum || !acc->loss_sum || !acc->cov_bits)
    REC_FAIL(e, REC_EINVAL, "rec_eval_batch: null option/accumulator");
  if (o->head_idx < 0 || o->head_idx >= e->cfg.n_heads || o->n_k < 0 || o->n_k > REC_MAX_KLIST || o->n_cov < 0 || o->n_cov > REC_MAX_KLIST)
    REC_FAIL(e, REC_EINVAL, "rec_eval_batch: bad head_idx / k lists");
  const int kmax = eval_kmax(o);
  if (kmax > e->cfg.max_topk) REC_FAIL(e, REC_EINVAL, "rec_eval_batch: k=%d exceeds max_topk=%d", kmax, e->cfg.max_topk);
  if (e->cfg.vocab_lo != 0 || e->cfg.vocab_hi != e->cfg.action_dim) REC_FAIL(e, REC_EINVAL, "rec_eval_batch on a sharded engine: use rec_eval_shard_candidates + rec_eval_merge");
  const int B = b->B;
  if ((rc = launch_gru_forward(e, net_id, b->s, b->true_len, B, e->h_state[0], false))) return rc;
  HeadStatsArgs a = {};
  a.net_id = net_id; a.h = e->h_state[0]; a.B = B; a.do_stats = 1; a.stats_head = o->head_idx; a.target = b->a; a.topk = kmax;
  int n_split = 0;
  if (tck_chunk_topk_supported(e, a)) {
    // large batch x large catalogue: chunk maxima on the tensor cores, exact top-k from the best chunks (heads_tck.cu)
    if ((rc = launch_head_topk_chunks(e, a, &n_split, nullptr))) return rc;
    if ((rc = launch_head_merge(e, e->part, n_split, B, 0, true, false, nullptr, &a))) return rc;
    return launch_eval_metrics(e, b, o, kmax, acc, extra(e).rowm, topk_ids, topk_scores);
  }
  if (e->timing) cudaEventRecord(e->ev[2], e->stream);
  if ((rc = head_stats_dispatch(e, a, &n_split))) return rc;
  if (e->timing) cudaEventRecord(e->ev[3], e->stream);
  if ((rc = launch_head_merge(e, e->part, n_split, B, kmax, true, false, nullptr, &a))) return rc;
  return launch_eval_metrics(e, b, o, kmax, acc, extra(e).rowm, topk_ids, topk_scores);
}

// ---- sharded / phase-split entry points (see DESIGN.md "multi-GPU") --------------------------------
// Every rank holds the FULL (all-gathered) batch, a replica of embedding + GRU, and rows
// [vocab_lo, vocab_hi) of every head.  Collectives are run by the caller between the phases.

static int shard_head_pass(rec_engine *e, int net_id, const float *h, const rec_batch *b, int stats_head, int topk,
                           int n_q, const float *w, bool want_stats, float *summary) {
  // per-shard statistics -> local merge -> one record per row in e->summary
  int rc, n_split = 0;
  if (want_stats || topk > 0) {
    HeadStatsArgs a = {};
    a.net_id = net_id; a.h = h; a.B = b->B; a.do_stats = want_stats ? 1 : 0; a.stats_head = stats_head; a.target = b->a;
    a.topk = topk;
    if (want_stats && n_q == 0 && tck_chunk_topk_supported(e, a)) {
      if ((rc = launch_head_topk_chunks(e, a, &n_split, summary))) return rc;
      if ((rc = launch_head_merge(e, e->part, n_split, b->B, 0, true, false, summary, &a))) return rc;
    } else {
      if ((rc = head_stats_dispatch(e, a, &n_split))) return rc;
      if ((rc = launch_head_merge(e, e->part, n_split, b->B, topk, want_stats, false, summary, &a))) return rc;
    }
  }
  if (n_q > 0) {
    HeadStatsArgs g = {};
    g.net_id = net_id; g.h = e->h_state[1]; g.B = b->B; g.n_arg = n_q;
    g.w[0] = w[0]; g.w[1] = w[1]; g.w[2] = w[2];
    if ((rc = head_stats_dispatch(e, g, &n_split))) return rc;
    if ((rc = launch_head_merge(e, e->part, n_split, b->B, 0, false, true, summary, &g))) return rc;
  }
  return REC_OK;
}

extern "C" int rec_record_floats(const rec_engine *e) {
  DevGuard dev_guard(e); return e ? e->part_stride : -1; }

static int phase_a_body(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp, int main_net, float *records_out,
                        bool skip_gru) {
  if (!e) return REC_EINVAL;
  if (!b || !hp || !records_out) REC_FAIL(e, REC_EINVAL, "rec_train_phase_a: null argument");
  if (hp->dropout_p < 0.f || hp->dropout_p >= 1.f) REC_FAIL(e, REC_EINVAL, "dropout_p must be in [0, 1)");
  const int n_q = e->cfg.n_heads - 1;
  if (main_net < 0 || main_net >= e->cfg.n_nets) REC_FAIL(e, REC_EINVAL, "main_net out of range");
  int rc = check_net(e, main_net, true);
  if (rc) return rc;
  for (int i = 0; i < 3; ++i) side_join(e, i);  // a step abandoned after phase C may still have side work in flight
  if ((rc = check_batch(e, b, n_q > 0))) return rc;
  if (n_q > 0 && e->cfg.n_nets != 2) REC_FAIL(e, REC_EINVAL, "Q heads need a twin-net engine");
  if (n_q == 3 && (!hp->div_emb || !hp->unpopular || hp->topk_div < 1 || hp->topk_nov < 1 ||
                   hp->topk_div > e->cfg.max_topk || hp->topk_nov > e->cfg.max_topk))
    REC_FAIL(e, REC_EINVAL, "SMORL step needs div_emb, unpopular and 1 <= topk_div/topk_nov <= max_topk");
  const int B = b->B, boot = 1 - main_net;
  e->cur_batch = *b; e->cur_hp = *hp; e->cur_main = main_net; e->cur_phase = 1;
  e->cur_topk = (n_q == 3) ? (hp->topk_div > hp->topk_nov ? hp->topk_div : hp->topk_nov) : 0;
  REC_CUDA(e, cudaMemsetAsync(records_out, 0, sizeof(float) * (size_t)B * e->part_stride, e->stream));
  if (skip_gru) {
    if (n_q > 0 && (rc = check_net(e, boot, false))) return rc;  // final states came from rec_dp_unpack
  } else if (n_q > 0) {
    if ((rc = check_net(e, boot, false))) return rc;
    const int nets[3] = {main_net, main_net, boot};
    const int64_t *ss[3] = {b->s, b->s_next, b->s_next};
    const int64_t *ll[3] = {b->true_len, b->true_next_len, b->true_len};
    float *hh[3] = {e->h_state[0], e->h_state[1], e->h_state[2]};
    const bool sv[3] = {true, false, false};
    if ((rc = launch_gru_forward_multi(e, 3, nets, ss, ll, hh, sv, B))) return rc;
  } else {
    if ((rc = launch_gru_forward(e, main_net, b->s, b->true_len, B, e->h_state[0], true))) return rc;
  }
  // BidirGRU4Rec dropout (supervised step only): the keep mask is a function of (seed, Adam step, element of the GLOBAL
  // batch) or injected for the global batch, so every rank drops the same elements without a collective
  if (n_q == 0 && hp->dropout_p > 0.f && (rc = launch_dropout(e, main_net, e->h_state[0], nullptr, B, hp, false))) return rc;
  float w[3] = {n_q == 3 ? hp->q_weights[0] : 1.f, hp->q_weights[1], hp->q_weights[2]};
  return shard_head_pass(e, main_net, e->h_state[0], b, 0, e->cur_topk, n_q, w, true, records_out);
}

extern "C" int rec_train_phase_a(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp, int main_net,
                                 float *records_out) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  if (e) e->dp_active = false;
  return phase_a_body(e, b, hp, main_net, records_out, false);
}

// ---- data-parallel trunk of the vocabulary-sharded step --------------------------------------------------
// The heads are sharded by vocabulary and see the GLOBAL batch; embedding + GRU are replicated.  Running the
// recurrent trunk on the global batch on every rank would make its cost grow with the number of GPUs, so each
// rank runs it on its OWN sessions only:
//   rec_dp_forward   local GRU passes -> packed record [batch fields | final states]      -> all-gather
//   rec_dp_unpack    global batch fields + global final states
//   rec_train_phase_a_heads / _b / _c   (unchanged head phases on the global batch)       -> all-reduce dL/dh
//   rec_dp_backward  BPTT + weight gradients of the local sessions -> GRU gradient vector -> all-reduce
//                    and the local dx rows                                                 -> all-gather
//   rec_dp_apply     identical Adam update of the replicated GRU + embedding table on every rank
__global__ void pack_h_kernel(const float *h0, const float *h1, const float *h2, int n_h, int n, float *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_h * n) return;
  const int k = i / n, j = i - k * n;
  out[i] = (k == 0 ? h0 : k == 1 ? h1 : h2)[j];
}
__global__ void unpack_h_kernel(const uint8_t *gathered, size_t stride, size_t h_off, int G, int n_h, int n, float *h0,
                                float *h1, float *h2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * n_h * n) return;
  const int g = i / (n_h * n), r = i - g * (n_h * n), k = r / n, j = r - k * n;
  const float *src = reinterpret_cast<const float *>(gathered + (size_t)g * stride + h_off);
  (k == 0 ? h0 : k == 1 ? h1 : h2)[(size_t)g * n + j] = src[r];
}
__global__ void sum_splits_kernel(const float *__restrict__ part, int splits, int n, float *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int sidx = 0; sidx < splits; ++sidx) acc += part[(size_t)sidx * n + i];
  out[i] = acc;
}

static int dp_n_h(const rec_engine *e) { return e->cfg.n_heads > 1 ? 3 : 1; }

extern "C" int64_t rec_dp_packed_bytes(const rec_engine *e, int B_local) {
  DevGuard dev_guard(e);
  if (!e || B_local < 1) return -1;
  return rec_packed_batch_bytes(e, B_local) + (int64_t)dp_n_h(e) * B_local * e->D * (int64_t)sizeof(float);
}
extern "C" int64_t rec_dp_grad_floats(const rec_engine *e) {
  DevGuard dev_guard(e);
  if (!e) return -1;
  const int E = e->cfg.embedding_dim, H = e->cfg.hidden_dim;
  return (int64_t)e->dirs * 2 * (3 * H) * ((E > H ? E : H) + 1);
}

extern "C" int rec_dp_forward(rec_engine *e, const rec_batch *local, int main_net, void *packed_out) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  const int n_q = e->cfg.n_heads - 1;
  if (main_net < 0 || main_net >= e->cfg.n_nets) REC_FAIL(e, REC_EINVAL, "main_net out of range");
  int rc = check_net(e, main_net, true);
  if (rc) return rc;
  if ((rc = check_batch(e, local, n_q > 0))) return rc;
  if (!packed_out) REC_FAIL(e, REC_EINVAL, "rec_dp_forward: null argument");
  for (int i = 0; i < 3; ++i) side_join(e, i);
  const int B = local->B, boot = 1 - main_net;
  rec_batch own = e->own;
  own.B = B;
  if (!local->r) { own.r = nullptr; own.s_next = nullptr; own.true_next_len = nullptr; own.is_end = nullptr; }
  copy_batch_kernel<<<cdiv(B * (2 * e->cfg.state_size + 5), 256), 256, 0, e->stream>>>(*local, own, e->cfg.state_size);
  REC_LAUNCH_CHECK(e);
  e->dp_local = own;
  if (n_q > 0) {
    if ((rc = check_net(e, boot, false))) return rc;
    const int nets[3] = {main_net, main_net, boot};
    const int64_t *ss[3] = {own.s, own.s_next, own.s_next};
    const int64_t *ll[3] = {own.true_len, own.true_next_len, own.true_len};
    float *hh[3] = {e->h_state[0], e->h_state[1], e->h_state[2]};
    const bool sv[3] = {true, false, false};
    if ((rc = launch_gru_forward_multi(e, 3, nets, ss, ll, hh, sv, B))) return rc;
  } else {
    if ((rc = launch_gru_forward(e, main_net, own.s, own.true_len, B, e->h_state[0], true))) return rc;
  }
  if ((rc = launch_pack_batch(e, &own, (uint8_t *)packed_out))) return rc;
  const int n = B * e->D, n_h = dp_n_h(e);
  pack_h_kernel<<<cdiv(n_h * n, 256), 256, 0, e->stream>>>(e->h_state[0], e->h_state[1], e->h_state[2], n_h, n,
                                                           (float *)((uint8_t *)packed_out + rec_packed_batch_bytes(e, B)));
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

extern "C" int rec_dp_unpack(rec_engine *e, const void *gathered, int n_ranks, int B_local, const rec_batch *out) {
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  if (!gathered || n_ranks < 1 || B_local < 1 || !out || !out->s || !out->s_next || !out->a || !out->true_len ||
      !out->true_next_len || !out->r || !out->is_end)
    REC_FAIL(e, REC_EINVAL, "rec_dp_unpack: bad argument");
  if (n_ranks * B_local > e->cfg.max_batch) REC_FAIL(e, REC_EINVAL, "rec_dp_unpack: global batch exceeds max_batch");
  const size_t stride = (size_t)rec_dp_packed_bytes(e, B_local);
  int rc = launch_unpack_batch(e, (const uint8_t *)gathered, n_ranks, B_local, stride, out);
  if (rc) return rc;
  const int n = B_local * e->D, n_h = dp_n_h(e);
  unpack_h_kernel<<<cdiv(n_ranks * n_h * n, 256), 256, 0, e->stream>>>((const uint8_t *)gathered, stride,
                                                                      (size_t)rec_packed_batch_bytes(e, B_local), n_ranks, n_h,
                                                                      n, e->h_state[0], e->h_state[1], e->h_state[2]);
  REC_LAUNCH_CHECK(e);
  return REC_OK;
}

extern "C" int rec_train_phase_a_heads(rec_engine *e, const rec_batch *global_b, const rec_train_hparams *hp, int main_net,
                                       float *records_out) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  if (e) e->dp_active = true;
  return phase_a_body(e, global_b, hp, main_net, records_out, true);
}

extern "C" int rec_dp_backward(rec_engine *e, const float *dh_reduced, int rank, float *gru_grads_out, float *dx_out) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  if (e->cur_phase != 3 || !e->dp_active) REC_FAIL(e, REC_EINVAL, "rec_dp_backward called out of order");
  if (!dh_reduced || !gru_grads_out || !dx_out || rank < 0) REC_FAIL(e, REC_EINVAL, "rec_dp_backward: bad argument");
  const rec_batch &lb = e->dp_local;
  const int B = lb.B, main_net = e->cur_main;
  if ((rank + 1) * B > e->cur_batch.B) REC_FAIL(e, REC_EINVAL, "rec_dp_backward: rank %d outside the global batch", rank);
  const float *dh_local = dh_reduced + (size_t)rank * B * e->D;
  int rc;
  if (e->cfg.n_heads == 1 && e->cur_hp.dropout_p > 0.f) {  // dL/dh of this rank's rows through their part of the keep mask
    REC_CUDA(e, cudaMemcpyAsync(e->dh, dh_local, sizeof(float) * (size_t)B * e->D, cudaMemcpyDeviceToDevice, e->stream));
    if ((rc = launch_dropout(e, main_net, nullptr, e->dh, B, &e->cur_hp, true, rank * B))) return rc;
    dh_local = e->dh;
  }
  rc = launch_gru_backward(e, main_net, lb.s, lb.true_len, B, dh_local, e->cur_step_size, e->cur_bc2_sqrt, &e->cur_hp, 1 | 2);
  if (rc) return rc;
  const int n = (int)rec_dp_grad_floats(e);
  sum_splits_kernel<<<cdiv(n, 256), 256, 0, e->stream>>>(e->wgrad_part, e->wgrad_used, n, gru_grads_out);
  REC_LAUNCH_CHECK(e);
  REC_CUDA(e, cudaMemcpyAsync(dx_out, e->dx, sizeof(float) * (size_t)B * e->cfg.state_size * e->dirs * e->cfg.embedding_dim,
                              cudaMemcpyDeviceToDevice, e->stream));
  e->cur_phase = 4;
  return REC_OK;
}

extern "C" int rec_dp_apply(rec_engine *e, const float *gru_grads_reduced, const float *dx_gathered) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  if (e->cur_phase != 4) REC_FAIL(e, REC_EINVAL, "rec_dp_apply called out of order");
  if (!gru_grads_reduced || !dx_gathered) REC_FAIL(e, REC_EINVAL, "rec_dp_apply: null argument");
  const rec_batch *gb = &e->cur_batch;
  const int main_net = e->cur_main;
  e->cur_phase = 0;
  // the launchers read the engine's buffers: point them at the reduced / gathered data for this call
  float *const own_part = e->wgrad_part, *const own_dx = e->dx;
  const int own_splits = e->wgrad_splits, own_used = e->wgrad_used;
  e->wgrad_part = const_cast<float *>(gru_grads_reduced); e->wgrad_splits = 1; e->wgrad_used = 1; e->dx = const_cast<float *>(dx_gathered);
  int rc;
  {
    SideScope side(e, 1);  // embedding chain over the GLOBAL positions next to the GRU Adam
    rc = launch_embedding_update(e, main_net, gb->s, gb->true_len, gb->B, e->cur_step_size, e->cur_bc2_sqrt, &e->cur_hp, 7);
  }
  if (!rc) rc = launch_gru_backward(e, main_net, gb->s, gb->true_len, gb->B, nullptr, e->cur_step_size, e->cur_bc2_sqrt, &e->cur_hp, 4);
  e->wgrad_part = own_part; e->wgrad_splits = own_splits; e->wgrad_used = own_used; e->dx = own_dx;
  for (int i = 0; i < 3; ++i) side_join(e, i);
  return rc;
}

extern "C" int rec_train_phase_b(rec_engine *e, const float *gathered, int n_shards, float *q_out) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  if (e->cur_phase != 1) REC_FAIL(e, REC_EINVAL, "rec_train_phase_b called out of order");
  if (!gathered || n_shards < 1 || (!q_out && e->cfg.n_heads > 1)) REC_FAIL(e, REC_EINVAL, "rec_train_phase_b: bad argument");
  const rec_batch *b = &e->cur_batch;
  const int B = b->B, n_q = e->cfg.n_heads - 1, main_net = e->cur_main;
  int rc;
  if ((rc = launch_head_merge(e, gathered, n_shards, B, e->cur_topk, true, n_q > 0, nullptr))) return rc;
  if (n_q > 0) {
    // this shard's contribution to Q(s,a) and Q_boot(s',a*) (zero when the row lives elsewhere)
    if ((rc = launch_row_dots(e, main_net, e->h_state[0], b->a, nullptr, B, 1, n_q, q_out))) return rc;
    if ((rc = launch_row_dots(e, 1 - main_net, e->h_state[2], nullptr, e->astar, B, 1, n_q, q_out + (int64_t)B * 3))) return rc;
  }
  e->cur_phase = 2;
  return REC_OK;
}

extern "C" int rec_train_phase_c(rec_engine *e, const float *boot_q_reduced, float *losses_out, float *dh_out) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  if (e->cur_phase != 2) REC_FAIL(e, REC_EINVAL, "rec_train_phase_c called out of order");
  if (!losses_out || !dh_out) REC_FAIL(e, REC_EINVAL, "rec_train_phase_c: null argument");
  const rec_batch *b = &e->cur_batch;
  const rec_train_hparams *hp = &e->cur_hp;
  const int B = b->B, n_q = e->cfg.n_heads - 1, main_net = e->cur_main;
  int rc;
  if (n_q > 0) {
    if (!boot_q_reduced) REC_FAIL(e, REC_EINVAL, "rec_train_phase_c: reduced Q buffer is null");
    REC_CUDA(e, cudaMemcpyAsync(e->q_sa, boot_q_reduced, sizeof(float) * (size_t)B * 3, cudaMemcpyDeviceToDevice, e->stream));
    REC_CUDA(e, cudaMemcpyAsync(e->q_boot, boot_q_reduced + (int64_t)B * 3, sizeof(float) * (size_t)B * 3,
                                cudaMemcpyDeviceToDevice, e->stream));
    const float alpha_eff = (n_q == 3) ? hp->alpha : 1.f;
    if ((rc = launch_td(e, b, hp, n_q, alpha_eff, extra(e).q_loss_rows))) return rc;
  }
  if ((rc = launch_loss_reduce(e, B, n_q > 0 ? extra(e).q_loss_rows : nullptr, e->loss_buf))) return rc;
  REC_CUDA(e, cudaMemcpyAsync(losses_out, e->loss_buf, (n_q > 0 ? 2 : 1) * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
  adam_scalars(e, main_net, hp, &e->cur_step_size, &e->cur_bc2_sqrt);
  if ((rc = upload_adam_scalars(e, main_net, hp))) return rc;
  if ((rc = launch_head_backward_adam(e, main_net, e->h_state[0], b, B, e->cur_step_size, e->cur_bc2_sqrt, hp, 1.f / (float)B))) return rc;
  REC_CUDA(e, cudaMemcpyAsync(dh_out, e->dh, sizeof(float) * (size_t)B * e->D, cudaMemcpyDeviceToDevice, e->stream));
  e->cur_phase = 3;
  return REC_OK;
}

extern "C" int rec_train_phase_d(rec_engine *e, const float *dh_reduced) {
  if (e) e->param_epoch++;  // parameters may change: derived operand images are stale (rec_eval_hold_params)
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  if (e->cur_phase != 3) REC_FAIL(e, REC_EINVAL, "rec_train_phase_d called out of order");
  if (!dh_reduced) REC_FAIL(e, REC_EINVAL, "rec_train_phase_d: null argument");
  const rec_batch *b = &e->cur_batch;
  const int main_net = e->cur_main;
  e->cur_phase = 0;
  int rc;
  if (e->cfg.n_heads == 1 && e->cur_hp.dropout_p > 0.f) {  // dL/dh through the keep mask of phase A
    if (dh_reduced != e->dh)
      REC_CUDA(e, cudaMemcpyAsync(e->dh, dh_reduced, sizeof(float) * (size_t)b->B * e->D, cudaMemcpyDeviceToDevice, e->stream));
    if ((rc = launch_dropout(e, main_net, nullptr, e->dh, b->B, &e->cur_hp, true))) return rc;
    dh_reduced = e->dh;
  }
  rc = trunk_backward(e, main_net, b->s, b->true_len, b->B, dh_reduced, e->cur_step_size, e->cur_bc2_sqrt, &e->cur_hp, false);
  for (int i = 0; i < 3; ++i) side_join(e, i);  // the Q-head sweep forked in phase C rejoins here
  return rc;
}

// Sharded evaluation: per-shard record (max, sumexp, target logit, top-k candidates) per row ...
extern "C" int rec_eval_shard_candidates(rec_engine *e, int net_id, const rec_batch *b, int head_idx, int kmax,
                                         float *records_out) {
  DevGuard dev_guard(e);
  int rc = check_net(e, net_id, false);
  if (rc) return rc;
  if ((rc = check_batch(e, b, false))) return rc;
  if (!records_out || head_idx < 0 || head_idx >= e->cfg.n_heads || kmax < 1 || kmax > e->cfg.max_topk)
    REC_FAIL(e, REC_EINVAL, "rec_eval_shard_candidates: bad argument");
  REC_CUDA(e, cudaMemsetAsync(records_out, 0, sizeof(float) * (size_t)b->B * e->part_stride, e->stream));
  if ((rc = launch_gru_forward(e, net_id, b->s, b->true_len, b->B, e->h_state[0], false))) return rc;
  float w[3] = {1.f, 0.f, 0.f};
  return shard_head_pass(e, net_id, e->h_state[0], b, head_idx, kmax, 0, w, true, records_out);
}

// ... and the merge of the all-gathered records of all shards + metric accumulation (replicated).
extern "C" int rec_eval_merge(rec_engine *e, const rec_batch *b, const rec_eval_opts *o, const float *gathered,
                              int n_shards, const rec_eval_accum *acc, int32_t *topk_ids, float *topk_scores) {
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  int rc = check_batch(e, b, false);
  if (rc) return rc;
  if (!o || !acc || !gathered || n_shards < 1) REC_FAIL(e, REC_EINVAL, "rec_eval_merge: bad argument");
  const int kmax = eval_kmax(o);
  if (kmax > e->cfg.max_topk) REC_FAIL(e, REC_EINVAL, "rec_eval_merge: k=%d exceeds max_topk=%d", kmax, e->cfg.max_topk);
  if ((rc = launch_head_merge(e, gathered, n_shards, b->B, kmax, true, false, nullptr))) return rc;
  return launch_eval_metrics(e, b, o, kmax, acc, extra(e).rowm, topk_ids, topk_scores);
}

// ---- packed batches for the input all-gather of sharded runs ----------------------------------------------
extern "C" int rec_build_replay_rows(rec_engine *e, const int64_t *session_offsets, int64_t n_sessions, const int64_t *items,
                                     const float *rewards, int64_t n_events, int64_t pad_id, int pad_pos_end,
                                     const rec_batch *out) {
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  if (!session_offsets || !items || !out || n_sessions < 1 || n_events < 1)
    REC_FAIL(e, REC_EINVAL, "rec_build_replay_rows: null argument or empty log");
  if (!out->s || !out->s_next || !out->a || !out->true_len || !out->true_next_len || !out->is_end)
    REC_FAIL(e, REC_EINVAL, "rec_build_replay_rows: every output column except r is mandatory");
  return launch_build_rows(e, session_offsets, n_sessions, items, rewards, n_events, pad_id, pad_pos_end ? 1 : 0, out);
}

extern "C" int rec_gather_batch(rec_engine *e, const rec_batch *columns, int64_t n_rows, const int64_t *idx, int B,
                                const rec_batch *out) {
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  if (!columns || !idx || !out || n_rows < 1 || B < 1) REC_FAIL(e, REC_EINVAL, "rec_gather_batch: null argument or empty buffer");
  if (!columns->s || !columns->a || !columns->true_len || !out->s || !out->a || !out->true_len)
    REC_FAIL(e, REC_EINVAL, "rec_gather_batch: s, a and true_len are mandatory on both sides");
  return launch_gather_batch(e, columns, n_rows, idx, B, out);
}

extern "C" int64_t rec_packed_batch_bytes(const rec_engine *e, int B) {
  DevGuard dev_guard(e);
  if (!e || B < 1) return -1;
  int64_t n = (int64_t)B * (2 * e->cfg.state_size + 3) * 8 + (int64_t)B * 5;
  return (n + 15) / 16 * 16;
}
extern "C" int rec_pack_batch(rec_engine *e, const rec_batch *b, void *packed_out) {
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  if (!b || !packed_out || !b->s || !b->a || !b->true_len || b->B < 1) REC_FAIL(e, REC_EINVAL, "rec_pack_batch: bad argument");
  return launch_pack_batch(e, b, (uint8_t *)packed_out);
}
extern "C" int rec_unpack_batch(rec_engine *e, const void *gathered, int n_ranks, int B_local, const rec_batch *out) {
  DevGuard dev_guard(e);
  if (!e) return REC_EINVAL;
  if (!gathered || n_ranks < 1 || B_local < 1 || !out || !out->s || !out->s_next || !out->a || !out->true_len ||
      !out->true_next_len || !out->r || !out->is_end)
    REC_FAIL(e, REC_EINVAL, "rec_unpack_batch: bad argument");
  return launch_unpack_batch(e, (const uint8_t *)gathered, n_ranks, B_local, (size_t)rec_packed_batch_bytes(e, B_local), out);
}
