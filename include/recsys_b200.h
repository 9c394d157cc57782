/*
 * recsys_b200.h -- C ABI of the B200-native GRU4Rec / BidirGRU4Rec / SQN / SMORL hot path.
 *
 * Drop-in boundary for the PyTorch path of adam-walsh-data/IKEA-Recommender-System
 * (`recommenders/`).  The reference is pure Python: the "FFI" a maintainer would add is a
 * ctypes binding; each entry point below cites the reference interface
 * it replaces (paths relative to the reference root, `recommenders/...`).
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is DEVICE memory owned by the caller
 *     (the engine owns only its opaque handle + private workspace);
 *   - all calls are asynchronous on the `stream` given at creation (a cudaStream_t passed as
 *     void*; NULL = legacy default stream) unless documented otherwise;
 *   - return 0 on success, negative REC_E* on failure, never throws across the ABI; the message
 *     is available from rec_last_error();
 *   - not thread-safe per handle; one handle per (GPU, process);
 *   - no CPU fallback: rec_create fails unless the device is compute capability 10.x.
 *
 * Tensor layouts are PyTorch's: embedding [N+1,E]; weight_ih [3H,E], weight_hh [3H,H], biases [3H]
 * with gate order (r,z,n); head weight [V,D], bias [V]; all fp32, contiguous.  Index tensors are
 * int64 exactly as the reference's DataLoader yields them.
 */
#ifndef RECSYS_B200_H
#define RECSYS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REC_ABI_VERSION 1
#define REC_MAX_HEADS 5   /* supervised + up to 3 Q heads (SMORL); SARM: 5 Q heads, head 0 doubles as the supervised head */
#define REC_MAX_NETS 2    /* double-Q twins */
#define REC_MAX_TOPK 32   /* largest k of any top-k consumer */
#define REC_MAX_KLIST 8   /* entries in topk_hr_ndcg / topk_cov lists */

#define REC_OK 0
#define REC_EINVAL (-22)
#define REC_ENOMEM (-12)
#define REC_ENODEV (-19)
#define REC_ECUDA (-5)

typedef struct rec_engine rec_engine; /* opaque */

/* Static shape of one model family.
 * Replaces the constructor arguments of GRU4Rec / BidirGRU4Rec / SQN_Network / SMORL_GRU_Net
 * (models/GRU4Rec/model.py:6-60, models/BidirGRU4Rec/model.py:7-68, models/SQN/sqn_gru.py:10-85,
 *  models/SMORL/smorl_gru.py:14-102). */
typedef struct rec_config {
  int32_t item_num;       /* N; the embedding table has N+1 rows */
  int32_t action_dim;     /* V, global vocabulary of every head */
  int32_t embedding_dim;  /* E (multiple of 4) */
  int32_t hidden_dim;     /* H (multiple of 4) */
  int32_t state_size;     /* L */
  int32_t bidirectional;  /* 0 | 1 ; D = H * (1 + bidirectional) */
  int32_t n_heads;        /* 1 supervised only | 2 SQN (sup,q) | 4 SMORL (sup,q_acc,q_div,q_nov) | 5 SARM (q_heads[0..4]) */
  int32_t n_nets;         /* 1 | 2 (double-Q twins) */
  int32_t use_packed_seq; /* 1: final state after exactly len tokens (pack_padded_sequence) */
  int32_t frozen_pad_row; /* -1: every embedding row trainable; else row that never gets gradient */
  int32_t max_batch;      /* largest B of any later call */
  int32_t vocab_lo;       /* head rows [vocab_lo, vocab_hi) live on this engine (vocabulary shard); */
  int32_t vocab_hi;       /*   0 / V for an unsharded engine                                          */
  int32_t max_topk;       /* largest k any call will ask for (<= REC_MAX_TOPK) */
} rec_config;

/* Parameter + Adam-state pointers of one net (nn.Module.state_dict() of the reference modules;
 * Adam state of torch.optim.Adam(lr) as in sqn_gru.py:173-181).  Head pointers address the LOCAL
 * shard rows [vocab_lo,vocab_hi).  m/v may be NULL for a net that is only evaluated. */
typedef struct rec_net_params {
  float *emb, *emb_m, *emb_v;                            /* [N+1,E] */
  float *w_ih[2], *w_ih_m[2], *w_ih_v[2];                /* [3H,E]  per direction */
  float *w_hh[2], *w_hh_m[2], *w_hh_v[2];                /* [3H,H] */
  float *b_ih[2], *b_ih_m[2], *b_ih_v[2];                /* [3H] */
  float *b_hh[2], *b_hh_m[2], *b_hh_v[2];                /* [3H] */
  float *head_w[REC_MAX_HEADS], *head_w_m[REC_MAX_HEADS], *head_w_v[REC_MAX_HEADS]; /* [Vloc,D] */
  float *head_b[REC_MAX_HEADS], *head_b_m[REC_MAX_HEADS], *head_b_v[REC_MAX_HEADS]; /* [Vloc] */
} rec_net_params;

/* One replay-buffer batch, the tuple of ikea/data_utils/replay_buffer.py:65-74 (device copies). */
typedef struct rec_batch {
  int32_t B;
  const int64_t *s;              /* [B,L] */
  const int64_t *a;              /* [B] */
  const float *r;                /* [B]   offline accuracy reward (NULL for supervised) */
  const int64_t *s_next;         /* [B,L] (NULL for supervised) */
  const int64_t *true_len;       /* [B] */
  const int64_t *true_next_len;  /* [B] (NULL for supervised) */
  const uint8_t *is_end;         /* [B] bool (NULL for supervised) */
} rec_batch;

/* Hyper-parameters of a train step: Adam defaults of torch.optim.Adam (betas .9/.999, eps 1e-8);
 * gamma / q_weights / alpha of sqn_gru.py:238-245 and smorl_gru.py:317-325. */
typedef struct rec_train_hparams {
  float lr, beta1, beta2, eps;
  float gamma;
  float alpha;        /* SMORL: loss = sup + alpha*q ; SQN: loss = q + sup (alpha ignored, =1) */
  float q_weights[3]; /* SMORL scalarisation weights w (SQN: {1,0,0}) */
  /* SMORL online rewards (evaluate/diversity.py:15-73, evaluate/novelty.py:12-47) */
  const float *div_emb;        /* frozen E_div [N+1, div_dim] */
  int32_t div_dim;
  int32_t topk_div, topk_nov;
  float nov_reward;
  const uint8_t *unpopular;    /* bitmap-as-bytes [V]: 1 if action id is in the unpopular set */
  const int64_t *out_to_in;    /* optional LUT [V]: output-token id -> input-token id (NULL = identity) */
  int32_t pad_pos_end;         /* 1: "end" padding (last action = s[len-1]); 0: "beg" (s[L-1]) */
  /* BidirGRU4Rec: nn.Dropout(p) on concat(h_fwd, h_bwd) in train mode (BidirGRU4Rec/model.py:60,93), supervised
   * step only.  The keep mask is drawn on the device from (dropout_seed, Adam step, element) unless
   * dropout_mask (device uint8 [B, D], 1 = keep) injects one -- torch's Philox stream cannot be reproduced, so
   * parity tests feed the same mask to the oracle.  dropout_p = 0 disables. */
  float dropout_p;
  uint64_t dropout_seed;
  const uint8_t *dropout_mask;
} rec_train_hparams;

/* Options of one evaluation sweep: the keyword arguments of evaluate()/update_train_metrics()
 * (evaluate/eval_protocol.py:123-139, 266-284). */
typedef struct rec_eval_opts {
  int32_t head_idx;                  /* which head's logits are scored (0 = supervised) */
  int32_t n_k, ks[REC_MAX_KLIST];    /* topk_hr_ndcg */
  int32_t n_cov, cov_ks[REC_MAX_KLIST]; /* topk_to_consider_cov */
  int32_t topk_div, topk_nov;
  float nov_reward;
  const float *div_emb;
  int32_t div_dim;
  const uint8_t *unpopular;          /* [V] */
  const int64_t *out_to_in;          /* optional [V] */
  int32_t pad_pos_end;
} rec_eval_opts;

/* Device-side accumulators of an evaluation sweep (caller allocates, zero-initialises). */
typedef struct rec_eval_accum {
  double *hits;        /* [REC_MAX_KLIST] */
  double *ndcg;        /* [REC_MAX_KLIST] */
  double *reps;        /* [REC_MAX_KLIST] */
  double *div_sum;     /* [1] */
  double *nov_sum;     /* [1] */
  double *loss_sum;    /* [1] sum of per-batch mean CE (eval_protocol.py:182,250) */
  uint32_t *cov_bits;  /* [REC_MAX_KLIST][ceil(V/32)] coverage bitmaps per k */
} rec_eval_accum;

/* ---- lifetime ------------------------------------------------------------------------------ */
int rec_abi_version(void);
/* Creates an engine on the current CUDA device. */
int rec_create(const rec_config *cfg, void *stream, rec_engine **out);
void rec_destroy(rec_engine *e);
const char *rec_last_error(const rec_engine *e); /* valid until the next call on e; e may be NULL */

/* Binds caller-owned parameter storage to net `net_id` (replaces holding an nn.Module). Must be
 * called again after the caller rewrites GRU weights out-of-band (load_state_dict). */
int rec_bind_params(rec_engine *e, int net_id, const rec_net_params *p);
/* Number of optimizer steps taken for a net (torch.optim.Adam state['step']). */
int rec_set_adam_step(rec_engine *e, int net_id, int64_t step);
int64_t rec_get_adam_step(const rec_engine *e, int net_id);

/* ---- forward (replaces model(s, lengths): GRU4Rec/model.py:62-82, sqn_gru.py:87-112) -------- */
/* Final layer-0 GRU state h_out[B, D].  `lengths` is a DEVICE copy of the CPU lengths tensor. */
int rec_forward_state(rec_engine *e, int net_id, const int64_t *s, const int64_t *lengths, int B,
                      float *h_out);
/* Full logits of one head for the local vocabulary shard: logits[B, Vloc] (row stride ld). Only
 * for API compatibility with `model(s, lengths)`; the hot path never materialises logits. */
int rec_head_logits(rec_engine *e, int net_id, int head, const float *h, int B, float *logits,
                    int64_t ld);

/* ---- training (replaces *_trainer.train_step) ----------------------------------------------- */
/* GRU4Rec_trainer/BidirGRU4Rec_trainer.train_step (GRU4Rec/model.py:129-155): writes the batch-mean
 * CE loss to loss_out[0] (device). */
int rec_train_step_supervised(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp,
                              float *loss_out);
/* SQN_trainer.train_step (sqn_gru.py:183-254) / SMORL_trainer.train_step (smorl_gru.py:233-334):
 * `main_net` in {0,1} is the twin picked by the caller's python RNG (sqn_gru.py:207-216).
 * losses_out[0] = sup_loss, losses_out[1] = q_loss (device). */
int rec_train_step_q(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp, int main_net,
                     float *losses_out);
/* SARM_trainer.train_step (models/SARM/sarm.py:117-149) on a single-net engine with n_heads = 5 (MultiObjectiveQNetwork,
 * sarm.py:5-76): for every head i  q_loss_i = mean_b (r + gamma * max_a Q_i(s', a) - Q_i(s, a))^2  with the next-state
 * values detached (no is_end masking: the reference masks only a value it never uses), sup_loss = CE(Q_0(s, .), a),
 * loss = sup_loss + mean_i q_loss_i, one Adam step on the net.  hp->gamma carries the trainer's 0.99 (sarm.py:112).
 * losses_out[0] = sup_loss, losses_out[1] = mean_i q_loss_i (device).  The python RNG draw of the reference
 * (random.randint, :123) only selects tensors that do not reach the loss; the host mirror consumes it. */
int rec_train_step_sarm(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp, float *losses_out);
/* The same two steps called the way the reference's trainers are: every pointer of `host_b` is HOST memory
 * (the CPU tensors a DataLoader yields; hp->div_emb / hp->unpopular stay device pointers) and the losses come
 * back as host floats (the reference returns `loss.item()`).  Synchronous: the batch is packed into a pinned
 * mirror, one H2D copy, the step and the D2H copy of the losses run as one CUDA graph, then the call waits. */
int rec_train_step_supervised_host(rec_engine *e, const rec_batch *host_b, const rec_train_hparams *hp,
                                   float *loss_host);
int rec_train_step_q_host(rec_engine *e, const rec_batch *host_b, const rec_train_hparams *hp, int main_net,
                          float *losses_host);

/* Phase-split variant of the same step for vocabulary-sharded multi-GPU runs: every rank holds the full
 * (all-gathered) batch, replicas of embedding + GRU and rows [vocab_lo,vocab_hi) of every head; the caller
 * runs the three (tiny) collectives between the phases.  All buffers are caller-owned device memory.
 *   phase A: GRU forwards + per-shard head statistics -> records_out[B, rec_record_floats()]: per row
 *            (max, sum-exp, target logit, argmax value/id of sum_h w_h Q_h(s',.), top-k candidates)
 *            -> caller ALL-GATHERS the records of all shards into gathered[n_shards, B, record]
 *   phase B: merges the gathered records (log-sum-exp combine; (score desc, id asc) order), then this
 *            shard's contribution to Q(s,a) and Q_boot(s',a*) -> q_out[2, B, 3] (zeros for rows owned by
 *            another shard) -> caller ALL-REDUCES (sum) q_out
 *   phase C: rewards, TD target, losses (losses_out[0]=sup, [1]=q), head backward + fused Adam on the
 *            shard -> dh_out[B, D] partial -> caller ALL-REDUCES (sum) dh_out
 *   phase D: GRU BPTT, GRU / embedding gradients + Adam (replicated, bit-identical on every rank). */
int rec_record_floats(const rec_engine *e);
int rec_train_phase_a(rec_engine *e, const rec_batch *b, const rec_train_hparams *hp, int main_net,
                      float *records_out);
int rec_train_phase_b(rec_engine *e, const float *gathered, int n_shards, float *q_out);
int rec_train_phase_c(rec_engine *e, const float *q_reduced, float *losses_out, float *dh_out);
int rec_train_phase_d(rec_engine *e, const float *dh_reduced);

/* Data-parallel trunk for the sharded step (the heads stay sharded by vocabulary and see the global batch; the
 * replicated embedding + GRU run on each rank's OWN sessions only, so their cost does not grow with the number
 * of GPUs).  Sequence per step, collectives by the caller:
 *   rec_dp_forward(local batch)            -> packed_out[rec_dp_packed_bytes(B_local)]          all-gather
 *   rec_dp_unpack(gathered, G, B_local)    -> global batch fields (+ final states inside the engine)
 *   rec_train_phase_a_heads / _b / _c      (as rec_train_phase_a/b/c, without the GRU forward)  all-reduce dh
 *   rec_dp_backward(dh_reduced, rank)      -> gru_grads_out[rec_dp_grad_floats()]               all-reduce
 *                                          -> dx_out[B_local * L * dirs * E]                    all-gather
 *   rec_dp_apply(gru_grads_reduced, dx_gathered)   identical Adam update of the replicas on every rank */
int64_t rec_dp_packed_bytes(const rec_engine *e, int B_local);
int64_t rec_dp_grad_floats(const rec_engine *e);
int rec_dp_forward(rec_engine *e, const rec_batch *local_b, int main_net, void *packed_out);
int rec_dp_unpack(rec_engine *e, const void *gathered, int n_ranks, int B_local, const rec_batch *global_out);
int rec_train_phase_a_heads(rec_engine *e, const rec_batch *global_b, const rec_train_hparams *hp, int main_net,
                            float *records_out);
int rec_dp_backward(rec_engine *e, const float *dh_reduced, int rank, float *gru_grads_out, float *dx_out);
int rec_dp_apply(rec_engine *e, const float *gru_grads_reduced, const float *dx_gathered);

/* Row-sharded embedding table ("the same row slice of the embedding table (+m,v) so the dense-Adam sweep
 * scales 1/G"; collectives C1 / C5).  The reference has one nn.Embedding per net (models/SQN/sqn_gru.py:50-62) updated by
 * dense Adam (sqn_gru.py:173-181).  rec_set_embedding_shard makes this rank the OWNER of rows [row_lo, row_hi) of every
 * net's table: its Adam sweep touches only those rows (+ their m, v); (0, item_num + 1) restores the unsharded sweep.
 * The rank keeps a full-size copy; rows it does not own go stale and are refreshed from their owners before they are read:
 *   rec_emb_rows_gather(net, ids[n])  -> rows_out[n, E]: the current row where this rank owns ids[i], zeros elsewhere
 *   all-reduce(sum) of rows_out by the caller (exactly one owner contributes a non-zero row: the sum is bit-exact)
 *   rec_emb_rows_scatter(net, ids[n], rows[n, E]): writes the rows this rank does NOT own into its copy
 * Gradient rows need no extra exchange: every rank already sees the dx rows of the global batch (phase D /
 * rec_dp_apply) and applies the ones that fall into its row range. */
int rec_set_embedding_shard(rec_engine *e, int64_t row_lo, int64_t row_hi);
int rec_emb_rows_gather(rec_engine *e, int net_id, const int64_t *ids, int64_t n, float *rows_out);
int rec_emb_rows_scatter(rec_engine *e, int net_id, const int64_t *ids, int64_t n, const float *rows);

/* Input plumbing of sharded runs: one rank's batch packed into ONE byte buffer (so a single all-gather moves
 * every field), and the inverse for the gathered [n_ranks][rec_packed_batch_bytes] buffer -> field arrays of
 * n_ranks*B_local rows (caller-owned; out->B is ignored). */
int64_t rec_packed_batch_bytes(const rec_engine *e, int B);
/* Device-resident replay buffer (replaces ReplayBuffer.__getitem__ + the DataLoader's default collate,
 * ikea/data_utils/replay_buffer.py:65-74 and ikea/training/trainSQN.py's `for batch in train_loader`): `columns` holds
 * the base pointers of the WHOLE buffer in HBM (n_rows rows; columns->B is ignored), `idx` [B] int64 device row
 * numbers (one slice of the epoch's permutation).  Row idx[b] of every non-NULL column -> row b of `out`
 * (device arrays of B rows, caller-owned).  Asynchronous on the engine's stream; indices are clamped to the buffer. */
int rec_gather_batch(rec_engine *e, const rec_batch *columns, int64_t n_rows, const int64_t *idx, int B,
                     const rec_batch *out);
/* Replay-buffer construction on the device (replaces the pandas groupby-apply of
 * recommenders/data_utils/preprocessing.py:199-268 with its helpers get_state :5-29 and get_next_state :143-170, shared
 * by ikea/data_utils/preprocessing.py:385-489): an event log sorted by session, given as CSR `session_offsets`
 * [n_sessions+1] (int64, device) over `items` [n_events] (+ optional per-event `rewards`), becomes one replay row per
 * event in `out` (n_events rows, device, caller-owned): state / next_state padded to cfg.state_size with `pad_id` at the
 * end (pad_pos_end != 0) or the beginning, action, true_len = clip(n_before, 1, L), true_next_len = min(n_before+1, L),
 * is_end = last event of its session, r = rewards (0 when NULL; out->r may be NULL). */
int rec_build_replay_rows(rec_engine *e, const int64_t *session_offsets, int64_t n_sessions, const int64_t *items,
                          const float *rewards, int64_t n_events, int64_t pad_id, int pad_pos_end, const rec_batch *out);
int rec_pack_batch(rec_engine *e, const rec_batch *b, void *packed_out);
int rec_unpack_batch(rec_engine *e, const void *gathered, int n_ranks, int B_local, const rec_batch *out);

/* ---- evaluation (replaces evaluate()/update_train_metrics(), eval_protocol.py:123-359) ------ */
/* One batch: forward, fused top-k (score desc, id asc), CE, and every metric accumulated on the
 * device.  topk_ids[B,kmax] (int32, global action ids) and topk_scores[B,kmax] may be NULL. */
int rec_eval_batch(rec_engine *e, int net_id, const rec_batch *b, const rec_eval_opts *o,
                   const rec_eval_accum *acc, int32_t *topk_ids, float *topk_scores);
/* An evaluation sweep calls rec_eval_batch / rec_eval_shard_candidates many times with unchanged parameters
 * (evaluate(), eval_protocol.py:154-171: model.eval(), one pass over the validation loader).  on = 1: the caller
 * promises not to modify the bound parameter tensors until on = 0; the engine may then keep operand images derived from
 * them (the packed bf16 hi/lo image of the scored head: 512 MB of traffic per batch at 1 M items) across calls.  Any
 * training entry point of this engine ends the reuse by itself. */
int rec_eval_hold_params(rec_engine *e, int on);
/* Sharded evaluation (vocabulary-sharded heads over several GPUs): each shard produces one record per
 * row -- (max, sum-exp, target logit, top-kmax candidates as score/id lists) -- in the caller's device
 * buffer records_out[B, rec_record_floats()]; the caller all-gathers the records of all shards ... */
int rec_eval_shard_candidates(rec_engine *e, int net_id, const rec_batch *b, int head_idx, int kmax,
                              float *records_out);
/* ... and every rank merges the gathered [n_shards, B, record_floats] records (score desc, id asc;
 * log-sum-exp combine) and accumulates the metrics exactly like rec_eval_batch. */
int rec_eval_merge(rec_engine *e, const rec_batch *b, const rec_eval_opts *o, const float *gathered,
                   int n_shards, const rec_eval_accum *acc, int32_t *topk_ids, float *topk_scores);

/* ---- introspection -------------------------------------------------------------------------- */
/* 1 (default): rec_train_step_* replay a captured CUDA graph of the whole step from its second call on (inputs are
 * first copied into engine-owned buffers, Adam scalars live in device memory); 0: plain stream launches. */
int rec_set_cuda_graphs(rec_engine *e, int on);
/* Switches the stream every later call launches on (the engine is created on the caller's current stream).
 * Needed when the caller captures the phase-split step, NCCL collectives included, into its own CUDA graph:
 * the engine must launch on the capturing stream (sharded.py: ShardedStep). */
int rec_set_stream(rec_engine *e, void *stream);
/* Debug/A-B switch: 0 forces the generic CUDA-core head kernels even when the tcgen05 path applies (D == 64). */
int rec_set_tensor_cores(rec_engine *e, int on);
/* Number of kernels this engine launched since creation (bench.py's gpu_launches). */
int64_t rec_launch_count(const rec_engine *e);
/* CUDA-event time (ms) of the most recent dominant-kernel launch when profiling is enabled. */
int rec_enable_kernel_timing(rec_engine *e, int on);
float rec_last_kernel_ms(rec_engine *e, int which); /* which: 0 supervised-head bwd+Adam, 1 eval head statistics,
                                                        2 embedding Adam sweep, 3 Q-heads Adam sweep,
                                                        4 supervised-head statistics (train), 5 greedy-action pass */

/* Debug / tests: copies the greedy actions a*[B] (int32, global action ids) of the most recent Q step into the caller's
 * device buffer (asynchronous on the engine's stream).  argmax of sqn_gru.py:229 / tensor_operations.py:73-84. */
int rec_debug_copy_astar(rec_engine *e, int32_t *out, int B);
/* Debug: device buffer [240] int64 receiving (tag, clock64) pairs of CTA 0 of the tensor-core backward kernel
 * (REC_TRACE_SEL=1: the supervised statistics kernel, =2: the greedy-action kernel). */
int rec_debug_set_trace(rec_engine *e, long long *dev_buf);
/* Self-test of the tcgen05/TMEM plumbing (bf16x3 split GEMM of one 128-row tile; see csrc/tc_selftest.cu):
 * mode 0: C[128,128] = A[128,64].B[128,64]^T ; mode 1: C[128,64] = P[128,128]^T.Q[128,64] ;
 * mode 2: C[128,64] = P[128,128].R[128,64].  Device pointers, fp32. */
int rec_debug_tc_gemm(int mode, const float *A, const float *B, float *C, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RECSYS_B200_H */
