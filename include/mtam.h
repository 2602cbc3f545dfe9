/* mtam.h -- C-ABI of libmtam_b200.so, the sm_100a implementation of MTAMRecommender's
 * training / scoring hot path.
 *
 * The reference (TensorFlow 1.14, pure Python) has no FFI of its own: the only seam is
 * `tf.Session.run(fetches, feed_dict)` under the Python classes in Embedding/ and Model/.  Each
 * entry point below names the reference call site (file:line under /root/reference) whose device
 * work it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference
 * side.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / C++ types.
 *  - every function returns 0 on success or a negative mtam_status; `mtam_last_error()` gives text.
 *  - all data pointers are DEVICE pointers owned by the caller unless the name says `host`.
 *    The library allocates nothing on the device: parameters, gradients, Adam slots and the
 *    workspace are arenas the caller allocates (sizes from `mtam_plan`) and hands over.
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no host sync inside
 *    unless stated.  A handle is not thread-safe.
 *  - floats are fp32, indices int32, exactly as the reference's placeholders
 *    (Embedding/Behavior_embedding_time_aware_attention.py:21-46).
 */
#ifndef MTAM_H_
#define MTAM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTAM_ABI_VERSION 2

typedef enum {
  MTAM_OK = 0,
  MTAM_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  MTAM_ERR_CUDA = -2,        /* CUDA runtime error (text in mtam_last_error) */
  MTAM_ERR_WORKSPACE = -3,   /* workspace too small */
  MTAM_ERR_UNSUPPORTED = -4  /* model kind / feature not built */
} mtam_status;

/* train_process.py:164-218 experiment_type dispatch */
typedef enum {
  MTAM_KIND_MTAM = 0,       /* 'MTAM'                              Model/MTAMRec_model.py:61-92 */
  MTAM_KIND_PISTREC = 1,    /* Time_Aware_self_Attention_model     Model/PISTRec_model.py:38-74 */
  MTAM_KIND_SASREC = 2,     /* 'SASrec'                            Model/attention_baseline_models.py:33-46 */
  MTAM_KIND_TA_SASREC = 3,  /* 'Time_Aware_Self_Attention_Model'   Model/attention_baseline_models.py:47-65 */
  MTAM_KIND_TISASREC = 4,   /* 'Ti_Self_Attention_Model'           Model/attention_baseline_models.py:66-84 */
  MTAM_KIND_BPRMF = 5,      /* 'bpr'                               Model/BPRMF.py:10-59 */
  MTAM_KIND_MTAM_VIA_T_GRU = 6, /* 'MTAM_via_T_GRU'                Model/MTAMRec_model.py:167-204: the hops read the
                               T-GRU's OUTPUT SEQUENCE as their memory, and the query is layer-normed first */
  MTAM_KIND_MTAM_NO_TIME_AWARE_RNN = 7,  /* 'MTAM_no_time_aware_rnn'  MTAMRec_model.py:93-125: MTAM with a plain GRU
                               (tf GRUCell, Model/Modules/gru.py:60-67) as the intent encoder */
  MTAM_KIND_MTAM_VIA_RNN = 8    /* 'MTAM_via_rnn'                  MTAMRec_model.py:206-233: MTAM_via_T_GRU with the plain GRU */
} mtam_kind;

/* How the dense contractions are computed. */
typedef enum {
  MTAM_GEMM_FP32 = 0,       /* fp32 FFMA everywhere (tightest parity) */
  MTAM_GEMM_TF32X3 = 1,     /* tcgen05 kind::tf32, 3-term error-compensated split (fp32-class accuracy) */
  MTAM_GEMM_TF32 = 2        /* the separately-toleranced fast mode (SURVEY 8c "bf16-operand GEMM mode"): as TF32X3, except
                               that the three products of the softmax cross-entropy (logits, dPred, dTable: over 90 % of
                               the FLOPs of a large-catalogue step) are issued as ONE kind::tf32 MMA each, operands
                               rounded to nearest tf32 (10 mantissa bits), fp32 accumulation.  Tolerance (tests):
                               loss rel 1e-2, gradients norm-wise rel 2e-2, recall@50 >= 0.99.  The top-k filter keeps
                               the 3-term product (its result is exact either way). */
} mtam_gemm_mode;

/* base_model.init_optimizer (base_model.py:71-80).  Adam is what every preset uses; any name the reference does not
 * know falls through to plain gradient descent.  RMSProp / Adadelta are not built: TF applies them lazily to the rows an
 * IndexedSlices gradient names, which needs per-row slot bookkeeping no preset exercises. */
typedef enum {
  MTAM_OPT_ADAM = 0,        /* tf.train.AdamOptimizer(lr): beta1 0.9, beta2 0.999, eps 1e-8, non-lazy sparse apply */
  MTAM_OPT_SGD = 1          /* tf.train.GradientDescentOptimizer(lr) */
} mtam_optimizer;

typedef struct {
  int32_t abi_version;      /* MTAM_ABI_VERSION */
  int32_t kind;             /* mtam_kind */
  int32_t max_batch;        /* largest B any call will pass */
  int32_t L;                /* FLAGS.length_of_user_history  (config/model_parameter.py:49) */
  int32_t D;                /* FLAGS.num_units   (:14)  64 or 128 */
  int32_t H;                /* FLAGS.num_heads   (:13) */
  int32_t N;                /* FLAGS.num_blocks  (:12) */
  int32_t user_rows;        /* user_count + 3      (Behavior_...py:64) */
  int32_t item_rows;        /* item_count + 3      (:71) */
  int32_t category_rows;    /* category_count + 3  (:78) */
  int32_t position_rows;    /* max_length_seq + 3  (:86) */
  float reg;                /* FLAGS.regulation_rate  (base_model.py:309) */
  float clip;               /* FLAGS.max_gradient_norm (base_model.py:294) */
  float beta1, beta2, eps;  /* tf.train.AdamOptimizer defaults (base_model.py:76) */
  int32_t gemm_mode;        /* mtam_gemm_mode */
  int32_t optimizer;        /* mtam_optimizer */
  float dropout;            /* FLAGS.dropout: attention dropout of SASREC / TISASREC (multihead_attention.py:179,
                               time_aware_attention.py:198); the other kinds have none.  Active in every forward pass,
                               evaluation included, as in the reference (is_training=True is hard-coded) */
  uint32_t dropout_seed;    /* keys the keep mask together with the forward-call counter (mtam_set_dropout_state) */
  int32_t reserved[4];
} mtam_config;

typedef struct {
  uint64_t param_floats;    /* length (floats) of each of the four arenas: params, grads, adam m, adam v */
  uint64_t workspace_bytes; /* scratch for max_batch */
} mtam_sizes;

/* The 11 feed arrays of init_placeholders (Behavior_...py:21-46), as device pointers. */
typedef struct {
  int32_t B;
  const int32_t* user_id;            /* [B]   */
  const int32_t* item_list;          /* [B,L] */
  const int32_t* category_list;      /* [B,L] */
  const int32_t* position_list;      /* [B,L] */
  const float* time_list;            /* [B,L] */
  const float* timelast_list;        /* [B,L] */
  const float* timenow_list;         /* [B,L]  (sliced but unused by the live graph) */
  const int32_t* target_item_id;     /* [B]   */
  const int32_t* target_item_category; /* [B] (unused by the live graph) */
  const float* target_item_time;     /* [B]   */
  const int32_t* seq_length;         /* [B]   */
} mtam_batch;

/* A data set of examples in columnar form (host memory): replaces the reference's Python list of 9-tuples
 * (user, items, cats, times, timelast, timenow, positions, [target id, cat, time], length) that
 * Prepare/prepare_data_base.py:79-92 reads back with eval() and make_feed_dic_new pads list by list
 * (Behavior_...py:146-192).  Record r owns steps [offsets[r], offsets[r+1]) of the six per-step columns. */
typedef struct {
  int64_t n_records;
  const int64_t* offsets;            /* [n_records + 1] */
  const int32_t* user_id;            /* [n_records] */
  const int32_t* target_item_id;     /* [n_records] */
  const int32_t* target_item_category;
  const float* target_item_time;
  const int32_t* seq_length;         /* [n_records]: the tuple's `length` field, fed verbatim */
  const int32_t* item;               /* [offsets[n_records]] */
  const int32_t* category;
  const int32_t* position;
  const float* time;
  const float* timelast;
  const float* timenow;
} mtam_record_store;

#define MTAM_NAME_MAX 160
#define MTAM_PARAM_DEAD 1    /* created by the reference but never reached by tf.gradients */
#define MTAM_PARAM_TABLE 2   /* embedding table (sparse + dense gradient pieces) */

typedef struct {
  char name[MTAM_NAME_MAX];  /* reference-compatible variable name */
  int32_t rows, cols;        /* logical shape (1-D variables: rows = 1) */
  int32_t ndim;              /* 1 or 2 */
  int32_t ld;                /* row stride in floats inside the arena (>= cols) */
  uint64_t offset;           /* first element, in floats, from the arena base */
  int32_t flags;
} mtam_param_info;

/* Scalars a train / forward call writes to `scalars_out` (device float[8]). */
enum {
  MTAM_S_LOSS = 0,         /* self.loss          base_model.py:322 */
  MTAM_S_LOSS_ORIGIN = 1,  /* mean(loss_origin)  :323 */
  MTAM_S_L2_NORM = 2,      /* l2_norm            :302-307 */
  MTAM_S_GLOBAL_NORM = 3,  /* tf.clip_by_global_norm's norm  :294 */
  MTAM_S_CLIP_SCALE = 4,
  MTAM_S_COUNT = 8
};

typedef struct mtam_model* mtam_handle;

/* ---- stand-alone bandwidth kernels (callable without a model) ----------------------------- */

/* out[i,:] = table[idx[i],:]   -- tf.nn.embedding_lookup, Behavior_...py:68,75,82,90.
 * D must be a multiple of 4; table/out 16-byte aligned.  Bit-exact copy. */
int mtam_gather(const float* table, int32_t table_rows, int32_t D, const int32_t* idx, int64_t n,
                float* out, void* stream);

/* Bytes of scratch mtam_scatter_add needs for n indices into a table of `table_rows` rows. */
size_t mtam_scatter_add_workspace(int64_t n, int32_t table_rows, int32_t D);

/* Aggregation of the rows of equal indices -- the gradient of tf.nn.embedding_lookup
 * (tf.gradients -> IndexedSlices -> unsorted_segment_sum, base_model.py:292,296):
 *   accumulate != 0:  dst[idx[i],:] += rows[i,:]  for i in [0,n)
 *   accumulate == 0:  dst[r,:] = sum of rows[i,:] over {i : idx[i] == r} for every r that occurs in idx; other rows of
 *                     dst are not touched (on a zeroed dst both forms give the same result; this one never reads dst).
 * Deterministic: stable radix sort of (idx, i), then a segmented reduction that adds the rows of one index in
 * ascending i and writes each distinct row of dst exactly once.  `ld_rows` (>= D) is the row stride of `rows` in
 * floats.  If `unique_idx`/`n_unique` are non-null they receive the distinct indices (ascending) and their count
 * (device). */
int mtam_scatter_add(float* dst, int32_t table_rows, int32_t D, const int32_t* idx, const float* rows,
                     int32_t ld_rows, int64_t n, int32_t accumulate, void* workspace, size_t workspace_bytes,
                     int32_t* unique_idx, int32_t* n_unique, void* stream);

/* The two halves of mtam_scatter_add, separately callable: the index sort depends only on the batch's ids, so a
 * training step runs it off the critical path (and once per id list, however many tables share it); the
 * segmented reduction is the HBM-bound part.
 *  mtam_sort_indices: stable radix sort of (keys[i], i), keys in [0, key_bound), one sweep over the data per 8-bit
 *    digit (decoupled look-back).  *keys_sorted and *perm (perm[j] = original position of the j-th smallest key)
 *    point INTO the workspace on return.
 *  mtam_scatter_add_sorted: dst[keys_sorted[j],:] (+)= rows[perm[j],:] (perm NULL: rows already in sorted order),
 *    rows of one key added in ascending j, `accumulate` as above.  dst rows are D floats apart. */
size_t mtam_sort_workspace(int64_t n, int32_t key_bound);
int mtam_sort_indices(const int32_t* keys, int64_t n, int32_t key_bound, void* workspace, size_t workspace_bytes,
                      const int32_t** keys_sorted, const int32_t** perm, void* stream);
size_t mtam_scatter_add_sorted_workspace(int64_t n, int32_t D);
int mtam_scatter_add_sorted(float* dst, int32_t D, const int32_t* keys_sorted, const int32_t* perm, const float* rows,
                            int32_t ld_rows, int64_t n, int32_t accumulate, void* workspace, size_t workspace_bytes,
                            void* stream);

/* ---- model --------------------------------------------------------------------------------- */

int mtam_plan(const mtam_config* cfg, mtam_sizes* out);

/* Arenas are param_floats long each; the library keeps the pointers, the caller keeps ownership. */
int mtam_create(const mtam_config* cfg, float* params, float* grads, float* adam_m, float* adam_v,
                void* workspace, size_t workspace_bytes, mtam_handle* out);
int mtam_destroy(mtam_handle h);
const char* mtam_last_error(mtam_handle h);   /* h may be NULL: last error of a create/plan call */

/* Variable inventory: tf.trainable_variables() of the model (base_model.py:328) */
int mtam_param_count(mtam_handle h);
int mtam_param_info_get(mtam_handle h, int32_t i, mtam_param_info* out);

/* Adam time step t (beta1_power/beta2_power state of tf.train.AdamOptimizer); 0 after create. */
int mtam_get_adam_step(mtam_handle h, int64_t* t);
int mtam_set_adam_step(mtam_handle h, int64_t t);

/* Forward only: the graph up to self.loss / self.predict_behavior_emb (base_model.py:300-323).
 * pred_out [B,D] and loss_origin_out [B] may be NULL. */
int mtam_forward(mtam_handle h, const mtam_batch* batch, float* scalars_out, float* loss_origin_out,
                 float* pred_out, void* stream);

/* One `sess.run([loss, merged, train_op])` (base_model.py:150-167): forward, tf.gradients,
 * clip_by_global_norm, Adam.apply_gradients.  lr is the float64 placeholder of base_model.py:25. */
int mtam_train_step(mtam_handle h, const mtam_batch* batch, double lr, float* scalars_out, void* stream);

/* The same step in three phases so a data-parallel driver can put its collectives between them.
 *  1. forward_backward: local loss pieces and gradients, with the mean taken over `global_batch`
 *     sequences.  Dense pieces land in the grads arena, the sparse embedding pieces (IndexedSlices
 *     values) stay in the workspace.  scalars_out gets the local partial sums.
 *     `norm_sq_sparse` (device float[1]) += sum of squares of the local un-deduplicated sparse values.
 *  2. finish_grads: adds the sum of squares of the (already all-reduced) dense pieces to the value
 *     in norm_sq (device float[1], pre-loaded with the all-reduced sparse part); if scatter_local != 0
 *     it then scatter-adds the local sparse pieces into the grads arena (deterministic sort +
 *     segmented reduce).  A data-parallel driver passes 0, all-gathers the pieces described by
 *     mtam_sparse_pieces and calls mtam_scatter_add on the grads arena itself.
 *  3. apply: clip by sqrt(*norm_sq) and run Adam over the arenas; zeroes the grads arena rows it
 *     dirtied. */
int mtam_forward_backward(mtam_handle h, const mtam_batch* batch, int32_t global_batch,
                          float* scalars_out, float* norm_sq_sparse, void* stream);
int mtam_finish_grads(mtam_handle h, float* norm_sq, int32_t scatter_local, void* stream);
int mtam_apply(mtam_handle h, double lr, const float* norm_sq, float* scalars_out, void* stream);
/* mtam_apply in three parts, for a data-parallel driver that wants to update the regions of the arena whose gradients
 * are complete while a collective for the others is still in flight: begin = clip scale + step preparation (lr_t),
 * range = the optimizer update of arena floats [float_begin, float_end) (multiples of 4; each float exactly once per
 * step), end = restores the gradient-arena invariants.  mtam_apply == begin, range(0, param_floats), end. */
int mtam_apply_begin(mtam_handle h, double lr, const float* norm_sq, float* scalars_out, void* stream);
int mtam_apply_range(mtam_handle h, uint64_t float_begin, uint64_t float_end, void* stream);
int mtam_apply_end(mtam_handle h, void* stream);

/* Data-parallel helpers.
 *  mtam_set_item_grad_event: `cuda_event` (a cudaEvent_t owned by the caller, or NULL to clear) is recorded on the
 *    step's stream as soon as the dense item-table gradient is complete (right after the softmax backward, the
 *    first thing the backward pass does), so a driver can start all-reducing it on another stream while the rest
 *    of the backward pass runs.
 *  mtam_scatter_sparse_into: scatter-adds the LOCAL sparse pieces of the last forward_backward (deterministic sort
 *    + segmented reduce; the sorts ran beside the forward pass) into caller-provided [rows, D] tables instead of the
 *    grads arena; a NULL destination skips that table.  Used when the tables are small enough that all-reducing a
 *    densified piece is cheaper than all-gathering the rows (multi-GPU). */
int mtam_set_item_grad_event(mtam_handle h, void* cuda_event);
int mtam_scatter_sparse_into(mtam_handle h, float* item_dst, float* category_dst, float* position_dst, float* user_dst,
                             void* stream);

/* BPR-MF only: fixes the negative item id that `tf.random_uniform([1], 0, item_count)` (BPRMF.py:43) would
 * draw, so a run can be reproduced; item_id < 0 restores the per-step draw from the handle's own generator. */
/* ---- row-sharded item table: the scalable training variant -----------------------------------------------
 * The reference keeps the whole item table on one device and both looks it up (Behavior_...py:68-75) and scores against
 * it (base_model.py:316).  With the table row-sharded over the ranks, a handle is created with item_rows = the rows of
 * THIS rank's shard (they live in the arenas under embedding_layer/item) and the caller owns the exchanges:
 *   item rows of the batch  <- all-to-all lookup from the owning ranks              (parallel.ShardedCatalogue.lookup)
 *   mtam_forward_rows       : forward pass up to pred with those rows in place of the table lookup (item_rows [B*L, D],
 *                             row t = table[item_list[t]]; the batch's item_list itself is not read); no softmax.
 *                             scalars_out[MTAM_S_L2_NORM] = the L2 term over this rank's rows.
 *   softmax cross-entropy   <- mtam_softmax_ce_forward / _backward on the shard + the combination of the per-shard
 *                             log-sum-exps; the shard's dense table gradient is purely local
 *   mtam_backward_rows      : backward pass from dpred (gradient of the loss mean over global_batch).  Dense parameter
 *                             gradients go to the grads arena (its item-table region is not touched); the IndexedSlices
 *                             values stay in the workspace (mtam_sparse_pieces: the first D columns of item_cat_rows are
 *                             d loss / d item_rows, to be sent back to the owners and scatter-added there); their squared
 *                             norm is added to *norm_sq_sparse.
 *   then mtam_scatter_sparse_into (category / position / user), mtam_sumsq for the dense pieces, mtam_apply. */
int mtam_forward_rows(mtam_handle h, const mtam_batch* batch, const float* item_rows, float* pred_out, float* scalars_out,
                      void* stream);
int mtam_backward_rows(mtam_handle h, const mtam_batch* batch, const float* item_rows, const float* dpred,
                       int32_t global_batch, float* norm_sq_sparse, void* stream);
/* out_accumulate[0] += sum x[i]^2 (a dense piece of tf.clip_by_global_norm's norm, base_model.py:294); x 16-byte aligned */
size_t mtam_sumsq_workspace(int64_t n);
int mtam_sumsq(const float* x, int64_t n, float* out_accumulate, void* workspace, size_t workspace_bytes, void* stream);

/* Attention dropout state: the keep mask of forward call number `counter` is a pure function of (seed, counter, block,
 * element) -- csrc/selfattn.cu sa_keep(); tests restate it to give the oracle the same mask.  The next forward pass
 * (train or eval) uses exactly `counter`, later ones count up from it. */
int mtam_set_dropout_state(mtam_handle h, uint32_t seed, uint32_t counter);
int mtam_set_bpr_negative(mtam_handle h, int32_t item_id);

/* Per-step host half of mtam_apply: advances the Adam step (beta powers) and enqueues a one-thread
 * kernel on `stream` that stores lr_t = lr*sqrt(1-b2^t)/(1-b1^t) in device memory (the value travels as
 * a kernel argument).  mtam_apply calls it itself unless it was already called for this step.  To replay
 * a CUDA graph of mtam_train_step: call mtam_prepare_step eagerly right before the capture (so the
 * captured step does not contain it) and before every replay. */
int mtam_prepare_step(mtam_handle h, double lr, void* stream);

/* Where the sparse gradient pieces (IndexedSlices values) of the last forward_backward live. */
typedef struct {
  int32_t B, L, D;
  const float* item_cat_rows;   /* [B*L, 2D]: item values in cols [0,D), category values in [D,2D) */
  const float* position_rows;   /* [B*L, D] */
  const float* user_rows;       /* [B, D]  (NULL-equivalent when has_user == 0) */
  int32_t has_user;
  uint64_t dense_begin;         /* grads arena: floats [dense_begin, param_floats) hold dense pieces */
  uint64_t user_offset, item_offset, category_offset, position_offset;  /* table regions in the arenas */
} mtam_sparse_view;
int mtam_sparse_pieces(mtam_handle h, mtam_sparse_view* out);

/* Per-phase device timing of a step (CUDA events on the step's stream). */
enum {
  MTAM_PH_EMBED_FWD = 0, MTAM_PH_GRU_X_GEMM, MTAM_PH_GRU_FWD, MTAM_PH_KV_GEMM, MTAM_PH_HOP_FWD, MTAM_PH_CE_FWD,
  MTAM_PH_CE_BWD, MTAM_PH_HOP_BWD, MTAM_PH_HOP_PARAM_GRADS, MTAM_PH_GRU_BWD, MTAM_PH_GRU_PARAM_GRADS,
  MTAM_PH_EMBED_BWD, MTAM_PH_DENSE_NORM, MTAM_PH_SCATTER, MTAM_PH_ADAM, MTAM_PHASE_COUNT
};
int mtam_profile_enable(mtam_handle h, int32_t on);
int mtam_profile_read(mtam_handle h, float* ms_out, int32_t n);   /* ms_out[MTAM_PHASE_COUNT]; host sync */

/* Stand-alone GEMM (diagnostic / tests): C[M,N] = act(op(A) op(B) + bias), row-major with leading dimensions;
 * transA: A stored [K,M]; transB: B stored [N,K]; mode = mtam_gemm_mode.  Replaces tf.matmul / tf.layers.dense
 * call sites (time_aware_attention.py:249-253, Behavior_...py:95-103). */
int mtam_gemm(int32_t mode, int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K, const float* A,
              int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc, const float* bias, int32_t relu,
              int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream);
size_t mtam_gemm_workspace(int32_t M, int32_t N, int32_t K);

/* make_feed_dic_new (Behavior_...py:146-192) for B records of a columnar store: records index[0..B) (or
 * first..first+B when index is NULL) are right-padded with 0 to L steps and written into the 11 HOST arrays `out`
 * points to (typically the pinned staging buffers of the step's single H2D copy).  Host code only; a record
 * longer than L is an error, as np.pad's negative width is in the reference. */
int mtam_pack_records(const mtam_record_store* rs, const int64_t* index, int64_t first, int32_t B, int32_t L,
                      const mtam_batch* out);

/* Kernels launched by the library since it was loaded (diagnostic). */
long long mtam_launch_count(void);

/* Full-catalogue scoring + top-k: tf.matmul(pred, item_table^T) + tf.nn.top_k (base_model.py:194-202).
 * Sorted descending, ties -> lower index.  idx_out [B,k] int32, score_out [B,k] (may be NULL).
 * `row_begin`/`row_end` restrict scoring to item rows [row_begin,row_end) (a table shard); indices
 * written are global row numbers.
 * gemm_mode (mtam_gemm_mode; mtam_eval_topk uses the handle's): MTAM_GEMM_FP32 computes every score with an fp32 FMA
 * chain; MTAM_GEMM_TF32X3 (num_units 32 / 64) runs the [B,V] product on tcgen05 as a filter (maximum logit per bucket
 * of 16 / 64 items), then rescores the items of the k+14 best buckets per row in fp32 (fixed summation order, so
 * sharded == unsharded bit for bit) -- exact top-k of those fp32 scores unless more than 14 bucket maxima tie with
 * the k-th within the 3xTF32 rounding error. */
int mtam_eval_topk(mtam_handle h, const mtam_batch* batch, int32_t k, int32_t* idx_out, float* score_out,
                   void* stream);
int mtam_score_topk(int32_t gemm_mode, const float* pred, int32_t B, int32_t D, const float* item_table,
                    int32_t row_begin, int32_t row_end, int32_t k, int32_t* idx_out, float* score_out,
                    void* workspace, size_t workspace_bytes, void* stream);
size_t mtam_score_topk_workspace(int32_t B, int32_t rows, int32_t k);
/* The filter pass of mtam_score_topk (MTAM_GEMM_TF32X3) on its own: bmax[b*ld + j] = max over the items of bucket j
 * (bucket_size = 16 or 64 consecutive rows of item_table[0, rows)) of <pred_b, row> computed on tcgen05 (3xTF32).
 * ld >= 8*ceil(ceil(rows/128)*(128/bucket_size) / 8) floats.  Entries past ceil(rows/bucket_size) are unspecified.
 * Exposed so that the filter can be checked against exhaustive scoring (base_model.py:194-195). */
int mtam_score_bucket_max(const float* pred, int32_t B, int32_t D, const float* item_table, int32_t rows,
                          int32_t bucket_size, float* bmax, int32_t ld, void* stream);
/* Tuning: row ranges longer than `rows` use buckets of 64 items in the filter pass, shorter ones buckets of 16
 * (default 2^21; 0 restores it).  Process-wide; the result of mtam_score_topk does not depend on it. */
int mtam_set_topk_bucket_crossover(int32_t rows);

/* Softmax cross-entropy of base_model.output (base_model.py:316-321) against item rows [table, table + rows*D) -- the
 * whole catalogue or one shard of a row-sharded table (sharded log-sum-exp).  `target` is relative to the
 * range; a target outside [0, rows) has no logit here.
 *   forward : lse_out[b] = log sum_v exp(<pred_b, row_v>) over the range; target_logit_out[b] = <pred_b, row_target>
 *             where the target lies in the range, else 0.  Over shards: lse = logsumexp_r(lse_r), target logit = sum_r.
 *   backward: with the GLOBAL lse, G = (exp(logit - lse) - onehot) * inv_batch; dtable[rows,D] = G^T pred (every row of
 *             the range, written once), dpred[B,D] = G table (this range's share: sum over shards).
 * The [B,rows] logits are never materialised; gemm_mode as in mtam_score_topk. */
size_t mtam_softmax_ce_workspace(int32_t B, int32_t D, int32_t rows);
int mtam_softmax_ce_forward(int32_t gemm_mode, const float* pred, int32_t B, int32_t D, const float* table, int32_t rows,
                            const int32_t* target, float* lse_out, float* target_logit_out, void* workspace,
                            size_t workspace_bytes, void* stream);
int mtam_softmax_ce_backward(int32_t gemm_mode, const float* pred, int32_t B, int32_t D, const float* table, int32_t rows,
                             const int32_t* target, const float* lse, float inv_batch, float* dtable, float* dpred,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Merge per-shard top-k lists: in_idx/in_score are [n_lists,B,k] (each list sorted, shard order =
 * index order) -> out [B,k].  Used after the all-gather of the row-sharded eval. */
int mtam_merge_topk(const int32_t* in_idx, const float* in_score, int32_t n_lists, int32_t B, int32_t k,
                    int32_t* out_idx, float* out_score, void* stream);

/* HR@k / NDCG@k of calculate_topK (base_model.py:215-242) for k in {1,5,10,30,50} from a top-50
 * list: out10 (device float[10]) = hr1,ndcg1,hr5,ndcg5,...,hr50,ndcg50. */
int mtam_hr_ndcg(const int32_t* topk_idx, int32_t B, int32_t k, const int32_t* target, float* out10,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MTAM_H_ */
