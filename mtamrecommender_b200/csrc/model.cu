// Model handle, arena layout and step orchestration behind the C-ABI (include/mtam.h).
// Replaces the device work of `sess.run([loss, merged, train_op], feed)` in
// base_model.train (Model/base_model.py:150-167) for the MTAM graph (Model/MTAMRec_model.py:61-92).
#include <math.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "model_kernels.h"
#include "selfattn.h"
#include "../../include/mtam.h"

namespace mtam {

static const char* kGruLive[8] = {"_time_kernel_w1", "_time_kernel_b1", "_time_history_w1", "_time_w1",
                                  "_time_b1", "_time_kernel_w2", "_time_w12", "_time_b12"};
static const char* kGruDead[6] = {"_time_history_b1", "_time_kernel_b2", "_time_history_w2",
                                  "_time_history_b2", "_time_w2", "_time_b2"};
static const char* kGateLive[5] = {"_time_input_w1", "_time_input_b1", "time_output_w1", "time_output_w2",
                                   "time_output_b"};

// Arena layout (floats).  Tables whose gradient is sparse-only come first so that the dense-piece
// norm can skip them; every block starts on a 4-float boundary.
struct Layout {
  size_t user = 0, cat = 0, pos = 0, dense_begin = 0, item = 0, Wemb = 0, Wgru = 0, bgru = 0, gruvec = 0, Wq = 0,
         bq = 0, Wkv = 0, bkv = 0, Wt = 0, gate = 0, gate_dead = 0, lnb = 0, lng = 0, lnfb = 0, lnfg = 0, lnsb = 0, lnsg = 0, item_b = 0,
         total = 0;
  SaLayout sa;
  std::vector<ParamDesc> params;
};

struct Workspace {
  // activations
  float *E2, *R, *X, *GX, *Hs, *RUCT, *RH, *KV;
  float *Qin, *Qr, *Qt, *AA, *PA, *ZZ, *DK, *GT, *XH, *RSTD, *XHF, *RSTDF, *pred;
  float *q0raw, *dq0raw, *XHS, *RSTDS;   // MTAM_via_T_GRU: the short-term intent before its layer norm, and the norm's state
  float *tlogit, *lse, *loss_origin;
  // gradients of activations
  float *dpred, *dX, *dKV, *DOUT, *DQP, *DQT, *BKV, *GB, *dq0, *WqT, *WtT, *dGX, *vec_partial, *dR, *dE2, *dEp, *dEu;
  // partial-sum buffers and device scalars
  float *l2_partial, *sq_partial, *ce_partial, *norm_partial, *dev_scalars;  // dev_scalars: [16]
  float *bU, *bIP, *bdot, *bbpos, *browsum, *bcolsum, *bloss, *bdU, *bdIP, *bdINp, *bdIN, *bdbneg, *bl2, *bsq;  // BPRMF
  int32_t* bneg_idx;
  int32_t* iota;     // 0..max_batch*L-1: token numbers as row ids of a caller-supplied [B*L, D] item-row array
  unsigned* user_marks;   // one bit per user row: named by the current batch (zero between steps)
  size_t user_marks_words;
  void *ce_ws, *gemm_ws, *colsum_ws, *gemm_ws2, *colsum_ws2, *scatter_ws, *sa_ws, *topk_ws, *sort_ws[4], *seg_ws;
  void* seg_ws_side[3];   // the category / position / user reductions run beside the item table's, each with its own scratch
  size_t ce_ws_bytes, gemm_ws_bytes, colsum_ws_bytes, scatter_ws_bytes, sa_ws_bytes, topk_ws_bytes, sort_ws_bytes[4],
      seg_ws_bytes;
  size_t total_bytes;
};

}  // namespace mtam

using namespace mtam;

struct mtam_model {
  mtam_config cfg;
  Layout lay;
  Workspace ws;
  float *params, *grads, *m, *v;
  int64_t adam_t = 0;
  float b1_pow = 1.f, b2_pow = 1.f;  // fp32 running products, like TF's beta1_power / beta2_power variables
  mtam_batch last_batch;
  int last_B = 0;
  bool grads_pending = false;
  float* lr_host = nullptr;          // pinned: lr_t of the current step (read by a captured H2D copy)
  bool step_prepared = false;        // mtam_prepare_step already advanced the Adam state for this step
  bool prof = false;
  cudaStream_t side = nullptr;       // index sorts run here, concurrently with forward/backward
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t side2 = nullptr;      // tensor-core CE: the dense item-table gradient runs here, beside the backward chain
  cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr;
  cudaStream_t side3 = nullptr;      // parameter gradients (weight-gradient GEMMs, column sums): nothing downstream of the
  cudaEvent_t ev_pg[3] = {}, ev_join3 = nullptr;   // backward chain reads them, so they run beside it
  cudaStream_t side4 = nullptr;      // the user table's gradient-free Adam rows, from the start of the step (optim.cu)
  cudaEvent_t ev_fork4 = nullptr, ev_join4 = nullptr;
  cudaEvent_t ev_sc_fork = nullptr, ev_sc[3] = {};   // the four segmented reductions of a step, side by side
  cudaEvent_t ev_ce_done = nullptr;  // caller-owned: recorded once the dense item-table gradient is complete
  const float* item_rows_ext = nullptr;   // row-sharded item table: the batch's item rows, fetched by the caller
  bool rows_mode = false;                 // ... and the softmax is the caller's too (mtam_forward_rows / mtam_backward_rows)
  bool sort_pending = false;
  const int32_t *sk[4] = {}, *sp[4] = {};   // sorted keys / permutations: item, category, position, user
  uint32_t drop_seed = 0, drop_calls = 0;   // attention dropout: seed and forward-call counter (device copy: dev_scalars[10])
  bool drop_prepared = false;        // the device copy already holds the counter of the next forward pass
  bool drop_explicit = false;        // ... and it was placed there by mtam_set_dropout_state: the next step must use it
  int bpr_neg = -1;                  // injected negative item id (mtam_set_bpr_negative); -1: draw one per step
  int bpr_neg_used = 0;
  uint64_t rng = 1234;
  cudaEvent_t ev[MTAM_PHASE_COUNT + 1] = {};
  bool ev_valid[MTAM_PHASE_COUNT + 1] = {};
  std::string err;
};

namespace mtam {

static inline bool is_mtam_family(int kind) {
  return kind == MTAM_KIND_MTAM || (kind >= MTAM_KIND_MTAM_VIA_T_GRU && kind <= MTAM_KIND_MTAM_VIA_RNN);
}
// the hops' memory is the intent encoder's output sequence (and the query is layer-normed)
static inline bool memory_is_rnn(int kind) { return kind == MTAM_KIND_MTAM_VIA_T_GRU || kind == MTAM_KIND_MTAM_VIA_RNN; }
// the intent encoder is tf's GRUCell, without the time gate
static inline bool plain_gru(int kind) { return kind == MTAM_KIND_MTAM_NO_TIME_AWARE_RNN || kind == MTAM_KIND_MTAM_VIA_RNN; }

// phase boundary marker (cudaEventRecord on the step's stream when profiling is on)
static inline void phase(mtam_model* h, int id, cudaStream_t st) {
  if (!h->prof) return;
  if (!h->ev[id]) cudaEventCreate(&h->ev[id]);
  cudaEventRecord(h->ev[id], st);
  h->ev_valid[id] = true;
}

static size_t a4(size_t x) { return (x + 3) / 4 * 4; }

static void add_param(Layout& l, const std::string& name, int rows, int cols, int ndim, int ld, size_t off,
                      int flags = 0) {
  l.params.push_back(ParamDesc{name, rows, cols, ndim, ld, off, flags});
}

static int build_layout(const mtam_config& c, Layout& l) {
  const int D = c.D, L = c.L, N = c.N;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o = a4(o + n); return r; };
  l.user = take((size_t)c.user_rows * D);
  l.cat = take((size_t)c.category_rows * D);
  l.pos = take((size_t)c.position_rows * D);
  l.dense_begin = o;
  l.item = take((size_t)c.item_rows * D);
  add_param(l, "embedding_layer/user", c.user_rows, D, 2, D, l.user, MTAM_PARAM_TABLE);
  add_param(l, "embedding_layer/item", c.item_rows, D, 2, D, l.item, MTAM_PARAM_TABLE);
  add_param(l, "embedding_layer/category", c.category_rows, D, 2, D, l.cat, MTAM_PARAM_TABLE);
  add_param(l, "embedding_layer/position", c.position_rows, D, 2, D, l.pos, MTAM_PARAM_TABLE);
  l.Wemb = take((size_t)2 * D * D);
  add_param(l, "position_embedding/dense4emb/kernel", 2 * D, D, 2, D, l.Wemb);
  if (c.kind == MTAM_KIND_BPRMF) {
    l.item_b = take((size_t)c.item_rows);
    add_param(l, "embedding_layer/item_b", c.item_rows, 1, 2, 1, l.item_b, MTAM_PARAM_TABLE);
    l.total = o;
    return 0;
  }
  if (is_mtam_family(c.kind)) {
    const std::string g = "ShortTermIntentEncoder/";
    l.Wgru = take((size_t)2 * D * 3 * D);
    add_param(l, g + "gates/kernel", 2 * D, 2 * D, 2, 3 * D, l.Wgru);
    add_param(l, g + "candidate/kernel", 2 * D, D, 2, 3 * D, l.Wgru + 2 * D);
    l.bgru = take((size_t)3 * D);
    add_param(l, g + "gates/bias", 1, 2 * D, 1, 2 * D, l.bgru);
    add_param(l, g + "candidate/bias", 1, D, 1, D, l.bgru + 2 * D);
    l.gruvec = take((size_t)14 * D);
    if (!plain_gru(c.kind)) {   // the time gate's vectors: the region stays (zeros) for the plain cell, unnamed
      for (int i = 0; i < 8; ++i) add_param(l, g + kGruLive[i], 1, D, 1, D, l.gruvec + (size_t)i * D);
      for (int i = 0; i < 6; ++i)
        add_param(l, g + kGruDead[i], 1, D, 1, D, l.gruvec + (size_t)(8 + i) * D, MTAM_PARAM_DEAD);
    }
    l.Wq = take((size_t)N * D * D);
    l.bq = take((size_t)N * D);
    l.Wkv = take((size_t)D * 2 * N * D);
    l.bkv = take((size_t)2 * N * D);
    l.Wt = take((size_t)N * D * D);
    l.gate = take((size_t)N * 5 * L);
    l.gate_dead = take((size_t)N * L);
    l.lnb = take((size_t)N * D);
    l.lng = take((size_t)N * D);
    for (int i = 0; i < N; ++i) {
      const std::string b = "NextItemDecoder/decoder/num_blocks_" + std::to_string(i) + "/";
      add_param(l, b + "dense/kernel", D, D, 2, D, l.Wq + (size_t)i * D * D);
      add_param(l, b + "dense/bias", 1, D, 1, D, l.bq + (size_t)i * D);
      add_param(l, b + "dense_1/kernel", D, D, 2, 2 * N * D, l.Wkv + (size_t)i * 2 * D);
      add_param(l, b + "dense_1/bias", 1, D, 1, D, l.bkv + (size_t)i * 2 * D);
      add_param(l, b + "dense_2/kernel", D, D, 2, 2 * N * D, l.Wkv + (size_t)i * 2 * D + D);
      add_param(l, b + "dense_2/bias", 1, D, 1, D, l.bkv + (size_t)i * 2 * D + D);
      add_param(l, b + "vanilla_attention/_time_input_w", D, D, 2, D, l.Wt + (size_t)i * D * D);
      for (int k = 0; k < 5; ++k)
        add_param(l, b + "vanilla_attention/" + kGateLive[k], 1, L, 2, L, l.gate + ((size_t)i * 5 + k) * L);
      add_param(l, b + "vanilla_attention/time_output_w3", 1, L, 2, L, l.gate_dead + (size_t)i * L, MTAM_PARAM_DEAD);
      add_param(l, b + "vanilla_attention/ln/beta", 1, D, 1, D, l.lnb + (size_t)i * D);
      add_param(l, b + "vanilla_attention/ln/gamma", 1, D, 1, D, l.lng + (size_t)i * D);
    }
    l.lnfb = take(D);
    l.lnfg = take(D);
    add_param(l, "NextItemDecoder/LayerNorm/beta", 1, D, 1, D, l.lnfb);
    add_param(l, "NextItemDecoder/LayerNorm/gamma", 1, D, 1, D, l.lnfg);
    if (memory_is_rnn(c.kind)) {   // layer_norm of the short-term intent (MTAMRec_model.py:187, :220)
      l.lnsb = take(D);
      l.lnsg = take(D);
      add_param(l, g + "LayerNorm/beta", 1, D, 1, D, l.lnsb);
      add_param(l, g + "LayerNorm/gamma", 1, D, 1, D, l.lnsg);
    }
    l.total = o;
    return 0;
  }
  // self-attention family (PISTRec / SASRec / TA-SASRec / TiSASRec)
  MTAM_TRY(sa_build_layout(c, o, l.params, l.sa));
  l.lnfb = l.sa.lnfb;
  l.lnfg = l.sa.lnfg;
  l.total = a4(o);
  return 0;
}

static int plan_workspace(const mtam_config& c, void* base, size_t cap, Workspace& w) {
  const int64_t B = c.max_batch, L = c.L, D = c.D, N = c.N, H = c.H;
  const int64_t T = B * L;
  Bump b(base, cap);
  memset(&w, 0, sizeof(w));
  w.E2 = b.take<float>(T * 2 * D);
  w.R = b.take<float>(T * D);
  w.X = b.take<float>(T * D);
  w.pred = b.take<float>(B * D);
  w.XHF = b.take<float>(B * D);
  w.RSTDF = b.take<float>(B);
  w.tlogit = b.take<float>(B);
  w.lse = b.take<float>(B);
  w.loss_origin = b.take<float>(B);
  w.dpred = b.take<float>(B * D);
  w.dX = b.take<float>(T * D);
  w.dR = b.take<float>(T * D);
  w.dE2 = b.take<float>(T * 2 * D);
  w.dEp = b.take<float>(T * D);
  w.dEu = b.take<float>(B * D);
  w.l2_partial = b.take<float>(kEmbedMaxBlocks);
  w.sq_partial = b.take<float>(kEmbedMaxBlocks);
  w.ce_partial = b.take<float>(cdiv(B, 32) + 1);
  w.norm_partial = b.take<float>(kNumSMs * 4 + 4);
  w.dev_scalars = b.take<float>(16);
  w.iota = b.take<int32_t>(T);
  size_t gemm_ws = 0, colsum_ws = 0;
  auto G = [&](int M, int Nn, int K) { gemm_ws = std::max(gemm_ws, gemm_any_workspace_bytes(M, Nn, K)); };
  auto CS = [&](int M, int Nn) { colsum_ws = std::max(colsum_ws, colsum_workspace_bytes(M, Nn)); };
  G(2 * D, D, T);  // dWemb
  if (is_mtam_family(c.kind)) {
    w.q0raw = b.take<float>(B * D);
    w.dq0raw = b.take<float>(B * D);
    w.XHS = b.take<float>(B * D);
    w.RSTDS = b.take<float>(B);
    w.GX = b.take<float>(T * 3 * D);
    w.Hs = b.take<float>((T + 1) * D);
    w.RUCT = b.take<float>(T * 4 * D);
    w.RH = b.take<float>(T * D);
    w.KV = b.take<float>(T * 2 * N * D);
    w.Qin = b.take<float>((N + 1) * B * D);
    w.Qr = b.take<float>(N * B * D);
    w.Qt = b.take<float>(N * B * D);
    w.AA = b.take<float>(N * B * H * L);
    w.PA = b.take<float>(N * B * H * L);
    w.ZZ = b.take<float>(N * B * L);
    w.DK = b.take<float>(N * B * L);
    w.GT = b.take<float>(N * B * L);
    w.XH = b.take<float>(N * B * D);
    w.RSTD = b.take<float>(N * B);
    w.dKV = b.take<float>(T * 2 * N * D);
    w.DOUT = b.take<float>(B * N * D);
    w.DQP = b.take<float>(B * N * D);
    w.BKV = b.take<float>(B * 2 * N * D);
    w.DQT = b.take<float>(B * N * D);
    w.GB = b.take<float>(B * 5 * N * L);
    w.dq0 = b.take<float>(B * D);
    w.WqT = b.take<float>(N * D * D);
    w.WtT = b.take<float>(N * D * D);
    w.dGX = b.take<float>(T * 3 * D);
    w.vec_partial = b.take<float>((int64_t)gru_num_blocks(B) * 8 * D);
    G(D, 2 * N * D, T); G(D, 3 * D, T); G(D, 2 * D, T); G(D, D, T); G(D, D, B);
    gemm_ws = std::max(gemm_ws, gemm_atb_batched_workspace_bytes(N, (int)D, (int)D, (int)B));
    CS(T, 2 * N * D); CS(T, 3 * D); CS(B, 5 * N * L); CS(B, N * D); CS(gru_num_blocks(B), 8 * D); CS(B, D);
  } else if (c.kind == MTAM_KIND_BPRMF) {
    w.bU = b.take<float>(B * D); w.bIP = b.take<float>(B * D); w.bdU = b.take<float>(B * D);
    w.bdIP = b.take<float>(B * D); w.bdINp = b.take<float>(B * D); w.bdIN = b.take<float>(D);
    w.bdot = b.take<float>(B); w.bbpos = b.take<float>(B); w.browsum = b.take<float>(B); w.bcolsum = b.take<float>(B);
    w.bloss = b.take<float>(B); w.bdbneg = b.take<float>(4); w.bl2 = b.take<float>(cdiv(B, 4) + 2);
    w.bsq = b.take<float>(cdiv(B, 4) + 2); w.bneg_idx = b.take<int32_t>(4);
  } else {
    w.sa_ws_bytes = sa_workspace_bytes(c);
    w.sa_ws = b.take<char>(w.sa_ws_bytes);
    G(D, 3 * D, T); G(D, D, T);
    CS(T, 3 * D); CS(T, D); CS(B, D); CS(B, 5 * L * L); CS(T, L);
  }
  w.ce_ws_bytes = ce_workspace_bytes(B, D, c.item_rows);
  w.ce_ws = b.take<char>(w.ce_ws_bytes);
  w.gemm_ws_bytes = gemm_ws + 256;
  w.gemm_ws = b.take<char>(w.gemm_ws_bytes);
  w.colsum_ws_bytes = colsum_ws + 256;
  w.colsum_ws = b.take<char>(w.colsum_ws_bytes);
  // the parameter-gradient products run on a side stream beside the backward chain: their own split-K / partial scratch
  w.gemm_ws2 = b.take<char>(w.gemm_ws_bytes);
  w.colsum_ws2 = b.take<char>(w.colsum_ws_bytes);
  size_t sc = 0;
  sc = std::max(sc, scatter_add_workspace_bytes(T, c.item_rows, D));
  sc = std::max(sc, scatter_add_workspace_bytes(T, c.category_rows, D));
  sc = std::max(sc, scatter_add_workspace_bytes(T, c.position_rows, D));
  sc = std::max(sc, scatter_add_workspace_bytes(B, c.user_rows, D));
  sc = std::max(sc, scatter_add_workspace_bytes(B, c.item_rows, D));
  w.scatter_ws_bytes = sc;
  w.scatter_ws = b.take<char>(sc);
  {
    const int rows[4] = {c.item_rows, c.category_rows, c.position_rows, c.user_rows};
    for (int k = 0; k < 4; ++k) {
      w.sort_ws_bytes[k] = sort_workspace_bytes(k == 3 ? B : T, rows[k]);
      w.sort_ws[k] = b.take<char>(w.sort_ws_bytes[k]);
    }
    w.seg_ws_bytes = seg_reduce_workspace_bytes(T, (int)D);
    w.seg_ws = b.take<char>(w.seg_ws_bytes);
    for (int k = 0; k < 3; ++k) w.seg_ws_side[k] = b.take<char>(w.seg_ws_bytes);
  }
  w.topk_ws_bytes = score_topk_workspace_bytes((int)B, c.item_rows, std::min(50, c.item_rows));
  w.topk_ws = b.take<char>(w.topk_ws_bytes);
  w.user_marks_words = (size_t)(c.user_rows + 31) / 32 + 8;
  w.user_marks = b.take<unsigned>(w.user_marks_words);
  w.total_bytes = b.off + 1024;
  if (base && !b.ok()) return set_error(MTAM_ERR_WORKSPACE, "workspace %zu bytes < required %zu", cap, w.total_bytes);
  return 0;
}

static int validate(const mtam_config* c) {
  if (!c) return set_error(MTAM_ERR_INVALID, "config is null");
  if (c->abi_version != MTAM_ABI_VERSION)
    return set_error(MTAM_ERR_INVALID, "abi_version %d != %d", c->abi_version, MTAM_ABI_VERSION);
  if (c->kind < 0 || c->kind > MTAM_KIND_MTAM_VIA_RNN) return set_error(MTAM_ERR_INVALID, "unknown model kind %d", c->kind);
  if (c->D != 32 && c->D != 64 && c->D != 128)
    return set_error(MTAM_ERR_INVALID, "num_units=%d not supported (32, 64 or 128)", c->D);
  if (c->max_batch < 1 || c->L < 1 || c->N < 0) return set_error(MTAM_ERR_INVALID, "bad max_batch/L/N");
  if (c->H < 1 || c->H > 32 || (32 % c->H) != 0 || (c->D % c->H) != 0)
    return set_error(MTAM_ERR_INVALID, "num_heads=%d must divide 32 and num_units", c->H);
  if (c->user_rows < 1 || c->item_rows < 1 || c->category_rows < 1 || c->position_rows < 1)
    return set_error(MTAM_ERR_INVALID, "table row counts must be positive");
  if (c->gemm_mode != MTAM_GEMM_FP32 && !gemm_mode_is_tc(c->gemm_mode))
    return set_error(MTAM_ERR_INVALID, "unknown gemm_mode %d", c->gemm_mode);
  if (c->optimizer != MTAM_OPT_ADAM && c->optimizer != MTAM_OPT_SGD)
    return set_error(MTAM_ERR_UNSUPPORTED, "optimizer %d not built (adam, sgd)", c->optimizer);
  if (!(c->dropout >= 0.f && c->dropout < 1.f)) return set_error(MTAM_ERR_INVALID, "dropout rate %g outside [0,1)", c->dropout);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// MTAM forward / backward
// ---------------------------------------------------------------------------------------------
static int gemm(mtam_model* h, int tA, int tB, int M, int N, int K, const float* A, int lda, const float* Bm, int ldb,
                float* C, int ldc, const GemmEpilogue& e, cudaStream_t st) {
  return gemm_any(h->cfg.gemm_mode, tA, tB, M, N, K, A, lda, Bm, ldb, C, ldc, e, h->ws.gemm_ws, h->ws.gemm_ws_bytes, st);
}
static int colsum(mtam_model* h, const float* A, int lda, const float* Bm, int ldb, int M, int N, float* out,
                  cudaStream_t st) {
  return colsum_f32(A, lda, Bm, ldb, M, N, out, 0, h->ws.colsum_ws, h->ws.colsum_ws_bytes, st);
}

// the same on the parameter-gradient side stream (own scratch)
static int gemm2(mtam_model* h, int tA, int tB, int M, int N, int K, const float* A, int lda, const float* Bm, int ldb,
                 float* C, int ldc, const GemmEpilogue& e, cudaStream_t st) {
  return gemm_any(h->cfg.gemm_mode, tA, tB, M, N, K, A, lda, Bm, ldb, C, ldc, e, h->ws.gemm_ws2, h->ws.gemm_ws_bytes, st);
}
static int colsum2(mtam_model* h, const float* A, int lda, const float* Bm, int ldb, int M, int N, float* out,
                   cudaStream_t st) {
  return colsum_f32(A, lda, Bm, ldb, M, N, out, 0, h->ws.colsum_ws2, h->ws.colsum_ws_bytes, st);
}

static HopArgs hop_args(mtam_model* h, const mtam_batch* bt) {
  const mtam_config& c = h->cfg;
  const Layout& l = h->lay;
  Workspace& w = h->ws;
  HopArgs a;
  a.B = bt->B; a.L = c.L; a.D = c.D; a.H = c.H; a.N = c.N;
  a.seq_len = bt->seq_length; a.target_time = bt->target_item_time; a.time_list = bt->time_list;
  // the hops' memory ("user_history"): the embedded behaviours, or the T-GRU's output sequence (Hs holds h_{t-1} per
  // row with a leading zero row, so the outputs -- zero from step seq_len-1 on, as dynamic_rnn leaves them -- are Hs + D)
  a.X = memory_is_rnn(c.kind) ? w.Hs + c.D : w.X;
  a.KV = w.KV;
  a.Wq = h->params + l.Wq; a.bq = h->params + l.bq; a.Wt = h->params + l.Wt; a.gate = h->params + l.gate;
  a.ln_gamma = h->params + l.lng; a.ln_beta = h->params + l.lnb;
  a.lnf_gamma = h->params + l.lnfg; a.lnf_beta = h->params + l.lnfb;
  a.Qin = w.Qin; a.Qr = w.Qr; a.Qt = w.Qt; a.AA = w.AA; a.PA = w.PA; a.ZZ = w.ZZ; a.DK = w.DK; a.GT = w.GT;
  a.XH = w.XH; a.RSTD = w.RSTD; a.XHF = w.XHF; a.RSTDF = w.RSTDF; a.pred = w.pred;
  return a;
}

// shared by all sequence models: embedding layer forward  (Behavior_...py:62-114)
static int embed_forward(mtam_model* h, const mtam_batch* bt, int include_user, int* n_l2, cudaStream_t st) {
  const mtam_config& c = h->cfg;
  const Layout& l = h->lay;
  Workspace& w = h->ws;
  const int B = bt->B, D = c.D;
  const int64_t T = (int64_t)B * c.L;
  float* P = h->params;
  // (row-sharded item table: the rows were fetched by the caller; token t reads row t of that array)
  const float* item_tab = h->item_rows_ext ? h->item_rows_ext : P + l.item;
  MTAM_TRY(embed_gather(item_tab, P + l.cat, P + l.pos, P + l.user, bt->item_list, bt->category_list,
                        bt->position_list, bt->user_id, B, c.L, D, include_user, w.E2, w.l2_partial, n_l2, st));
  GemmEpilogue e;
  e.relu = 1;
  MTAM_TRY(gemm(h, 0, 0, (int)T, D, 2 * D, w.E2, 2 * D, P + l.Wemb, D, w.R, D, e, st));
  MTAM_TRY(add_pos(w.R, P + l.pos, bt->position_list, T, D, w.X, st));
  return 0;
}

// shared: loss scalars  (base_model.py:302-323)
static int loss_scalars(mtam_model* h, int n_l2, int n_ce, int global_batch, float* scalars_out, cudaStream_t st) {
  Workspace& w = h->ws;
  float* ds = w.dev_scalars;
  MTAM_TRY(loss_scalars_sum(w.l2_partial, n_l2, w.ce_partial, n_ce, 1.0f / (float)global_batch, h->cfg.reg,
                            ds + MTAM_S_L2_NORM, ds + MTAM_S_LOSS_ORIGIN, ds + MTAM_S_LOSS, st));
  if (scalars_out)
    MTAM_CUDA_CHECK(cudaMemcpyAsync(scalars_out, ds, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

static int mtam_fwd(mtam_model* h, const mtam_batch* bt, int global_batch, float* scalars_out, bool with_loss,
                    cudaStream_t st) {
  const mtam_config& c = h->cfg;
  const Layout& l = h->lay;
  Workspace& w = h->ws;
  const int B = bt->B, D = c.D, N = c.N, L = c.L;
  const int64_t T = (int64_t)B * L;
  float* P = h->params;
  int n_l2 = 0, n_ce = 0;
  phase(h, MTAM_PH_EMBED_FWD, st);
  MTAM_TRY(embed_forward(h, bt, 1, &n_l2, st));
  phase(h, MTAM_PH_GRU_X_GEMM, st);
  {  // x-side GRU pre-activations: [T,D] x [D,3D] + [bg|bc]
    GemmEpilogue e;
    e.bias = P + l.bgru;
    MTAM_TRY(gemm(h, 0, 0, (int)T, 3 * D, D, w.X, D, P + l.Wgru, 3 * D, w.GX, 3 * D, e, st));
  }
  phase(h, MTAM_PH_GRU_FWD, st);
  MTAM_CUDA_CHECK(cudaMemsetAsync(w.Hs, 0, (size_t)(T + 1) * D * sizeof(float), st));
  MTAM_CUDA_CHECK(cudaMemsetAsync(w.RH, 0, (size_t)T * D * sizeof(float), st));
  const bool via = memory_is_rnn(c.kind);
  MTAM_TRY(gru_forward(D, w.X, w.GX, bt->timelast_list, bt->seq_length, P + l.Wgru, P + l.gruvec, B, L, w.Hs, w.RUCT,
                       w.RH, via ? w.q0raw : w.Qin, st, plain_gru(c.kind) ? 1 : 0, gemm_mode_is_tc(c.gemm_mode) ? 1 : 0));
  if (via)   // short_term_intent = layer_norm(gather(...))  (MTAMRec_model.py:182-187)
    MTAM_TRY(ln_rows_forward(w.q0raw, B, D, P + l.lnsg, P + l.lnsb, w.Qin, w.XHS, w.RSTDS, st));
  phase(h, MTAM_PH_KV_GEMM, st);
  {  // K,V of all hops: relu(memory [T,D] x [D,2ND] + b)
    GemmEpilogue e;
    e.bias = P + l.bkv;
    e.relu = 1;
    MTAM_TRY(gemm(h, 0, 0, (int)T, 2 * N * D, D, via ? w.Hs + D : w.X, D, P + l.Wkv, 2 * N * D, w.KV, 2 * N * D, e, st));
  }
  phase(h, MTAM_PH_HOP_FWD, st);
  HopArgs a = hop_args(h, bt);
  MTAM_TRY(hop_forward(a, st));
  phase(h, MTAM_PH_CE_FWD, st);
  if (h->rows_mode) {   // the softmax against the sharded table is the caller's: leave the L2 term of this rank's rows
    MTAM_TRY(finalize_sum(w.l2_partial, n_l2, 0.5f, w.dev_scalars + MTAM_S_L2_NORM, 0, st));
    if (scalars_out)
      MTAM_CUDA_CHECK(cudaMemcpyAsync(scalars_out + MTAM_S_L2_NORM, w.dev_scalars + MTAM_S_L2_NORM, sizeof(float),
                                      cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  if (!with_loss) return 0;
  MTAM_CUDA_CHECK(cudaMemsetAsync(w.tlogit, 0, (size_t)B * sizeof(float), st));
  MTAM_TRY(ce_forward(c.gemm_mode, D, w.pred, P + l.item, bt->target_item_id, B, c.item_rows, w.ce_ws, w.tlogit, w.lse,
                      w.loss_origin, w.ce_partial, &n_ce, st));
  MTAM_TRY(loss_scalars(h, n_l2, n_ce, global_batch, scalars_out, st));
  return 0;
}

// shared: embedding-layer backward from dX; leaves the IndexedSlices values in dE2/dEp/dEu and
// adds their squared norm to *norm_sq_sparse.
static int embed_backward(mtam_model* h, const mtam_batch* bt, int include_user, float* norm_sq_sparse,
                          cudaStream_t st, cudaStream_t pg = nullptr, cudaEvent_t ev_pg = nullptr) {
  const mtam_config& c = h->cfg;
  const Layout& l = h->lay;
  Workspace& w = h->ws;
  const int B = bt->B, D = c.D;
  const int64_t T = (int64_t)B * c.L;
  float* P = h->params;
  float* G = h->grads;
  MTAM_TRY(relu_mask(w.dX, w.R, T * D, w.dR, st));
  GemmEpilogue e;
  if (pg) {   // the weight gradient runs beside the rest of the chain (the caller joins `pg` afterwards)
    MTAM_CUDA_CHECK(cudaEventRecord(ev_pg, st));
    MTAM_CUDA_CHECK(cudaStreamWaitEvent(pg, ev_pg, 0));
    MTAM_TRY(gemm2(h, 1, 0, 2 * D, D, (int)T, w.E2, 2 * D, w.dR, D, G + l.Wemb, D, e, pg));
  } else {
    MTAM_TRY(gemm(h, 1, 0, 2 * D, D, (int)T, w.E2, 2 * D, w.dR, D, G + l.Wemb, D, e, st));     // dW = E2^T dR
  }
  MTAM_TRY(gemm(h, 0, 1, (int)T, 2 * D, D, w.dR, D, P + l.Wemb, D, w.dE2, 2 * D, e, st));    // dE2 = dR W^T
  int n_sq = 0;
  MTAM_TRY(embed_bwd_tail(w.E2, w.dX, P + l.pos, P + l.user, bt->position_list, bt->user_id, B, c.L, D, c.reg,
                          include_user, w.dE2, w.dEp, w.dEu, w.sq_partial, &n_sq, st));
  MTAM_TRY(finalize_sum(w.sq_partial, n_sq, 1.0f, norm_sq_sparse, 1, st));
  return 0;
}

static int mtam_bwd(mtam_model* h, const mtam_batch* bt, int global_batch, float* norm_sq_sparse, cudaStream_t st) {
  const mtam_config& c = h->cfg;
  const Layout& l = h->lay;
  Workspace& w = h->ws;
  const int B = bt->B, D = c.D, N = c.N, L = c.L;
  const int64_t T = (int64_t)B * L;
  float* P = h->params;
  float* G = h->grads;
  GemmEpilogue e0, eacc;
  eacc.accumulate = 1;
  // softmax CE: dense item-table gradient straight into the arena, dpred
  phase(h, MTAM_PH_CE_BWD, st);
  // The dense item-table gradient (dT = G^T pred, a full-machine tensor-core pass) feeds only the norm and Adam; the
  // chain behind dpred (hops, T-GRU, embedding) is latency-bound kernels on a fraction of the SMs.  Run dT beside it.
  // (Not while profiling: the per-phase times would no longer add up.)
  const bool dt_aside = gemm_mode_is_tc(c.gemm_mode) && ce_tc_supported(D) && !h->prof && !h->rows_mode;
  const int ce_terms = gemm_mode_ce_terms(c.gemm_mode);
  if (h->rows_mode) {
    // dpred was placed in the workspace by mtam_backward_rows (softmax against the sharded table: the caller's)
  } else if (dt_aside) {
    MTAM_CUDA_CHECK(cudaEventRecord(h->ev_fork2, st));
    MTAM_CUDA_CHECK(cudaStreamWaitEvent(h->side2, h->ev_fork2, 0));
    MTAM_TRY(ce_backward_tc(D, w.pred, P + l.item, bt->target_item_id, w.lse, B, c.item_rows, 1.0f / (float)global_batch,
                            w.ce_ws, G + l.item, w.dpred, h->side2, 2, ce_terms));
    if (h->ev_ce_done) MTAM_CUDA_CHECK(cudaEventRecord(h->ev_ce_done, h->side2));
    MTAM_CUDA_CHECK(cudaEventRecord(h->ev_join2, h->side2));
    MTAM_TRY(ce_backward_tc(D, w.pred, P + l.item, bt->target_item_id, w.lse, B, c.item_rows, 1.0f / (float)global_batch,
                            w.ce_ws, G + l.item, w.dpred, st, 1, ce_terms));
  } else {
    MTAM_TRY(ce_backward(c.gemm_mode, D, w.pred, P + l.item, bt->target_item_id, w.lse, B, c.item_rows, 1.0f / (float)global_batch,
                         w.ce_ws, G + l.item, w.dpred, st));
    if (h->ev_ce_done) MTAM_CUDA_CHECK(cudaEventRecord(h->ev_ce_done, st));
  }
  // hops
  phase(h, MTAM_PH_HOP_BWD, st);
  const bool via = memory_is_rnn(c.kind);
  // gradient w.r.t. the hops' memory: dX itself, or (MTAM_via_T_GRU) the gradient of every T-GRU output step, kept in
  // the dR buffer (free until the embedding backward)
  float* dMem = via ? w.dR : w.dX;
  const float* mem = via ? w.Hs + D : w.X;
  MTAM_CUDA_CHECK(cudaMemsetAsync(w.dX, 0, (size_t)T * D * sizeof(float), st));
  if (via) MTAM_CUDA_CHECK(cudaMemsetAsync(dMem, 0, (size_t)T * D * sizeof(float), st));
  const bool kv_bias_fused = hop_backward_writes_all_dkv(D, c.H, L, N);   // = the CTA-per-sequence path is taken
  if (!kv_bias_fused)   // the warp-per-sequence kernels leave masked keys untouched
    MTAM_CUDA_CHECK(cudaMemsetAsync(w.dKV, 0, (size_t)T * 2 * N * D * sizeof(float), st));
  MTAM_CUDA_CHECK(cudaMemsetAsync(w.GB, 0, (size_t)B * 5 * N * L * sizeof(float), st));
  MTAM_TRY(transpose_dd(P + l.Wq, w.WqT, D, N, st));
  MTAM_TRY(transpose_dd(P + l.Wt, w.WtT, D, N, st));
  HopArgs a = hop_args(h, bt);
  HopGradArgs g;
  g.dpred = w.dpred; g.WqT = w.WqT; g.WtT = w.WtT; g.dX = dMem; g.dKV = w.dKV; g.DOUT = w.DOUT; g.DQP = w.DQP; g.BKV = w.BKV;
  g.DQT = w.DQT; g.GB = w.GB; g.dq0 = w.dq0;
  MTAM_TRY(hop_backward(a, g, st));
  // Parameter gradients (weight-gradient GEMMs with K = B*L, column sums) feed nothing but the norm and the optimizer, so
  // they run on a side stream (`pg`) beside the chain dX -> T-GRU backward -> embedding backward, joining before the
  // norm.  (Not while profiling: the per-phase times would no longer add up.)
  const bool pg_aside = !h->prof;
  cudaStream_t pg = pg_aside ? h->side3 : st;
  auto pg_after_main = [&](int i) -> int {   // pg continues once the main stream has reached this point
    if (!pg_aside) return 0;
    MTAM_CUDA_CHECK(cudaEventRecord(h->ev_pg[i], st));
    MTAM_CUDA_CHECK(cudaStreamWaitEvent(pg, h->ev_pg[i], 0));
    return 0;
  };
  // hop parameter gradients
  phase(h, MTAM_PH_HOP_PARAM_GRADS, st);
  MTAM_TRY(pg_after_main(0));
  {  // the six per-sequence column sums (LN gains / biases, query bias, gate vectors) in one launch
    ColsumBatch cb{};
    cb.job[0] = ColsumJob{w.dpred, D, nullptr, 0, D, G + l.lnfb};
    cb.job[1] = ColsumJob{w.dpred, D, w.XHF, D, D, G + l.lnfg};
    cb.job[2] = ColsumJob{w.DOUT, N * D, nullptr, 0, N * D, G + l.lnb};
    cb.job[3] = ColsumJob{w.DOUT, N * D, w.XH, N * D, N * D, G + l.lng};
    cb.job[4] = ColsumJob{w.DQP, N * D, nullptr, 0, N * D, G + l.bq};
    cb.job[5] = ColsumJob{w.GB, 5 * N * L, nullptr, 0, 5 * N * L, G + l.gate};
    cb.n_jobs = 6;
    // the K,V bias gradient: the CTA-per-sequence backward leaves per-sequence column sums of dKV ([B, 2ND] instead of
    // a second pass over the [B*L, 2ND] rows)
    if (kv_bias_fused) cb.job[cb.n_jobs++] = ColsumJob{w.BKV, 2 * N * D, nullptr, 0, 2 * N * D, G + l.bkv};
    MTAM_TRY(colsum_multi_f32(cb, B, pg));
    if (via) {   // the short-term intent's layer norm: gain and bias gradients from d loss / d (its output) = dq0
      ColsumBatch c2{};
      c2.job[0] = ColsumJob{w.dq0, D, nullptr, 0, D, G + l.lnsb};
      c2.job[1] = ColsumJob{w.dq0, D, w.XHS, D, D, G + l.lnsg};
      c2.n_jobs = 2;
      MTAM_TRY(colsum_multi_f32(c2, B, pg));
    }
  }
  // dWq_i = Qin_i^T dQpre_i, dWt_i = Qin_i^T dQt_i: the N hops of each in one batched split-K launch
  MTAM_TRY(gemm_atb_batched_f32(N, D, D, B, w.Qin, D, (int64_t)B * D, w.DQP, N * D, D, G + l.Wq, D, (int64_t)D * D,
                                w.gemm_ws2, w.gemm_ws_bytes, pg));
  MTAM_TRY(gemm_atb_batched_f32(N, D, D, B, w.Qin, D, (int64_t)B * D, w.DQT, N * D, D, G + l.Wt, D, (int64_t)D * D,
                                w.gemm_ws2, w.gemm_ws_bytes, pg));
  if (!kv_bias_fused) MTAM_TRY(colsum2(h, w.dKV, 2 * N * D, nullptr, 0, (int)T, 2 * N * D, G + l.bkv, pg));
  MTAM_TRY(gemm2(h, 1, 0, D, 2 * N * D, (int)T, mem, D, w.dKV, 2 * N * D, G + l.Wkv, 2 * N * D, e0, pg));
  MTAM_TRY(gemm(h, 0, 1, (int)T, D, 2 * N * D, w.dKV, 2 * N * D, P + l.Wkv, 2 * N * D, dMem, D, eacc, st));
  const float* dq0 = w.dq0;
  if (via) {   // through the layer norm of the query
    MTAM_TRY(ln_rows_backward(w.dq0, w.XHS, w.RSTDS, B, D, P + l.lnsg, w.dq0raw, st));
    dq0 = w.dq0raw;
  }
  // T-GRU
  phase(h, MTAM_PH_GRU_BWD, st);
  MTAM_CUDA_CHECK(cudaMemsetAsync(w.dGX, 0, (size_t)T * 3 * D * sizeof(float), st));
  MTAM_TRY(gru_backward(D, w.X, bt->timelast_list, bt->seq_length, P + l.Wgru, P + l.gruvec, w.Hs, w.RUCT, dq0,
                        via ? dMem : nullptr, B, L, w.dGX, w.dX, w.vec_partial, st, gemm_mode_is_tc(c.gemm_mode) ? 1 : 0));
  phase(h, MTAM_PH_GRU_PARAM_GRADS, st);
  MTAM_TRY(pg_after_main(1));
  MTAM_TRY(colsum2(h, w.vec_partial, 8 * D, nullptr, 0, gru_num_blocks(B), 8 * D, G + l.gruvec, pg));
  MTAM_TRY(colsum2(h, w.dGX, 3 * D, nullptr, 0, (int)T, 3 * D, G + l.bgru, pg));
  MTAM_TRY(gemm2(h, 1, 0, D, 3 * D, (int)T, w.X, D, w.dGX, 3 * D, G + l.Wgru, 3 * D, e0, pg));
  // h-side: h_{t-1} = Hs shifted by one row (leading zero row), r*h_{t-1} = RH
  MTAM_TRY(gemm2(h, 1, 0, D, 2 * D, (int)T, w.Hs, D, w.dGX, 3 * D, G + l.Wgru + (size_t)D * 3 * D, 3 * D, e0, pg));
  MTAM_TRY(gemm2(h, 1, 0, D, D, (int)T, w.RH, D, w.dGX + 2 * D, 3 * D, G + l.Wgru + (size_t)D * 3 * D + 2 * D, 3 * D, e0, pg));
  MTAM_TRY(gemm(h, 0, 1, (int)T, D, 3 * D, w.dGX, 3 * D, P + l.Wgru, 3 * D, w.dX, D, eacc, st));
  phase(h, MTAM_PH_EMBED_BWD, st);
  MTAM_TRY(embed_backward(h, bt, 1, norm_sq_sparse, st, pg_aside ? pg : nullptr, pg_aside ? h->ev_pg[2] : nullptr));
  if (pg_aside) {
    MTAM_CUDA_CHECK(cudaEventRecord(h->ev_join3, pg));
    MTAM_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_join3, 0));
  }
  if (dt_aside) MTAM_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_join2, 0));
  phase(h, MTAM_PH_DENSE_NORM, st);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// self-attention family: PISTRec / SASRec / TA-SASRec / TiSASRec
// ---------------------------------------------------------------------------------------------
static SaCtx sa_ctx(mtam_model* h, const mtam_batch* bt) {
  SaCtx c;
  c.cfg = h->cfg; c.sl = h->lay.sa; c.params = h->params; c.grads = h->grads; c.bt = bt;
  c.X = h->ws.X; c.dX = h->ws.dX; c.pred = h->ws.pred; c.dpred = h->ws.dpred; c.XHF = h->ws.XHF; c.RSTDF = h->ws.RSTDF;
  c.ws = h->ws.sa_ws; c.ws_bytes = h->ws.sa_ws_bytes;
  c.gemm_ws = h->ws.gemm_ws; c.gemm_ws_bytes = h->ws.gemm_ws_bytes;
  c.colsum_ws = h->ws.colsum_ws; c.colsum_ws_bytes = h->ws.colsum_ws_bytes;
  c.drop_rate = h->cfg.dropout; c.drop_seed = h->drop_seed;
  c.drop_counter = reinterpret_cast<const uint32_t*>(h->ws.dev_scalars + 10);
  return c;
}

// every forward pass of a model with attention dropout draws a fresh mask: advance the call counter unless
// mtam_prepare_step / mtam_set_dropout_state already placed the one this pass is to use on the device
static int next_dropout_call(mtam_model* h, cudaStream_t st) {
  const int k = h->cfg.kind;
  if (h->cfg.dropout <= 0.f || (k != MTAM_KIND_SASREC && k != MTAM_KIND_TISASREC)) return 0;
  h->drop_explicit = false;
  if (h->drop_prepared) { h->drop_prepared = false; return 0; }
  h->drop_calls += 1;
  return set_scalar_u32(reinterpret_cast<uint32_t*>(h->ws.dev_scalars + 10), h->drop_calls, st);
}

static int sa_fwd(mtam_model* h, const mtam_batch* bt, int global_batch, float* scalars_out, bool with_loss,
                  cudaStream_t st) {
  const mtam_config& c = h->cfg;
  Workspace& w = h->ws;
  const int include_user = c.kind != MTAM_KIND_PISTREC;   // PISTRec_model.py:56-60 omits the user L2 term
  int n_l2 = 0, n_ce = 0;
  phase(h, MTAM_PH_EMBED_FWD, st);
  MTAM_TRY(embed_forward(h, bt, include_user, &n_l2, st));
  phase(h, MTAM_PH_HOP_FWD, st);
  MTAM_TRY(sa_forward(sa_ctx(h, bt), st));
  phase(h, MTAM_PH_CE_FWD, st);
  if (!with_loss) return 0;
  MTAM_CUDA_CHECK(cudaMemsetAsync(w.tlogit, 0, (size_t)bt->B * sizeof(float), st));
  MTAM_TRY(ce_forward(c.gemm_mode, c.D, w.pred, h->params + h->lay.item, bt->target_item_id, bt->B, c.item_rows, w.ce_ws, w.tlogit,
                      w.lse, w.loss_origin, w.ce_partial, &n_ce, st));
  MTAM_TRY(loss_scalars(h, n_l2, n_ce, global_batch, scalars_out, st));
  return 0;
}

static int sa_bwd(mtam_model* h, const mtam_batch* bt, int global_batch, float* norm_sq_sparse, cudaStream_t st) {
  const mtam_config& c = h->cfg;
  Workspace& w = h->ws;
  phase(h, MTAM_PH_CE_BWD, st);
  MTAM_TRY(ce_backward(c.gemm_mode, c.D, w.pred, h->params + h->lay.item, bt->target_item_id, w.lse, bt->B, c.item_rows,
                       1.0f / (float)global_batch, w.ce_ws, h->grads + h->lay.item, w.dpred, st));
  if (h->ev_ce_done) MTAM_CUDA_CHECK(cudaEventRecord(h->ev_ce_done, st));
  phase(h, MTAM_PH_HOP_BWD, st);
  MTAM_TRY(sa_backward(sa_ctx(h, bt), st));
  phase(h, MTAM_PH_EMBED_BWD, st);
  MTAM_TRY(embed_backward(h, bt, c.kind != MTAM_KIND_PISTREC, norm_sq_sparse, st));
  phase(h, MTAM_PH_DENSE_NORM, st);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// BPR-MF
// ---------------------------------------------------------------------------------------------
static BprArgs bpr_args(mtam_model* h, const mtam_batch* bt) {
  Workspace& w = h->ws;
  const Layout& l = h->lay;
  BprArgs a;
  a.B = bt->B; a.D = h->cfg.D; a.neg = h->bpr_neg_used;
  a.Tu = h->params + l.user; a.Ti = h->params + l.item; a.Tb = h->params + l.item_b;
  a.user = bt->user_id; a.target = bt->target_item_id;
  a.U = w.bU; a.IP = w.bIP; a.dot = w.bdot; a.bpos = w.bbpos; a.rowsum = w.browsum; a.colsum = w.bcolsum;
  a.loss_partial = w.bloss; a.dU = w.bdU; a.dIP = w.bdIP; a.dINpart = w.bdINp; a.dIN = w.bdIN; a.dbneg = w.bdbneg;
  a.l2_partial = w.bl2; a.sq_partial = w.bsq;
  return a;
}

static int bpr_fwd(mtam_model* h, const mtam_batch* bt, float* scalars_out, bool with_loss, cudaStream_t st) {
  Workspace& w = h->ws;
  if (h->bpr_neg >= 0) {
    h->bpr_neg_used = h->bpr_neg;
  } else {   // tf.random_uniform([1], 0, item_count) (BPRMF.py:43): a fresh negative per step
    h->rng = h->rng * 6364136223846793005ULL + 1442695040888963407ULL;
    h->bpr_neg_used = (int)((h->rng >> 33) % (uint64_t)std::max(1, h->cfg.item_rows - 3));
  }
  int n_l2 = 0, n_loss = 0;
  MTAM_TRY(bpr_forward(bpr_args(h, bt), &n_l2, &n_loss, st));
  MTAM_CUDA_CHECK(cudaMemcpyAsync(w.pred, w.bU, (size_t)bt->B * h->cfg.D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (!with_loss) return 0;
  float* ds = w.dev_scalars;
  MTAM_TRY(finalize_sum(w.bl2, n_l2, 0.5f, ds + MTAM_S_L2_NORM, 0, st));
  MTAM_TRY(finalize_sum(w.bloss, n_loss, 1.0f / ((float)bt->B * (float)bt->B), ds + MTAM_S_LOSS_ORIGIN, 0, st));
  MTAM_TRY(finalize_sum(ds + MTAM_S_L2_NORM, 1, 5e-5f, ds + MTAM_S_LOSS, 0, st));
  MTAM_TRY(finalize_sum(ds + MTAM_S_LOSS_ORIGIN, 1, 1.0f, ds + MTAM_S_LOSS, 1, st));
  if (scalars_out) MTAM_CUDA_CHECK(cudaMemcpyAsync(scalars_out, ds, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

static int bpr_bwd(mtam_model* h, const mtam_batch* bt, float* norm_sq_sparse, cudaStream_t st) {
  int n_sq = 0;
  MTAM_TRY(bpr_backward(bpr_args(h, bt), &n_sq, st));
  MTAM_TRY(finalize_sum(h->ws.bsq, n_sq, 1.0f, norm_sq_sparse, 1, st));
  // every piece is sparse: the dense region of the grads arena must read as zero for the norm
  MTAM_CUDA_CHECK(cudaMemsetAsync(h->grads + h->lay.dense_begin, 0, (h->lay.total - h->lay.dense_begin) * sizeof(float), st));
  return 0;
}

static int bpr_scatter(mtam_model* h, cudaStream_t st) {
  const mtam_config& c = h->cfg;
  const Layout& l = h->lay;
  Workspace& w = h->ws;
  const mtam_batch& bt = h->last_batch;
  const int B = h->last_B, D = c.D;
  float* G = h->grads;
  int32_t neg = h->bpr_neg_used;
  MTAM_CUDA_CHECK(cudaMemcpyAsync(w.bneg_idx, &h->bpr_neg_used, sizeof(int32_t), cudaMemcpyHostToDevice, st));
  (void)neg;
  MTAM_TRY(scatter_add_rows(G + l.user, c.user_rows, D, D, bt.user_id, w.bdU, D, B, w.scatter_ws, w.scatter_ws_bytes, nullptr, nullptr, st));
  MTAM_TRY(scatter_add_rows(G + l.item, c.item_rows, D, D, bt.target_item_id, w.bdIP, D, B, w.scatter_ws, w.scatter_ws_bytes, nullptr, nullptr, st));
  MTAM_TRY(scatter_add_rows(G + l.item, c.item_rows, D, D, w.bneg_idx, w.bdIN, D, 1, w.scatter_ws, w.scatter_ws_bytes, nullptr, nullptr, st));
  MTAM_TRY(scatter_add_rows(G + l.item_b, c.item_rows, 1, 1, bt.target_item_id, w.browsum, 1, B, w.scatter_ws, w.scatter_ws_bytes, nullptr, nullptr, st));
  MTAM_TRY(scatter_add_rows(G + l.item_b, c.item_rows, 1, 1, w.bneg_idx, w.bdbneg, 1, 1, w.scatter_ws, w.scatter_ws_bytes, nullptr, nullptr, st));
  return 0;
}

static int check_batch(mtam_model* h, const mtam_batch* bt) {
  if (!h) return set_error(MTAM_ERR_INVALID, "null handle");
  if (!bt) return set_error(MTAM_ERR_INVALID, "null batch");
  if (bt->B < 1 || bt->B > h->cfg.max_batch)
    return set_error(MTAM_ERR_INVALID, "batch size %d outside [1, max_batch=%d]", bt->B, h->cfg.max_batch);
  if (!bt->user_id || !bt->item_list || !bt->category_list || !bt->position_list || !bt->time_list ||
      !bt->timelast_list || !bt->target_item_id || !bt->target_item_time || !bt->seq_length)
    return set_error(MTAM_ERR_INVALID, "batch has a null feed array");
  return 0;
}

static int fwd_dispatch(mtam_model* h, const mtam_batch* bt, int gb, float* scalars_out, bool with_loss,
                        cudaStream_t st) {
  switch (h->cfg.kind) {
    case MTAM_KIND_MTAM:
    case MTAM_KIND_MTAM_VIA_T_GRU:
    case MTAM_KIND_MTAM_NO_TIME_AWARE_RNN:
    case MTAM_KIND_MTAM_VIA_RNN: return mtam_fwd(h, bt, gb, scalars_out, with_loss, st);
    case MTAM_KIND_BPRMF: return bpr_fwd(h, bt, scalars_out, with_loss, st);
    default:
      MTAM_TRY(next_dropout_call(h, st));
      return sa_fwd(h, bt, gb, scalars_out, with_loss, st);
  }
}
static int bwd_dispatch(mtam_model* h, const mtam_batch* bt, int gb, float* nsq, cudaStream_t st) {
  switch (h->cfg.kind) {
    case MTAM_KIND_MTAM:
    case MTAM_KIND_MTAM_VIA_T_GRU:
    case MTAM_KIND_MTAM_NO_TIME_AWARE_RNN:
    case MTAM_KIND_MTAM_VIA_RNN: return mtam_bwd(h, bt, gb, nsq, st);
    case MTAM_KIND_BPRMF: return bpr_bwd(h, bt, nsq, st);
    default: return sa_bwd(h, bt, gb, nsq, st);
  }
}

}  // namespace mtam

// =============================================================================================
// C-ABI
// =============================================================================================
extern "C" {

const char* mtam_last_error(mtam_handle h) {
  (void)h;
  return mtam::last_error_slot().c_str();
}

int mtam_plan(const mtam_config* cfg, mtam_sizes* out) {
  MTAM_TRY(validate(cfg));
  if (!out) return set_error(MTAM_ERR_INVALID, "out is null");
  Layout l;
  MTAM_TRY(build_layout(*cfg, l));
  Workspace w;
  MTAM_TRY(plan_workspace(*cfg, nullptr, 0, w));
  out->param_floats = l.total;
  out->workspace_bytes = w.total_bytes;
  return 0;
}

int mtam_create(const mtam_config* cfg, float* params, float* grads, float* adam_m, float* adam_v, void* workspace,
                size_t workspace_bytes, mtam_handle* out) {
  MTAM_TRY(validate(cfg));
  if (!params || !grads || !adam_m || !adam_v || !workspace || !out)
    return set_error(MTAM_ERR_INVALID, "mtam_create: null arena / out pointer");
  if (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)adam_m | (uintptr_t)adam_v | (uintptr_t)workspace) & 15)
    return set_error(MTAM_ERR_INVALID, "mtam_create: arenas must be 16-byte aligned");
  mtam_model* h = new mtam_model();
  h->cfg = *cfg;
  int s = build_layout(*cfg, h->lay);
  if (s == 0) s = plan_workspace(*cfg, workspace, workspace_bytes, h->ws);
  if (s != 0) {
    delete h;
    return s;
  }
  h->params = params; h->grads = grads; h->m = adam_m; h->v = adam_v;
  if (cudaMallocHost((void**)&h->lr_host, 64) != cudaSuccess) {
    delete h;
    return set_error(MTAM_ERR_CUDA, "cudaMallocHost failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  memset(&h->last_batch, 0, sizeof(h->last_batch));
  if (fill_iota(h->ws.iota, (int64_t)cfg->max_batch * cfg->L, nullptr) != 0 ||
      cudaMemsetAsync(h->ws.user_marks, 0, h->ws.user_marks_words * sizeof(unsigned), nullptr) != cudaSuccess ||
      cudaStreamSynchronize(nullptr) != cudaSuccess) {
    delete h;
    return set_error(MTAM_ERR_CUDA, "workspace initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->side2, cudaStreamNonBlocking, 0) != cudaSuccess ||   /* lowest priority */
      cudaEventCreateWithFlags(&h->ev_fork2, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_join2, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->side3, cudaStreamNonBlocking, 0) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_pg[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_pg[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_pg[2], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_join3, cudaEventDisableTiming) != cudaSuccess ||
      cudaStreamCreateWithPriority(&h->side4, cudaStreamNonBlocking, 0) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork4, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_join4, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_sc_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_sc[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_sc[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_sc[2], cudaEventDisableTiming) != cudaSuccess) {
    int e = set_error(MTAM_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    mtam_destroy(h);
    return e;
  }
  *out = h;
  return 0;
}

int mtam_destroy(mtam_handle h) {
  if (h) {
    if (h->lr_host) cudaFreeHost(h->lr_host);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->side2) cudaStreamDestroy(h->side2);
    if (h->ev_fork2) cudaEventDestroy(h->ev_fork2);
    if (h->ev_join2) cudaEventDestroy(h->ev_join2);
    if (h->side3) cudaStreamDestroy(h->side3);
    for (int i = 0; i < 3; ++i)
      if (h->ev_pg[i]) cudaEventDestroy(h->ev_pg[i]);
    if (h->ev_join3) cudaEventDestroy(h->ev_join3);
    if (h->side4) cudaStreamDestroy(h->side4);
    if (h->ev_fork4) cudaEventDestroy(h->ev_fork4);
    if (h->ev_join4) cudaEventDestroy(h->ev_join4);
    if (h->ev_sc_fork) cudaEventDestroy(h->ev_sc_fork);
    for (int i = 0; i < 3; ++i)
      if (h->ev_sc[i]) cudaEventDestroy(h->ev_sc[i]);
    for (int i = 0; i <= MTAM_PHASE_COUNT; ++i)
      if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  }
  delete h;
  return 0;
}

int mtam_param_count(mtam_handle h) { return h ? (int)h->lay.params.size() : set_error(MTAM_ERR_INVALID, "null handle"); }

int mtam_param_info_get(mtam_handle h, int32_t i, mtam_param_info* out) {
  if (!h || !out || i < 0 || i >= (int)h->lay.params.size()) return set_error(MTAM_ERR_INVALID, "bad param index");
  const ParamDesc& p = h->lay.params[i];
  memset(out, 0, sizeof(*out));
  strncpy(out->name, p.name.c_str(), MTAM_NAME_MAX - 1);
  out->rows = p.rows; out->cols = p.cols; out->ndim = p.ndim; out->ld = p.ld; out->offset = p.off; out->flags = p.flags;
  return 0;
}

int mtam_get_adam_step(mtam_handle h, int64_t* t) {
  if (!h || !t) return set_error(MTAM_ERR_INVALID, "null argument");
  *t = h->adam_t;
  return 0;
}
int mtam_set_adam_step(mtam_handle h, int64_t t) {
  if (!h || t < 0) return set_error(MTAM_ERR_INVALID, "bad argument");
  h->adam_t = t;
  h->b1_pow = 1.f; h->b2_pow = 1.f;
  for (int64_t i = 0; i < t; ++i) { h->b1_pow *= h->cfg.beta1; h->b2_pow *= h->cfg.beta2; }
  return 0;
}

int mtam_set_bpr_negative(mtam_handle h, int32_t item_id) {
  if (!h) return set_error(MTAM_ERR_INVALID, "null handle");
  if (item_id >= h->cfg.item_rows) return set_error(MTAM_ERR_INVALID, "negative item id out of range");
  h->bpr_neg = item_id;
  return 0;
}

int mtam_prepare_step(mtam_handle h, double lr, void* stream) {
  if (!h) return set_error(MTAM_ERR_INVALID, "null handle");
  h->adam_t += 1;
  h->b1_pow *= h->cfg.beta1;
  h->b2_pow *= h->cfg.beta2;
  const float lr32 = (float)lr;  // float64 placeholder cast to fp32 (base_model.py:25)
  const float lr_t = h->cfg.optimizer == MTAM_OPT_SGD ? lr32 : lr32 * sqrtf(1.f - h->b2_pow) / (1.f - h->b1_pow);
  // lr_t travels as a kernel argument (captured by value at launch): no host buffer can be overwritten
  // before the device has read it, however far the host runs ahead.
  MTAM_TRY(set_scalar(h->ws.dev_scalars + 9, lr_t, (cudaStream_t)stream));
  if (h->cfg.dropout > 0.f && (h->cfg.kind == MTAM_KIND_SASREC || h->cfg.kind == MTAM_KIND_TISASREC)) {
    // a replayed CUDA graph runs no host code: the mask counter of the NEXT forward pass travels the same way as lr_t.
    // (In an eager step this call comes after the step's own forward pass and prepares the following one.)
    if (h->drop_explicit) {
      h->drop_explicit = false;          // mtam_set_dropout_state chose the counter: keep it
    } else {
      h->drop_calls += 1;
      MTAM_TRY(set_scalar_u32(reinterpret_cast<uint32_t*>(h->ws.dev_scalars + 10), h->drop_calls, (cudaStream_t)stream));
    }
    h->drop_prepared = true;
  }
  h->step_prepared = true;
  return 0;
}

int mtam_set_dropout_state(mtam_handle h, uint32_t seed, uint32_t counter) {
  if (!h) return set_error(MTAM_ERR_INVALID, "null handle");
  h->drop_seed = seed;
  h->drop_calls = counter;
  MTAM_TRY(set_scalar_u32(reinterpret_cast<uint32_t*>(h->ws.dev_scalars + 10), counter, nullptr));
  MTAM_CUDA_CHECK(cudaStreamSynchronize(nullptr));
  h->drop_prepared = true;
  h->drop_explicit = true;
  return 0;
}

int mtam_profile_enable(mtam_handle h, int32_t on) {
  if (!h) return set_error(MTAM_ERR_INVALID, "null handle");
  h->prof = on != 0;
  for (int i = 0; i <= MTAM_PHASE_COUNT; ++i) h->ev_valid[i] = false;
  return 0;
}

int mtam_profile_read(mtam_handle h, float* ms_out, int32_t n) {
  if (!h || !ms_out || n < MTAM_PHASE_COUNT) return set_error(MTAM_ERR_INVALID, "mtam_profile_read: bad argument");
  if (!h->ev_valid[MTAM_PHASE_COUNT]) return set_error(MTAM_ERR_INVALID, "no profiled step recorded");
  MTAM_CUDA_CHECK(cudaEventSynchronize(h->ev[MTAM_PHASE_COUNT]));
  for (int i = 0; i < MTAM_PHASE_COUNT; ++i) {
    ms_out[i] = 0.f;
    if (!h->ev_valid[i]) continue;
    int j = i + 1;
    while (j < MTAM_PHASE_COUNT && !h->ev_valid[j]) ++j;
    MTAM_CUDA_CHECK(cudaEventElapsedTime(&ms_out[i], h->ev[i], h->ev[j]));
  }
  return 0;
}

int mtam_sparse_pieces(mtam_handle h, mtam_sparse_view* out) {
  if (!h || !out) return set_error(MTAM_ERR_INVALID, "null argument");
  memset(out, 0, sizeof(*out));
  out->B = h->last_B; out->L = h->cfg.L; out->D = h->cfg.D;
  out->item_cat_rows = h->ws.dE2; out->position_rows = h->ws.dEp; out->user_rows = h->ws.dEu;
  out->dense_begin = h->lay.dense_begin;
  out->user_offset = h->lay.user; out->item_offset = h->lay.item; out->category_offset = h->lay.cat;
  out->position_offset = h->lay.pos;
  out->has_user = h->cfg.kind != MTAM_KIND_PISTREC;
  return 0;
}

int mtam_forward(mtam_handle h, const mtam_batch* batch, float* scalars_out, float* loss_origin_out, float* pred_out,
                 void* stream) {
  MTAM_TRY(check_batch(h, batch));
  cudaStream_t st = (cudaStream_t)stream;
  MTAM_TRY(fwd_dispatch(h, batch, batch->B, scalars_out, true, st));
  if (loss_origin_out)
    MTAM_CUDA_CHECK(cudaMemcpyAsync(loss_origin_out, h->ws.loss_origin, (size_t)batch->B * sizeof(float),
                                    cudaMemcpyDeviceToDevice, st));
  if (pred_out)
    MTAM_CUDA_CHECK(cudaMemcpyAsync(pred_out, h->ws.pred, (size_t)batch->B * h->cfg.D * sizeof(float),
                                    cudaMemcpyDeviceToDevice, st));
  return 0;
}

// ---- row-sharded item table ("scalable training variant") ----------------------------------------------
static int rows_batch(mtam_handle h, const mtam_batch* batch, const float* item_rows, mtam_batch* out) {
  MTAM_TRY(check_batch(h, batch));
  if (!is_mtam_family(h->cfg.kind)) return set_error(MTAM_ERR_UNSUPPORTED, "sharded item table: MTAM kinds only");
  if (!item_rows) return set_error(MTAM_ERR_INVALID, "item_rows is null");
  *out = *batch;
  out->item_list = h->ws.iota;
  return 0;
}

int mtam_forward_rows(mtam_handle h, const mtam_batch* batch, const float* item_rows, float* pred_out, float* scalars_out,
                      void* stream) {
  mtam_batch bt;
  MTAM_TRY(rows_batch(h, batch, item_rows, &bt));
  cudaStream_t st = (cudaStream_t)stream;
  const mtam_config& c = h->cfg;
  Workspace& w = h->ws;
  const int64_t T = (int64_t)batch->B * c.L;
  // the sorts of the replicated tables' ids run beside the forward pass, as in mtam_forward_backward
  MTAM_CUDA_CHECK(cudaEventRecord(h->ev_fork, st));
  MTAM_CUDA_CHECK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  MTAM_TRY(sort_by_row(batch->category_list, T, c.category_rows, w.sort_ws[1], w.sort_ws_bytes[1], &h->sk[1], &h->sp[1], h->side));
  MTAM_TRY(sort_by_row(batch->position_list, T, c.position_rows, w.sort_ws[2], w.sort_ws_bytes[2], &h->sk[2], &h->sp[2], h->side));
  MTAM_TRY(sort_by_row(batch->user_id, batch->B, c.user_rows, w.sort_ws[3], w.sort_ws_bytes[3], &h->sk[3], &h->sp[3], h->side));
  MTAM_CUDA_CHECK(cudaEventRecord(h->ev_join, h->side));
  h->sort_pending = true;
  h->item_rows_ext = item_rows;
  h->rows_mode = true;
  const int r = fwd_dispatch(h, &bt, batch->B, scalars_out, false, st);
  h->item_rows_ext = nullptr;
  h->rows_mode = false;
  MTAM_TRY(r);
  if (pred_out)
    MTAM_CUDA_CHECK(cudaMemcpyAsync(pred_out, w.pred, (size_t)batch->B * c.D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

int mtam_backward_rows(mtam_handle h, const mtam_batch* batch, const float* item_rows, const float* dpred,
                       int32_t global_batch, float* norm_sq_sparse, void* stream) {
  mtam_batch bt;
  MTAM_TRY(rows_batch(h, batch, item_rows, &bt));
  if (!dpred || !norm_sq_sparse) return set_error(MTAM_ERR_INVALID, "dpred / norm_sq_sparse is null");
  cudaStream_t st = (cudaStream_t)stream;
  MTAM_CUDA_CHECK(cudaMemcpyAsync(h->ws.dpred, dpred, (size_t)batch->B * h->cfg.D * sizeof(float), cudaMemcpyDeviceToDevice, st));
  h->item_rows_ext = item_rows;
  h->rows_mode = true;
  const int r = bwd_dispatch(h, &bt, global_batch, norm_sq_sparse, st);
  h->item_rows_ext = nullptr;
  h->rows_mode = false;
  MTAM_TRY(r);
  h->last_batch = *batch;
  h->last_B = batch->B;
  h->grads_pending = true;
  return 0;
}

// out[0] += sum of x[i]^2 (the dense pieces of tf.clip_by_global_norm's norm, base_model.py:294)
int mtam_sumsq(const float* x, int64_t n, float* out_accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !out_accumulate || n < 0) return set_error(MTAM_ERR_INVALID, "mtam_sumsq: bad argument");
  if (n == 0) return 0;
  if ((uintptr_t)x % 16) return set_error(MTAM_ERR_INVALID, "mtam_sumsq: x must be 16-byte aligned");
  const int np = sumsq_num_partials(n);
  if (!workspace || workspace_bytes < (size_t)np * sizeof(float)) return set_error(MTAM_ERR_WORKSPACE, "mtam_sumsq: workspace too small");
  int got = 0;
  MTAM_TRY(sumsq_partials(x, n, (float*)workspace, &got, (cudaStream_t)stream));
  return finalize_sum((float*)workspace, got, 1.0f, out_accumulate, 1, (cudaStream_t)stream);
}
size_t mtam_sumsq_workspace(int64_t n) { return (size_t)sumsq_num_partials(n) * sizeof(float) + 256; }

int mtam_forward_backward(mtam_handle h, const mtam_batch* batch, int32_t global_batch, float* scalars_out,
                          float* norm_sq_sparse, void* stream) {
  MTAM_TRY(check_batch(h, batch));
  if (!norm_sq_sparse) return set_error(MTAM_ERR_INVALID, "norm_sq_sparse is null");
  if (global_batch < batch->B) return set_error(MTAM_ERR_INVALID, "global_batch < local batch");
  cudaStream_t st = (cudaStream_t)stream;
  if (h->cfg.kind != MTAM_KIND_BPRMF) {
    // The scatter-add's sort depends on the batch indices only: run it on the side stream, concurrently
    // with forward/backward (fork/join with events, so it is also captured into a CUDA graph).
    const mtam_config& c = h->cfg;
    Workspace& w = h->ws;
    const int64_t T = (int64_t)batch->B * c.L;
    MTAM_CUDA_CHECK(cudaEventRecord(h->ev_fork, st));
    MTAM_CUDA_CHECK(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
    MTAM_TRY(sort_by_row(batch->item_list, T, c.item_rows, w.sort_ws[0], w.sort_ws_bytes[0], &h->sk[0], &h->sp[0], h->side));
    MTAM_TRY(sort_by_row(batch->category_list, T, c.category_rows, w.sort_ws[1], w.sort_ws_bytes[1], &h->sk[1], &h->sp[1], h->side));
    MTAM_TRY(sort_by_row(batch->position_list, T, c.position_rows, w.sort_ws[2], w.sort_ws_bytes[2], &h->sk[2], &h->sp[2], h->side));
    if (c.kind != MTAM_KIND_PISTREC)
      MTAM_TRY(sort_by_row(batch->user_id, batch->B, c.user_rows, w.sort_ws[3], w.sort_ws_bytes[3], &h->sk[3], &h->sp[3], h->side));
    MTAM_CUDA_CHECK(cudaEventRecord(h->ev_join, h->side));
    h->sort_pending = true;
  }
  MTAM_TRY(fwd_dispatch(h, batch, global_batch, scalars_out, true, st));
  MTAM_TRY(bwd_dispatch(h, batch, global_batch, norm_sq_sparse, st));
  h->last_batch = *batch;
  h->last_B = batch->B;
  h->grads_pending = true;
  return 0;
}

int mtam_finish_grads(mtam_handle h, float* norm_sq, int32_t scatter_local, void* stream) {
  if (!h || !norm_sq) return set_error(MTAM_ERR_INVALID, "null argument");
  if (!h->grads_pending) return set_error(MTAM_ERR_INVALID, "mtam_finish_grads without a pending forward_backward");
  cudaStream_t st = (cudaStream_t)stream;
  const mtam_config& c = h->cfg;
  const Layout& l = h->lay;
  Workspace& w = h->ws;
  const mtam_batch& bt = h->last_batch;
  const int B = h->last_B, D = c.D;
  const int64_t T = (int64_t)B * c.L;
  // dense pieces: everything from the item table to the end of the arena (the sparse-only table
  // regions in front are excluded; their pieces were counted un-deduplicated by forward_backward)
  int np = 0;
  MTAM_TRY(sumsq_partials(h->grads + l.dense_begin, (int64_t)(l.total - l.dense_begin), w.norm_partial, &np, st));
  MTAM_TRY(finalize_sum(w.norm_partial, np, 1.0f, norm_sq, 1, st));
  phase(h, MTAM_PH_SCATTER, st);
  if (h->sort_pending) {   // always re-join the side stream (also closes the fork inside a graph capture)
    MTAM_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_join, 0));
    h->sort_pending = false;
  }
  if (!scatter_local) return 0;
  if (c.kind == MTAM_KIND_BPRMF) return bpr_scatter(h, st);
  float* G = h->grads;
  // Four independent reductions into four tables.  At a few ten thousand rows each is a single partial wave of CTAs (one per
  // SM: 192 KB of shared memory), so they run side by side on the side streams -- all idle here: their work of this step
  // was joined before the norm -- instead of one after the other.  (Not while profiling: the phase is timed on `st`.)
  cudaStream_t s1 = st, s2 = st, s3 = st;
  const bool aside = !h->prof;
  if (aside) {
    s1 = h->side; s2 = h->side2; s3 = h->side3;
    MTAM_CUDA_CHECK(cudaEventRecord(h->ev_sc_fork, st));
    MTAM_CUDA_CHECK(cudaStreamWaitEvent(s1, h->ev_sc_fork, 0));
    MTAM_CUDA_CHECK(cudaStreamWaitEvent(s2, h->ev_sc_fork, 0));
    MTAM_CUDA_CHECK(cudaStreamWaitEvent(s3, h->ev_sc_fork, 0));
  }
  MTAM_TRY(seg_reduce_sorted(h->sk[0], h->sp[0], w.dE2, 2 * D, T, D, G + l.item, D, w.seg_ws, w.seg_ws_bytes, st));
  MTAM_TRY(seg_reduce_sorted(h->sk[1], h->sp[1], w.dE2 + D, 2 * D, T, D, G + l.cat, D, aside ? w.seg_ws_side[0] : w.seg_ws,
                             w.seg_ws_bytes, s1));
  MTAM_TRY(seg_reduce_sorted(h->sk[2], h->sp[2], w.dEp, D, T, D, G + l.pos, D, aside ? w.seg_ws_side[1] : w.seg_ws,
                             w.seg_ws_bytes, s2));
  if (c.kind != MTAM_KIND_PISTREC)
    MTAM_TRY(seg_reduce_sorted(h->sk[3], h->sp[3], w.dEu, D, B, D, G + l.user, D, aside ? w.seg_ws_side[2] : w.seg_ws,
                               w.seg_ws_bytes, s3));
  if (aside) {
    MTAM_CUDA_CHECK(cudaEventRecord(h->ev_sc[0], s1));
    MTAM_CUDA_CHECK(cudaEventRecord(h->ev_sc[1], s2));
    MTAM_CUDA_CHECK(cudaEventRecord(h->ev_sc[2], s3));
    for (int i = 0; i < 3; ++i) MTAM_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_sc[i], 0));
  }
  (void)bt;
  return 0;
}

int mtam_set_item_grad_event(mtam_handle h, void* cuda_event) {
  if (!h) return set_error(MTAM_ERR_INVALID, "null handle");
  h->ev_ce_done = (cudaEvent_t)cuda_event;
  return 0;
}

int mtam_scatter_sparse_into(mtam_handle h, float* item_dst, float* category_dst, float* position_dst, float* user_dst,
                             void* stream) {
  if (!h) return set_error(MTAM_ERR_INVALID, "null handle");
  if (!h->grads_pending) return set_error(MTAM_ERR_INVALID, "mtam_scatter_sparse_into without a pending forward_backward");
  if (h->cfg.kind == MTAM_KIND_BPRMF) return set_error(MTAM_ERR_UNSUPPORTED, "mtam_scatter_sparse_into: not for BPR-MF");
  cudaStream_t st = (cudaStream_t)stream;
  const mtam_config& c = h->cfg;
  Workspace& w = h->ws;
  const int B = h->last_B, D = c.D;
  const int64_t T = (int64_t)B * c.L;
  if (h->sort_pending) {
    MTAM_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_join, 0));
    h->sort_pending = false;
  }
  // side by side on the library's side streams (idle here: joined at the end of the backward pass), as in mtam_finish_grads
  cudaStream_t ss[3] = {h->side, h->side2, h->side3};
  MTAM_CUDA_CHECK(cudaEventRecord(h->ev_sc_fork, st));
  for (int i = 0; i < 3; ++i) MTAM_CUDA_CHECK(cudaStreamWaitEvent(ss[i], h->ev_sc_fork, 0));
  if (item_dst) MTAM_TRY(seg_reduce_sorted(h->sk[0], h->sp[0], w.dE2, 2 * D, T, D, item_dst, D, w.seg_ws, w.seg_ws_bytes, st));
  if (category_dst)
    MTAM_TRY(seg_reduce_sorted(h->sk[1], h->sp[1], w.dE2 + D, 2 * D, T, D, category_dst, D, w.seg_ws_side[0], w.seg_ws_bytes, ss[0]));
  if (position_dst)
    MTAM_TRY(seg_reduce_sorted(h->sk[2], h->sp[2], w.dEp, D, T, D, position_dst, D, w.seg_ws_side[1], w.seg_ws_bytes, ss[1]));
  if (user_dst && c.kind != MTAM_KIND_PISTREC)
    MTAM_TRY(seg_reduce_sorted(h->sk[3], h->sp[3], w.dEu, D, B, D, user_dst, D, w.seg_ws_side[2], w.seg_ws_bytes, ss[2]));
  for (int i = 0; i < 3; ++i) {
    MTAM_CUDA_CHECK(cudaEventRecord(h->ev_sc[i], ss[i]));
    MTAM_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_sc[i], 0));
  }
  return 0;
}

int mtam_apply_begin(mtam_handle h, double lr, const float* norm_sq, float* scalars_out, void* stream) {
  if (!h || !norm_sq) return set_error(MTAM_ERR_INVALID, "null argument");
  if (!h->grads_pending) return set_error(MTAM_ERR_INVALID, "mtam_apply without pending gradients");
  cudaStream_t st = (cudaStream_t)stream;
  float* ds = h->ws.dev_scalars;
  MTAM_TRY(clip_scale(norm_sq, h->cfg.clip, ds + MTAM_S_GLOBAL_NORM, ds + MTAM_S_CLIP_SCALE, st));
  phase(h, MTAM_PH_ADAM, st);
  if (!h->step_prepared) MTAM_TRY(mtam_prepare_step(h, lr, stream));
  h->step_prepared = false;
  if (scalars_out)
    MTAM_CUDA_CHECK(cudaMemcpyAsync(scalars_out + MTAM_S_GLOBAL_NORM, ds + MTAM_S_GLOBAL_NORM, 2 * sizeof(float),
                                    cudaMemcpyDeviceToDevice, st));
  return 0;
}

int mtam_apply_range(mtam_handle h, uint64_t float_begin, uint64_t float_end, void* stream) {
  if (!h) return set_error(MTAM_ERR_INVALID, "null handle");
  if (!h->grads_pending) return set_error(MTAM_ERR_INVALID, "mtam_apply_range without pending gradients");
  const mtam_config& c = h->cfg;
  if (float_end > h->lay.total || float_begin > float_end || (float_begin % 4) || (float_end % 4))
    return set_error(MTAM_ERR_INVALID, "mtam_apply_range: [%llu, %llu) is not a 4-aligned range of the arena",
                     (unsigned long long)float_begin, (unsigned long long)float_end);
  if (float_begin == float_end) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  float* ds = h->ws.dev_scalars;
  const int64_t n = (int64_t)(float_end - float_begin);
  if (c.optimizer == MTAM_OPT_SGD)
    return sgd_apply(h->params + float_begin, h->grads + float_begin, n, ds + MTAM_S_CLIP_SCALE, ds + 9, st);
  return adam_apply(h->params + float_begin, h->m + float_begin, h->v + float_begin, h->grads + float_begin, n,
                    ds + MTAM_S_CLIP_SCALE, ds + 9, c.beta1, c.beta2, c.eps, st);
}

int mtam_apply_end(mtam_handle h, void* stream) {
  if (!h) return set_error(MTAM_ERR_INVALID, "null handle");
  if (!h->grads_pending) return set_error(MTAM_ERR_INVALID, "mtam_apply_end without pending gradients");
  cudaStream_t st = (cudaStream_t)stream;
  const mtam_config& c = h->cfg;
  const Layout& l = h->lay;
  // restore the invariant "sparse-only table regions of the grads arena are zero"
  MTAM_CUDA_CHECK(cudaMemsetAsync(h->grads + l.cat, 0, (l.dense_begin - l.cat) * sizeof(float), st));
  if (c.kind != MTAM_KIND_PISTREC) MTAM_TRY(zero_rows(h->grads + l.user, h->last_batch.user_id, h->last_B, c.D, st));
  h->grads_pending = false;
  phase(h, MTAM_PHASE_COUNT, st);
  return 0;
}

int mtam_apply(mtam_handle h, double lr, const float* norm_sq, float* scalars_out, void* stream) {
  MTAM_TRY(mtam_apply_begin(h, lr, norm_sq, scalars_out, stream));
  MTAM_TRY(mtam_apply_range(h, 0, h->lay.total, stream));
  return mtam_apply_end(h, stream);
}

int mtam_eval_topk(mtam_handle h, const mtam_batch* batch, int32_t k, int32_t* idx_out, float* score_out, void* stream) {
  MTAM_TRY(check_batch(h, batch));
  if (!idx_out) return set_error(MTAM_ERR_INVALID, "idx_out is null");
  cudaStream_t st = (cudaStream_t)stream;
  MTAM_TRY(fwd_dispatch(h, batch, batch->B, nullptr, false, st));
  return score_topk(h->cfg.gemm_mode, h->ws.pred, batch->B, h->cfg.D, h->params + h->lay.item, 0, h->cfg.item_rows, k, idx_out, score_out,
                    h->ws.topk_ws, h->ws.topk_ws_bytes, st);
}

int mtam_train_step(mtam_handle h, const mtam_batch* batch, double lr, float* scalars_out, void* stream) {
  MTAM_TRY(check_batch(h, batch));
  cudaStream_t st = (cudaStream_t)stream;
  const mtam_config& c = h->cfg;
  const Layout& l = h->lay;
  float* nsq = h->ws.dev_scalars + 8;
  float* ds = h->ws.dev_scalars;
  // The user table's rows that this batch does not name have a zero gradient: their Adam update needs lr_t only and
  // starts now, on its own low-priority stream beside the forward and backward pass (optim.cu).  (MTAM kinds: the
  // forward pass reads the batch's user rows only.  Not while profiling: the phase times would no longer add up.)
  const bool early = c.optimizer == MTAM_OPT_ADAM && is_mtam_family(c.kind) && !h->prof;
  if (early) {
    if (!h->step_prepared) MTAM_TRY(mtam_prepare_step(h, lr, stream));      // lr_t is on the device from here on
    MTAM_TRY(mark_rows(batch->user_id, batch->B, h->ws.user_marks, st));
    MTAM_CUDA_CHECK(cudaEventRecord(h->ev_fork4, st));
    MTAM_CUDA_CHECK(cudaStreamWaitEvent(h->side4, h->ev_fork4, 0));
    MTAM_TRY(adam_rows_nograd(h->params + l.user, h->m + l.user, h->v + l.user, c.user_rows, c.D, h->ws.user_marks, ds + 9,
                              c.beta1, c.beta2, c.eps, h->side4));
    MTAM_CUDA_CHECK(cudaEventRecord(h->ev_join4, h->side4));
  }
  MTAM_CUDA_CHECK(cudaMemsetAsync(nsq, 0, sizeof(float), st));
  MTAM_TRY(mtam_forward_backward(h, batch, batch->B, scalars_out, nsq, stream));
  MTAM_TRY(mtam_finish_grads(h, nsq, 1, stream));
  if (!early) return mtam_apply(h, lr, nsq, scalars_out, stream);
  MTAM_TRY(mtam_apply_begin(h, lr, nsq, scalars_out, stream));
  MTAM_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_join4, 0));
  // the batch's user rows (sorted ids: the scatter-add's sort), then everything behind the user table
  MTAM_TRY(adam_rows_listed(h->params + l.user, h->m + l.user, h->v + l.user, h->grads + l.user, h->sk[3], batch->B, c.D,
                            ds + MTAM_S_CLIP_SCALE, ds + 9, c.beta1, c.beta2, c.eps, st));
  MTAM_TRY(mtam_apply_range(h, l.cat, l.total, stream));
  MTAM_TRY(clear_marks(batch->user_id, batch->B, h->ws.user_marks, st));
  return mtam_apply_end(h, stream);
}

}  // extern "C"
