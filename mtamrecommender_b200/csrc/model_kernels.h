// Internal launch API of the model-specific kernels (embed / T-GRU / hops / CE / optimiser).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mtam {

constexpr int kEmbedMaxBlocks = 148 * 8;

// ---- embed.cu -------------------------------------------------------------------------------
int embed_gather(const float* Ti, const float* Tc, const float* Tp, const float* Tu, const int32_t* item,
                 const int32_t* cat, const int32_t* pos, const int32_t* user, int B, int L, int D,
                 int include_user, float* E2, float* l2_partial, int* n_partial, cudaStream_t st);
int embed_bwd_tail(const float* E2, const float* dX, const float* Tp, const float* Tu, const int32_t* pos,
                   const int32_t* user, int B, int L, int D, float reg, int include_user, float* dE2, float* dEp,
                   float* dEu, float* partial, int* n_partial, cudaStream_t st);
int finalize_sum(const float* partial, int n, float scale, float* out, int accumulate, cudaStream_t st);
int loss_scalars_sum(const float* l2_partial, int n_l2, const float* ce_partial, int n_ce, float inv_batch, float reg,
                     float* l2_out, float* origin_out, float* loss_out, cudaStream_t st);
int add_pos(const float* R, const float* Tp, const int32_t* pos, int64_t T, int D, float* X, cudaStream_t st);
int relu_mask(const float* x, const float* mask, int64_t n, float* out, cudaStream_t st);
int zero_rows(float* dst, const int32_t* idx, int64_t n, int D, cudaStream_t st);

// ---- gru.cu ---------------------------------------------------------------------------------
int gru_num_blocks(int B);
int gru_forward(int D, const float* X, const float* GX, const float* timelast, const int32_t* seq_len,
                const float* Wgru, const float* vecs, int B, int L, float* Hs, float* RUCT, float* RH, float* q0,
                cudaStream_t st, int plain = 0, int tensor_cores = 0);
int gru_backward(int D, const float* X, const float* timelast, const int32_t* seq_len, const float* Wgru,
                 const float* vecs, const float* Hs, const float* RUCT, const float* dq0, const float* dOut, int B, int L,
                 float* dGX, float* dX, float* vec_partial, cudaStream_t st, int tensor_cores = 0);

// ---- hops.cu --------------------------------------------------------------------------------
struct HopArgs {
  int B, L, D, H, N;
  const int32_t* seq_len;      // [B] key_length
  const float* target_time;    // [B] t_querys
  const float* time_list;      // [B,L] t_keys
  const float* X;              // [B*L, D] keys (raw)
  const float* KV;             // [B*L, 2*N*D]  relu projections, hop i: K at col 2iD, V at col 2iD+D
  // parameters (arena pointers)
  const float* Wq;             // [N][D][D]
  const float* bq;             // [N][D]
  const float* Wt;             // [N][D][D]  _time_input_w
  const float* gate;           // [N][5][L]  _time_input_w1, _time_input_b1, time_output_w1, time_output_w2, time_output_b
  const float* ln_gamma;       // [N][D]
  const float* ln_beta;        // [N][D]
  const float* lnf_gamma;      // [D]
  const float* lnf_beta;       // [D]
  // saved activations
  float* Qin;                  // [N+1][B][D]  Qin[0] = short_term_intent (input), Qin[i+1] = hop i output
  float* Qr;                   // [N][B][D]
  float* Qt;                   // [N][B][D]
  float* AA;                   // [N][B][H][L]  raw Q_h.K_h
  float* PA;                   // [N][B][H][L]  attention probabilities
  float* ZZ;                   // [N][B][L]
  float* DK;                   // [N][B][L]
  float* GT;                   // [N][B][L]   sigmoid gate
  float* XH;                   // [B][N][D]   normalised pre-affine output
  float* RSTD;                 // [B][N]
  float* XHF;                  // [B][D]      final layer norm
  float* RSTDF;                // [B]
  float* pred;                 // [B][D]
};
struct HopGradArgs {
  const float* dpred;          // [B][D]
  const float* WqT;            // [N][D][D] transposed copies
  const float* WtT;            // [N][D][D]
  float* dX;                   // [B*L][D]   accumulated (+=), pre-zeroed
  float* dKV;                  // [B*L][2ND] pre-zeroed (masked keys stay 0), relu-masked
  float* DOUT;                 // [B][N][D]  d loss / d hop output
  float* DQP;                  // [B][N][D]  dQ pre-activation
  float* DQT;                  // [B][N][D]  d(q Wt)
  float* GB;                   // [B][N*5*L] per-sequence gate-parameter gradients, pre-zeroed
  float* dq0;                  // [B][D]
  float* BKV;                  // [B][N][2D] per-sequence column sums of dKV (CTA-per-sequence path only)
};
size_t hop_smem_bytes(int D, int H, int L, bool bwd);
int hop_forward(const HopArgs& a, cudaStream_t st);
bool hop_backward_writes_all_dkv(int D, int H, int L, int N);
int hop_backward(const HopArgs& a, const HopGradArgs& g, cudaStream_t st);
int transpose_dd(const float* src, float* dst, int D, int count, cudaStream_t st);

// ---- bpr.cu ---------------------------------------------------------------------------------
struct BprArgs {
  int B, D, neg;
  const float *Tu, *Ti, *Tb;           // user table, item table, item_b table
  const int32_t *user, *target;
  float *U, *IP;                       // [B,D] gathered rows (U doubles as predict_behavior_emb)
  float *dot, *bpos, *rowsum, *colsum, *loss_partial;   // [B]
  float *dU, *dIP, *dINpart;           // [B,D] IndexedSlices values
  float *dIN;                          // [D]   negative item row gradient
  float *dbneg;                        // [1]   (d b_pos = rowsum)
  float *l2_partial, *sq_partial;      // [ceil(B/4)+1]
};
int bpr_forward(const BprArgs& a, int* n_l2, int* n_loss, cudaStream_t st);
int bpr_backward(const BprArgs& a, int* n_sq, cudaStream_t st);

// ---- ce.cu ----------------------------------------------------------------------------------
int ce_grid(int V);
size_t ce_workspace_bytes(int B, int D, int V);
// mode = mtam_gemm_mode: exact-fp32 FFMA tiles (ce.cu) or tcgen05 (ce_tc.cu; num_units 32 / 64): 3xTF32, or one tf32
// MMA per product (MTAM_GEMM_TF32)
inline bool gemm_mode_is_tc(int mode) { return mode == 1 || mode == 2; }     // MTAM_GEMM_TF32X3, MTAM_GEMM_TF32
inline int gemm_mode_ce_terms(int mode) { return mode == 2 ? 1 : 3; }
int ce_forward(int mode, int D, const float* pred, const float* table, const int32_t* target, int B, int V, void* ws,
               float* tlogit, float* lse, float* loss_origin, float* block_partial, int* n_partial, cudaStream_t st);
int ce_backward(int mode, int D, const float* pred, const float* table, const int32_t* target, const float* lse, int B,
                int V, float inv_batch, void* ws, float* dTable, float* dpred, cudaStream_t st);
int ce_finalize(const float2* ms_partial, int G, int B, const float* tlogit, float* lse, float* loss_origin,
                float* block_partial, int* n_partial, cudaStream_t st);
int ce_reduce_partials(const float* partial, int G, int64_t n, float* out, cudaStream_t st);
bool ce_tc_supported(int D);
int ce_tc_ranges(int B, int V);
size_t ce_ms_region_bytes(int B, int V);
int ce_forward_tc(int D, const float* pred, const float* table, const int32_t* target, int B, int V, void* ws,
                  float* tlogit, float* lse, float* loss_origin, float* block_partial, int* n_partial, cudaStream_t st,
                  int terms = 3);
int ce_backward_tc(int D, const float* pred, const float* table, const int32_t* target, const float* lse, int B, int V,
                   float inv_batch, void* ws, float* dTable, float* dpred, cudaStream_t st, int parts = 3, int terms = 3);

// top-k filter pass: maxima of the logits over buckets of bs (16 / 64) consecutive table rows, [B][ld]
int ce_bucket_max_tc(int D, const float* pred, int B, const float* table, int V, int bs, float* bmax, int ld,
                     cudaStream_t st);

// ---- optim.cu -------------------------------------------------------------------------------
int sumsq_num_partials(int64_t n);
int sumsq_partials(const float* x, int64_t n, float* partial, int* n_partial, cudaStream_t st);
int clip_scale(const float* norm_sq, float clip, float* gn_out, float* scale_out, cudaStream_t st);
int set_scalar(float* p, float v, cudaStream_t st);
int set_scalar_u32(uint32_t* p, uint32_t v, cudaStream_t st);
int fill_iota(int32_t* p, int64_t n, cudaStream_t st);
int sgd_apply(float* w, const float* g, int64_t n, const float* scale_p, const float* lr, cudaStream_t st);
// the user table's Adam in two parts: rows without a gradient (bitmap of the batch's rows clear) early, the batch's rows late
int mark_rows(const int32_t* idx, int n, unsigned* marks, cudaStream_t st);
int clear_marks(const int32_t* idx, int n, unsigned* marks, cudaStream_t st);
int adam_rows_nograd(float* w, float* m, float* v, int64_t rows, int D, const unsigned* marks, const float* lr_t, float b1,
                     float b2, float eps, cudaStream_t st);
int adam_rows_listed(float* w, float* m, float* v, const float* g, const int32_t* ids_sorted, int n, int D,
                     const float* scale_p, const float* lr_t, float b1, float b2, float eps, cudaStream_t st);
int adam_apply(float* w, float* m, float* v, const float* g, int64_t n, const float* scale_p, const float* lr_t, float b1,
               float b2, float eps, cudaStream_t st);

}  // namespace mtam
