// Self-attention model family (PISTRec / SASRec / TA-SASRec / TiSASRec): layout, workspace and
// forward / backward drivers.  The embedding layer, the softmax CE and the optimiser are shared with
// MTAM and stay in model.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include <string>
#include <vector>

#include "../../include/mtam.h"

namespace mtam {

struct ParamDesc {
  std::string name;
  int rows, cols, ndim, ld;
  size_t off;
  int flags;
};

// offsets (floats) of the self-attention parameter blocks inside the arenas
struct SaLayout {
  size_t W3 = 0;        // [N][D][3D]   dense | dense_1 | dense_2 kernels side by side
  size_t b3 = 0;        // [N][3D]
  size_t Wt = 0;        // [N][D][D]    _time_input_w           (time-aware kinds only)
  size_t gate = 0;      // [N][5][L*L]  _time_input_w1, _time_input_b1, time_output_w1, time_output_w2, time_output_b
  size_t gate_dead = 0; // [N][L*L]     time_output_w3
  size_t lnb = 0, lng = 0;  // [N][D]
  size_t lnfb = 0, lnfg = 0;
};

int sa_build_layout(const mtam_config& c, size_t& offset, std::vector<ParamDesc>& params, SaLayout& sl);
size_t sa_workspace_bytes(const mtam_config& c);

struct SaCtx {
  mtam_config cfg;
  SaLayout sl;
  float* params;
  float* grads;
  const mtam_batch* bt;
  float* X;        // [T,D] embedding-layer output (input of block 0)
  float* dX;       // [T,D] gradient w.r.t. X (output of sa_backward)
  float* pred;     // [B,D]
  float* dpred;    // [B,D]
  float* XHF;      // [B,D]
  float* RSTDF;    // [B]
  void* ws;        // sa workspace
  size_t ws_bytes;
  void* gemm_ws; size_t gemm_ws_bytes;
  void* colsum_ws; size_t colsum_ws_bytes;
  float drop_rate = 0.f;                 // attention dropout of SASRec / TiSASRec (0: off)
  uint32_t drop_seed = 0;
  const uint32_t* drop_counter = nullptr;   // device: the forward-call counter the mask is keyed on
};
int sa_forward(const SaCtx& c, cudaStream_t st);
// tf.contrib.layers.layer_norm (eps 1e-12) of B rows of D floats, and its input gradient (net_utils.py:229-232)
int ln_rows_forward(const float* x, int B, int D, const float* gamma, const float* beta, float* out, float* xhat, float* rstd,
                    cudaStream_t st);
int ln_rows_backward(const float* dout, const float* xhat, const float* rstd, int B, int D, const float* gamma, float* dx,
                     cudaStream_t st);
int sa_backward(const SaCtx& c, cudaStream_t st);

}  // namespace mtam
