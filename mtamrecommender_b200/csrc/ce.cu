// Full-catalogue softmax cross-entropy against the item table, exact-fp32 path (MTAM_GEMM_FP32).
// The [B,V] logits are never materialised: item tiles stay stationary in shared memory while the
// (L2-resident) prediction rows stream past; forward keeps an online (max, sum-exp) per row,
// backward recomputes the logits tile and emits dPred and the dense item-table gradient.
//
// Reference: base_model.output  Model/base_model.py:300-328
//   logits = pred x item_table^T (:316); log_softmax (:317); one-hot pick (:318-321)
//   backward per SURVEY 9.9: dlogits = (softmax - onehot)/B; dpred = dlogits T; dT = dlogits^T pred
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "model_kernels.h"
#include "../../include/mtam.h"

namespace mtam {

constexpr int CT = 64;        // tile edge (rows and items)
constexpr int CPAD = CT + 4;  // k-major tiles: [k][CT+4]

__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  float mn = fmaxf(m, m2);
  if (mn == -INFINITY) { m = mn; s = 0.f; return; }
  s = s * expf(m - mn) + s2 * expf(m2 - mn);
  m = mn;
}

// loads a [CT x D] row-major tile (rows r0.., clipped to nrows) into k-major smem dst[k][CPAD]
template <int D>
__device__ __forceinline__ void load_tile_kmajor(const float* __restrict__ src, int r0, int nrows, float* dst) {
  // consecutive threads take consecutive rows: bank-conflict-free transposed stores
  for (int i = threadIdx.x; i < CT * (D / 4); i += 256) {
    int r = i % CT, c4 = (i / CT) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < nrows) v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)(r0 + r) * D + c4));
    dst[(c4 + 0) * CPAD + r] = v.x;
    dst[(c4 + 1) * CPAD + r] = v.y;
    dst[(c4 + 2) * CPAD + r] = v.z;
    dst[(c4 + 3) * CPAD + r] = v.w;
  }
}
template <int D>
__device__ __forceinline__ void load_tile_rowmajor(const float* __restrict__ src, int r0, int nrows, float* dst) {
  for (int i = threadIdx.x; i < CT * (D / 4); i += 256) {
    int r = i / (D / 4), c4 = (i % (D / 4)) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < nrows) v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)(r0 + r) * D + c4));
    *reinterpret_cast<float4*>(dst + r * (D + 4) + c4) = v;
  }
}

template <int D>
__device__ __forceinline__ void logits_tile(const float* Ps, const float* Ts, int tx, int ty, float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
  for (int k = 0; k < D; ++k) {
    float4 a = *reinterpret_cast<const float4*>(Ps + k * CPAD + ty * 4);
    float4 b = *reinterpret_cast<const float4*>(Ts + k * CPAD + tx * 4);
    float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// ms_partial[cta][B] (float2 = running max, sum-exp over the item tiles this CTA owns)
template <int D>
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ table,
                                                     const int32_t* __restrict__ target, int B, int V,
                                                     float2* __restrict__ ms_partial, float* __restrict__ tlogit) {
  extern __shared__ __align__(16) float sm[];
  float* Ts = sm;              // [D][CPAD]
  float* Ps = Ts + D * CPAD;   // [D][CPAD]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int ntiles = (V + CT - 1) / CT;
  float2* my_ms = ms_partial + (int64_t)blockIdx.x * B;
  bool first = true;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int v0 = tile * CT;
    __syncthreads();
    load_tile_kmajor<D>(table, v0, V, Ts);
    for (int m0 = 0; m0 < B; m0 += CT) {
      __syncthreads();
      load_tile_kmajor<D>(pred, m0, B, Ps);
      __syncthreads();
      float acc[4][4];
      logits_tile<D>(Ps, Ts, tx, ty, acc);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = m0 + ty * 4 + i;
        const int tg = (row < B) ? target[row] : -1;
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int col = v0 + tx * 4 + j;
          if (col >= V) acc[i][j] = -INFINITY;
          else if (col == tg) tlogit[row] = acc[i][j];
          m = fmaxf(m, acc[i][j]);
        }
        float s = 0.f;
        if (m > -INFINITY) {
#pragma unroll
          for (int j = 0; j < 4; ++j) s += expf(acc[i][j] - m);
        }
        // merge across the 16 threads (tx) that share this row: lanes of one half-warp
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
          lse_merge(m, s, m2, s2);
        }
        if (tx == 0 && row < B) {
          if (!first) {
            float2 old = my_ms[row];
            lse_merge(m, s, old.x, old.y);
          }
          my_ms[row] = make_float2(m, s);
        }
      }
    }
    first = false;
  }
  if (first) {  // this CTA owned no tile: neutral element
    for (int r = threadIdx.x; r < B; r += 256) my_ms[r] = make_float2(-INFINITY, 0.f);
  }
}

// lse[b], loss_origin[b] = lse - tlogit, block partial sums of loss_origin.
// One warp per row (lanes across the G per-CTA partials), 32 rows per block.
constexpr int CF_ROWS = 32;
__global__ void __launch_bounds__(256) ce_finalize_kernel(const float2* __restrict__ ms_partial, int G, int B,
                                                          const float* __restrict__ tlogit, float* __restrict__ lse,
                                                          float* __restrict__ loss_origin,
                                                          float* __restrict__ block_partial) {
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float lo_sum = 0.f;
  for (int r = warp; r < CF_ROWS; r += 8) {
    int b = blockIdx.x * CF_ROWS + r;
    if (b >= B) continue;   // warp-uniform
    float m = -INFINITY, s = 0.f;
    for (int g = lane; g < G; g += 32) {
      float2 p = ms_partial[(int64_t)g * B + b];
      lse_merge(m, s, p.x, p.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
      lse_merge(m, s, m2, s2);
    }
    float l = m + logf(s);
    if (lane == 0) {
      lse[b] = l;
      float lo = l - tlogit[b];
      loss_origin[b] = lo;
      lo_sum += lo;
    }
  }
  float tot = block_sum(lo_sum, red);
  if (threadIdx.x == 0) block_partial[blockIdx.x] = tot;
}

// dTable[v,:] = sum_b G[b,v] pred[b,:]  (complete, written once);  dpred_partial[cta][b,:] = sum over the
// CTA's item tiles of G[b,v] T[v,:];  G = (exp(logit - lse) - onehot) * inv_batch
template <int D>
__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ table,
                                                     const int32_t* __restrict__ target, const float* __restrict__ lse,
                                                     int B, int V, float inv_batch, float* __restrict__ dTable,
                                                     float* __restrict__ dpred_partial) {
  constexpr int DT = D / 16;  // d-columns per thread in the gradient mini-GEMMs
  constexpr int RS = D + 4;   // row-major tile stride
  extern __shared__ __align__(16) float sm[];
  float* Ts = sm;                // [D][CPAD]  k-major item tile
  float* Ps = Ts + D * CPAD;     // [D][CPAD]  k-major pred tile
  float* Tr = Ps + D * CPAD;     // [CT][RS]   row-major item tile
  float* Pr = Tr + CT * RS;      // [CT][RS]   row-major pred tile
  float* Gs = Pr + CT * RS;      // [CT rows][CPAD items]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int ntiles = (V + CT - 1) / CT;
  float* my_dp = dpred_partial + (int64_t)blockIdx.x * B * D;
  bool first = true;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int v0 = tile * CT;
    __syncthreads();
    load_tile_kmajor<D>(table, v0, V, Ts);
    load_tile_rowmajor<D>(table, v0, V, Tr);
    float dT[4][DT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < DT; ++j) dT[i][j] = 0.f;
    for (int m0 = 0; m0 < B; m0 += CT) {
      __syncthreads();
      load_tile_kmajor<D>(pred, m0, B, Ps);
      load_tile_rowmajor<D>(pred, m0, B, Pr);
      __syncthreads();
      float acc[4][4];
      logits_tile<D>(Ps, Ts, tx, ty, acc);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = m0 + ty * 4 + i;
        const bool rv = row < B;
        const float l = rv ? lse[row] : 0.f;
        const int tg = rv ? target[row] : -1;
        float gv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int col = v0 + tx * 4 + j;
          float p = (rv && col < V) ? expf(acc[i][j] - l) : 0.f;
          if (col == tg) p -= 1.f;
          gv[j] = p * inv_batch;
        }
        *reinterpret_cast<float4*>(Gs + (ty * 4 + i) * CPAD + tx * 4) = make_float4(gv[0], gv[1], gv[2], gv[3]);
      }
      __syncthreads();
      // dT[item = ty*4+i][d = tx*DT+j] += sum_row G[row][item] * pred[row][d]
#pragma unroll 4
      for (int r = 0; r < CT; ++r) {
        float4 ga = *reinterpret_cast<const float4*>(Gs + r * CPAD + ty * 4);
        float gi[4] = {ga.x, ga.y, ga.z, ga.w};
        float pv[DT];
#pragma unroll
        for (int j = 0; j < DT; ++j) pv[j] = Pr[r * RS + tx * DT + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < DT; ++j) dT[i][j] = fmaf(gi[i], pv[j], dT[i][j]);
      }
      // dP[row = ty*4+i][d = tx*DT+j] = sum_item G[row][item] * T[item][d]
      float dP[4][DT];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < DT; ++j) dP[i][j] = 0.f;
#pragma unroll 4
      for (int c = 0; c < CT; ++c) {
        float gi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) gi[i] = Gs[(ty * 4 + i) * CPAD + c];
        float tv[DT];
#pragma unroll
        for (int j = 0; j < DT; ++j) tv[j] = Tr[c * RS + tx * DT + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < DT; ++j) dP[i][j] = fmaf(gi[i], tv[j], dP[i][j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int row = m0 + ty * 4 + i;
        if (row < B) {
          float* q = my_dp + (int64_t)row * D + tx * DT;
#pragma unroll
          for (int j = 0; j < DT; ++j) q[j] = first ? dP[i][j] : q[j] + dP[i][j];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int item = v0 + ty * 4 + i;
      if (item < V) {
        float* q = dTable + (int64_t)item * D + tx * DT;
#pragma unroll
        for (int j = 0; j < DT; ++j) q[j] = dT[i][j];
      }
    }
    first = false;
  }
  if (first) {
    for (int64_t i = threadIdx.x; i < (int64_t)B * D; i += 256) my_dp[i] = 0.f;
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, int G, int64_t n, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int g = 0; g < G; ++g) s += partial[(int64_t)g * n + i];
  out[i] = s;
}

int ce_grid(int V) { return std::max(1, std::min(cdiv(V, CT), 2 * kNumSMs)); }

// shared with the tensor-core path (ce_tc.cu)
int ce_finalize(const float2* ms_partial, int G, int B, const float* tlogit, float* lse, float* loss_origin,
                float* block_partial, int* n_partial, cudaStream_t st) {
  int nb = cdiv(B, CF_ROWS);
  ce_finalize_kernel<<<nb, 256, 0, st>>>(ms_partial, G, B, tlogit, lse, loss_origin, block_partial);
  MTAM_LAUNCH_CHECK();
  *n_partial = nb;
  return 0;
}
int ce_reduce_partials(const float* partial, int G, int64_t n, float* out, cudaStream_t st) {
  reduce_partials_kernel<<<cdiv(n, 256), 256, 0, st>>>(partial, G, n, out);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// partial-result slots of the workspace: enough for either path
static int ce_ws_ranges(int B, int V) { return std::max(ce_grid(V), 2 * ce_tc_ranges(B, V)); }
size_t ce_ms_region_bytes(int B, int V) { return align_up((size_t)ce_ws_ranges(B, V) * B * sizeof(float2), 256); }

size_t ce_workspace_bytes(int B, int D, int V) {
  int G = ce_ws_ranges(B, V);
  return align_up((size_t)G * B * sizeof(float2), 256) + align_up((size_t)G * B * D * sizeof(float), 256) +
         align_up((size_t)cdiv(B, CF_ROWS) * sizeof(float), 256) + 1024;
}

template <int D>
static int ce_fwd_launch(const float* pred, const float* table, const int32_t* target, int B, int V, void* ws,
                         float* tlogit, float* lse, float* loss_origin, float* block_partial, int* n_partial,
                         cudaStream_t st) {
  int G = ce_grid(V);
  float2* ms = (float2*)ws;
  size_t smem = (size_t)2 * D * CPAD * sizeof(float);
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(ce_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ce_fwd_kernel<D><<<G, 256, smem, st>>>(pred, table, target, B, V, ms, tlogit);
  MTAM_LAUNCH_CHECK();
  return ce_finalize(ms, G, B, tlogit, lse, loss_origin, block_partial, n_partial, st);
}

int ce_forward_f32(int D, const float* pred, const float* table, const int32_t* target, int B, int V, void* ws,
               float* tlogit, float* lse, float* loss_origin, float* block_partial, int* n_partial,
               cudaStream_t st) {
  switch (D) {
    case 32: return ce_fwd_launch<32>(pred, table, target, B, V, ws, tlogit, lse, loss_origin, block_partial, n_partial, st);
    case 64: return ce_fwd_launch<64>(pred, table, target, B, V, ws, tlogit, lse, loss_origin, block_partial, n_partial, st);
    case 128: return ce_fwd_launch<128>(pred, table, target, B, V, ws, tlogit, lse, loss_origin, block_partial, n_partial, st);
  }
  return set_error(-1, "softmax CE: num_units=%d not supported (32, 64, 128)", D);
}

template <int D>
static int ce_bwd_launch(const float* pred, const float* table, const int32_t* target, const float* lse, int B, int V,
                         float inv_batch, void* ws, float* dTable, float* dpred, cudaStream_t st) {
  int G = ce_grid(V);
  float* dpp = (float*)((char*)ws + ce_ms_region_bytes(B, V));
  size_t smem = (size_t)(2 * D * CPAD + 2 * CT * (D + 4) + CT * CPAD) * sizeof(float);
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(ce_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ce_bwd_kernel<D><<<G, 256, smem, st>>>(pred, table, target, lse, B, V, inv_batch, dTable, dpp);
  MTAM_LAUNCH_CHECK();
  return ce_reduce_partials(dpp, G, (int64_t)B * D, dpred, st);
}

int ce_backward_f32(int D, const float* pred, const float* table, const int32_t* target, const float* lse, int B, int V,
                float inv_batch, void* ws, float* dTable, float* dpred, cudaStream_t st) {
  switch (D) {
    case 32: return ce_bwd_launch<32>(pred, table, target, lse, B, V, inv_batch, ws, dTable, dpred, st);
    case 64: return ce_bwd_launch<64>(pred, table, target, lse, B, V, inv_batch, ws, dTable, dpred, st);
    case 128: return ce_bwd_launch<128>(pred, table, target, lse, B, V, inv_batch, ws, dTable, dpred, st);
  }
  return set_error(-1, "softmax CE: num_units=%d not supported (32, 64, 128)", D);
}

int ce_forward(int mode, int D, const float* pred, const float* table, const int32_t* target, int B, int V, void* ws,
               float* tlogit, float* lse, float* loss_origin, float* block_partial, int* n_partial, cudaStream_t st) {
  if (gemm_mode_is_tc(mode) && ce_tc_supported(D))
    return ce_forward_tc(D, pred, table, target, B, V, ws, tlogit, lse, loss_origin, block_partial, n_partial, st,
                         gemm_mode_ce_terms(mode));
  return ce_forward_f32(D, pred, table, target, B, V, ws, tlogit, lse, loss_origin, block_partial, n_partial, st);
}
int ce_backward(int mode, int D, const float* pred, const float* table, const int32_t* target, const float* lse, int B,
                int V, float inv_batch, void* ws, float* dTable, float* dpred, cudaStream_t st) {
  if (gemm_mode_is_tc(mode) && ce_tc_supported(D))
    return ce_backward_tc(D, pred, table, target, lse, B, V, inv_batch, ws, dTable, dpred, st, 3, gemm_mode_ce_terms(mode));
  return ce_backward_f32(D, pred, table, target, lse, B, V, inv_batch, ws, dTable, dpred, st);
}

}  // namespace mtam

// ---- stand-alone softmax cross-entropy over a row range of the catalogue (the sharded log-sum-exp of SURVEY 8e) ----
extern "C" {

// [softmax kernels' partial results | loss_origin scratch [B] | per-block loss partials]
size_t mtam_softmax_ce_workspace(int32_t B, int32_t D, int32_t rows) {
  return mtam::ce_workspace_bytes(B, D, rows) + mtam::align_up((size_t)B * sizeof(float), 256) +
         mtam::align_up((size_t)mtam::cdiv(B, 32) * sizeof(float), 256);
}

int mtam_softmax_ce_forward(int32_t gemm_mode, const float* pred, int32_t B, int32_t D, const float* table, int32_t rows,
                            const int32_t* target, float* lse_out, float* target_logit_out, void* workspace,
                            size_t workspace_bytes, void* stream) {
  using namespace mtam;
  if (!pred || !table || !target || !lse_out || !target_logit_out || !workspace || B < 1 || rows < 1)
    return set_error(MTAM_ERR_INVALID, "mtam_softmax_ce_forward: bad argument");
  if (workspace_bytes < mtam_softmax_ce_workspace(B, D, rows)) return set_error(MTAM_ERR_WORKSPACE, "mtam_softmax_ce_forward: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  // rows whose target lies outside this range keep a target logit of 0 (the range that owns it writes it)
  MTAM_CUDA_CHECK(cudaMemsetAsync(target_logit_out, 0, (size_t)B * sizeof(float), st));
  const size_t base = ce_workspace_bytes(B, D, rows);
  float* loss_origin = (float*)((char*)workspace + base);                // lse - target logit of THIS range: scratch
  float* block_partial = (float*)((char*)workspace + base + align_up((size_t)B * sizeof(float), 256));
  int n_partial = 0;
  return ce_forward(gemm_mode, D, pred, table, target, B, rows, workspace, target_logit_out, lse_out, loss_origin, block_partial,
                    &n_partial, st);
}

int mtam_softmax_ce_backward(int32_t gemm_mode, const float* pred, int32_t B, int32_t D, const float* table, int32_t rows,
                             const int32_t* target, const float* lse, float inv_batch, float* dtable, float* dpred,
                             void* workspace, size_t workspace_bytes, void* stream) {
  using namespace mtam;
  if (!pred || !table || !target || !lse || !dtable || !dpred || !workspace || B < 1 || rows < 1)
    return set_error(MTAM_ERR_INVALID, "mtam_softmax_ce_backward: bad argument");
  if (workspace_bytes < mtam_softmax_ce_workspace(B, D, rows)) return set_error(MTAM_ERR_WORKSPACE, "mtam_softmax_ce_backward: workspace too small");
  return ce_backward(gemm_mode, D, pred, table, target, lse, B, rows, inv_batch, workspace, dtable, dpred, (cudaStream_t)stream);
}

}  // extern "C"

