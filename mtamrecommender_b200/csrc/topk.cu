// Full-catalogue scoring with fused top-k selection, list merge and HR/NDCG.
//
// Reference: base_model.metrics_topK  Model/base_model.py:188-213
//   item_result = tf.matmul(pred, item_table^T)  (:194-195); tf.nn.top_k(item_result, k) for
//   k in {1,5,10,30,50} (:196-200) -- one top-50 pass, the others are prefixes (top_k is sorted
//   descending, equal values -> lower index first);  calculate_topK  (:215-242).
// The [B,V] score matrix is never written: a CTA owns 64 prediction rows and a contiguous chunk
// of items, keeps the rows' running top-k in shared memory and only appends tile candidates that
// beat the current k-th score.
#include <limits.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "model_kernels.h"
#include "../../include/mtam.h"

namespace mtam {

constexpr int KT = 64;          // tile edge
constexpr int KPAD = KT + 4;
constexpr int KMAX = 64;        // largest supported k

__device__ __forceinline__ bool better(float s1, int i1, float s2, int i2) {
  return s1 > s2 || (s1 == s2 && i1 < i2);
}

template <int D>
__global__ void __launch_bounds__(256) score_topk_kernel(const float* __restrict__ pred, const float* __restrict__ table,
                                                         int B, int row_begin, int row_end, int k, int chunk_items,
                                                         float* __restrict__ cand_score, int32_t* __restrict__ cand_idx) {
  extern __shared__ __align__(16) float sm[];
  float* Ts = sm;                       // [D][KPAD]
  float* Ps = Ts + D * KPAD;            // [D][KPAD]
  float* ls = Ps + D * KPAD;            // [KT][KMAX] list scores (sorted, best first)
  int* li = (int*)(ls + KT * KMAX);     // [KT][KMAX] list indices
  float* qs = (float*)(li + KT * KMAX); // [KT][KT] candidate queue scores
  int* qi = (int*)(qs + KT * KT);       // [KT][KT]
  int* qn = qi + KT * KT;               // [KT] queue lengths
  float* thr = (float*)(qn + KT);       // [KT] current k-th score
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * KT;
  const int c0 = row_begin + blockIdx.x * chunk_items;
  const int c1 = min(row_end, c0 + chunk_items);
  for (int i = threadIdx.x; i < KT * KMAX; i += 256) { ls[i] = -INFINITY; li[i] = INT_MAX; }
  if (threadIdx.x < KT) { thr[threadIdx.x] = -INFINITY; qn[threadIdx.x] = 0; }
  // prediction tile, k-major
  for (int i = threadIdx.x; i < KT * (D / 4); i += 256) {
    int r = i % KT, c4 = (i / KT) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + r < B) v = __ldg(reinterpret_cast<const float4*>(pred + (int64_t)(m0 + r) * D + c4));
    Ps[(c4 + 0) * KPAD + r] = v.x; Ps[(c4 + 1) * KPAD + r] = v.y;
    Ps[(c4 + 2) * KPAD + r] = v.z; Ps[(c4 + 3) * KPAD + r] = v.w;
  }
  for (int v0 = c0; v0 < c1; v0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * (D / 4); i += 256) {
      int r = i % KT, c4 = (i / KT) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v0 + r < c1) v = __ldg(reinterpret_cast<const float4*>(table + (int64_t)(v0 + r) * D + c4));
      Ts[(c4 + 0) * KPAD + r] = v.x; Ts[(c4 + 1) * KPAD + r] = v.y;
      Ts[(c4 + 2) * KPAD + r] = v.z; Ts[(c4 + 3) * KPAD + r] = v.w;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
    for (int kk = 0; kk < D; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(Ps + kk * KPAD + ty * 4);
      float4 b = *reinterpret_cast<const float4*>(Ts + kk * KPAD + tx * 4);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const float t = thr[r];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = v0 + tx * 4 + j;
        // items arrive in increasing index inside a CTA, so an equal score can never displace
        if (col < c1 && acc[i][j] > t) {
          int slot = atomicAdd(&qn[r], 1);
          qs[r * KT + slot] = acc[i][j];
          qi[r * KT + slot] = col;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < KT) {
      const int r = threadIdx.x;
      const int n = qn[r];
      float* S = ls + r * KMAX;
      int* I = li + r * KMAX;
      for (int c = 0; c < n; ++c) {
        float s = qs[r * KT + c];
        int id = qi[r * KT + c];
        if (!better(s, id, S[k - 1], I[k - 1])) continue;
        int p = k - 1;
        while (p > 0 && better(s, id, S[p - 1], I[p - 1])) { S[p] = S[p - 1]; I[p] = I[p - 1]; --p; }
        S[p] = s; I[p] = id;
      }
      qn[r] = 0;
      thr[r] = S[k - 1];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KT * k; i += 256) {
    int r = i / k, c = i % k;
    if (m0 + r < B) {
      int64_t o = ((int64_t)blockIdx.x * B + (m0 + r)) * k + c;
      cand_score[o] = ls[r * KMAX + c];
      cand_idx[o] = li[r * KMAX + c];
    }
  }
}

// one CTA per row: pick the best k of n_lists*k candidates under (score desc, index asc)
__global__ void __launch_bounds__(128) merge_topk_kernel(const int32_t* __restrict__ in_idx, const float* __restrict__ in_score,
                                                         int n_lists, int B, int k, int32_t* __restrict__ out_idx,
                                                         float* __restrict__ out_score) {
  extern __shared__ __align__(16) float sm[];
  const int n = n_lists * k;
  float* S = sm;
  int* I = (int*)(sm + n);
  __shared__ float ws[4];
  __shared__ int wi[4], wp[4];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += 128) {
    int l = i / k, c = i % k;
    int64_t o = ((int64_t)l * B + b) * k + c;
    S[i] = in_score[o];
    I[i] = in_idx[o];
  }
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    float bs = -INFINITY;
    int bi = INT_MAX, bp = -1;
    for (int i = threadIdx.x; i < n; i += 128)
      if (better(S[i], I[i], bs, bi) || bp < 0) {
        if (bp < 0 || better(S[i], I[i], bs, bi)) { bs = S[i]; bi = I[i]; bp = i; }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float s2 = __shfl_xor_sync(0xffffffffu, bs, o);
      int i2 = __shfl_xor_sync(0xffffffffu, bi, o), p2 = __shfl_xor_sync(0xffffffffu, bp, o);
      if (p2 >= 0 && (bp < 0 || better(s2, i2, bs, bi) || (s2 == bs && i2 == bi && p2 < bp))) { bs = s2; bi = i2; bp = p2; }
    }
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = bs; wi[threadIdx.x >> 5] = bi; wp[threadIdx.x >> 5] = bp; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 4; ++w)
        if (wp[w] >= 0 && (bp < 0 || better(ws[w], wi[w], bs, bi) || (ws[w] == bs && wi[w] == bi && wp[w] < bp))) {
          bs = ws[w]; bi = wi[w]; bp = wp[w];
        }
      out_idx[(int64_t)b * k + r] = bi;
      if (out_score) out_score[(int64_t)b * k + r] = bs;
      if (bp >= 0) { S[bp] = -INFINITY; I[bp] = INT_MAX; }
    }
    __syncthreads();
  }
}

// HR@k / NDCG@k for k in {1,5,10,30,50} (calculate_topK): single block, fixed order
__global__ void __launch_bounds__(256) hr_ndcg_kernel(const int32_t* __restrict__ topk, int B, int k,
                                                      const int32_t* __restrict__ target, float* __restrict__ out10) {
  __shared__ float red[32];
  const int ks[5] = {1, 5, 10, 30, 50};
  float hit[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, nd[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int b = threadIdx.x; b < B; b += 256) {
    int t = target[b], rank = -1;
    for (int c = 0; c < k; ++c)
      if (topk[(int64_t)b * k + c] == t) { rank = c; break; }
    if (rank >= 0) {
      float g = logf(2.f) / logf((float)rank + 2.f);
#pragma unroll
      for (int q = 0; q < 5; ++q)
        if (rank < ks[q] && ks[q] <= k) { hit[q] += 1.f; nd[q] += g; }
    }
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    float h = block_sum(hit[q], red);
    float n = block_sum(nd[q], red);
    if (threadIdx.x == 0) { out10[2 * q] = h / B; out10[2 * q + 1] = n / B; }
  }
}

static int topk_chunks(int B, int rows) {
  int row_tiles = cdiv(B, KT);
  int want = std::max(1, cdiv(2 * kNumSMs, row_tiles));
  int maxc = std::max(1, cdiv(rows, KT));
  return std::min(want, maxc);
}

size_t score_topk_workspace_bytes(int B, int rows, int k) {
  int nc = topk_chunks(B, rows);
  return align_up((size_t)nc * B * k * sizeof(float), 256) + align_up((size_t)nc * B * k * sizeof(int32_t), 256) + 512;
}

int merge_topk(const int32_t* in_idx, const float* in_score, int n_lists, int B, int k, int32_t* out_idx,
               float* out_score, cudaStream_t st) {
  size_t smem = (size_t)n_lists * k * 8;
  if (smem > 200 * 1024) return set_error(MTAM_ERR_INVALID, "merge_topk: %d lists x k=%d exceed shared memory", n_lists, k);
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
  merge_topk_kernel<<<B, 128, smem, st>>>(in_idx, in_score, n_lists, B, k, out_idx, out_score);
  MTAM_LAUNCH_CHECK();
  return 0;
}

template <int D>
static int score_topk_launch(const float* pred, int B, const float* table, int row_begin, int row_end, int k,
                             int32_t* idx_out, float* score_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int rows = row_end - row_begin;
  const int nc = topk_chunks(B, rows);
  if (ws_bytes < score_topk_workspace_bytes(B, rows, k)) return set_error(MTAM_ERR_WORKSPACE, "score_topk: workspace too small");
  float* cs = (float*)ws;
  int32_t* ci = (int32_t*)((char*)ws + align_up((size_t)nc * B * k * sizeof(float), 256));
  int chunk_items = cdiv(cdiv(rows, nc), KT) * KT;
  size_t smem = (size_t)(2 * D * KPAD + 2 * KT * KMAX + 2 * KT * KT + 2 * KT) * sizeof(float);
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(score_topk_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(nc, cdiv(B, KT));
  score_topk_kernel<D><<<grid, 256, smem, st>>>(pred, table, B, row_begin, row_end, k, chunk_items, cs, ci);
  MTAM_LAUNCH_CHECK();
  return merge_topk(ci, cs, nc, B, k, idx_out, score_out, st);
}

int score_topk(const float* pred, int B, int D, const float* table, int row_begin, int row_end, int k, int32_t* idx_out,
               float* score_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (k < 1 || k > KMAX) return set_error(MTAM_ERR_INVALID, "top-k: k=%d outside [1,%d]", k, KMAX);
  if (row_end - row_begin < k) return set_error(MTAM_ERR_INVALID, "top-k: fewer than k=%d rows to score", k);
  switch (D) {
    case 32: return score_topk_launch<32>(pred, B, table, row_begin, row_end, k, idx_out, score_out, ws, ws_bytes, st);
    case 64: return score_topk_launch<64>(pred, B, table, row_begin, row_end, k, idx_out, score_out, ws, ws_bytes, st);
    case 128: return score_topk_launch<128>(pred, B, table, row_begin, row_end, k, idx_out, score_out, ws, ws_bytes, st);
  }
  return set_error(MTAM_ERR_INVALID, "top-k: num_units=%d not supported (32, 64, 128)", D);
}

}  // namespace mtam

extern "C" {

size_t mtam_score_topk_workspace(int32_t B, int32_t rows, int32_t k) { return mtam::score_topk_workspace_bytes(B, rows, k); }

int mtam_score_topk(const float* pred, int32_t B, int32_t D, const float* item_table, int32_t row_begin, int32_t row_end,
                    int32_t k, int32_t* idx_out, float* score_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!pred || !item_table || !idx_out || !workspace || B < 1)
    return mtam::set_error(MTAM_ERR_INVALID, "mtam_score_topk: null argument");
  return mtam::score_topk(pred, B, D, item_table, row_begin, row_end, k, idx_out, score_out, workspace, workspace_bytes,
                          (cudaStream_t)stream);
}

int mtam_merge_topk(const int32_t* in_idx, const float* in_score, int32_t n_lists, int32_t B, int32_t k, int32_t* out_idx,
                    float* out_score, void* stream) {
  if (!in_idx || !in_score || !out_idx || n_lists < 1 || B < 1 || k < 1)
    return mtam::set_error(MTAM_ERR_INVALID, "mtam_merge_topk: bad argument");
  return mtam::merge_topk(in_idx, in_score, n_lists, B, k, out_idx, out_score, (cudaStream_t)stream);
}

int mtam_hr_ndcg(const int32_t* topk_idx, int32_t B, int32_t k, const int32_t* target, float* out10, void* stream) {
  if (!topk_idx || !target || !out10 || B < 1 || k < 1) return mtam::set_error(MTAM_ERR_INVALID, "mtam_hr_ndcg: bad argument");
  mtam::hr_ndcg_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(topk_idx, B, k, target, out10);
  MTAM_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
