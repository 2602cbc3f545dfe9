// Behaviour-sequence embedding layer, forward and backward element-wise parts.
//   forward : get_embedding  Embedding/Behavior_embedding_time_aware_attention.py:62-114
//             E2[t] = [ Ti[item_t] | Tc[cat_t] ]  (the tf.concat operand, :95) and the four
//             tf.nn.l2_loss terms of Model/base_model.py:302-307 fused into the same pass.
//   backward: values of the four IndexedSlices that tf.gradients (base_model.py:292) hands to
//             clip_by_global_norm, and their un-deduplicated squared norm (SURVEY trap T1).
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "model_kernels.h"

namespace mtam {

// E2[t, 0:D] = Ti[item[t]], E2[t, D:2D] = Tc[cat[t]];  partial[block] = sum of squares of
// E2 rows, of Tp[pos[t]] rows and (if user != null) of Tu[user[b]] rows handled by this block.
__global__ void __launch_bounds__(256) embed_gather_kernel(
    const float4* __restrict__ Ti, const float4* __restrict__ Tc, const float4* __restrict__ Tp,
    const float4* __restrict__ Tu, const int32_t* __restrict__ item, const int32_t* __restrict__ cat,
    const int32_t* __restrict__ pos, const int32_t* __restrict__ user, int64_t T, int B, int vpr /*D/4*/,
    int include_user, float4* __restrict__ E2, float* __restrict__ partial) {
  __shared__ float red[32];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float ss = 0.f;
  const int64_t total = T * vpr;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
    int64_t t = g / vpr;
    int c = (int)(g % vpr);
    float4 a = __ldg(Ti + (int64_t)__ldg(item + t) * vpr + c);
    float4 b = __ldg(Tc + (int64_t)__ldg(cat + t) * vpr + c);
    float4 p = __ldg(Tp + (int64_t)__ldg(pos + t) * vpr + c);
    E2[t * (2 * vpr) + c] = a;
    E2[t * (2 * vpr) + vpr + c] = b;
    ss += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
    ss += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    ss += p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w;
  }
  if (include_user) {
    const int64_t tu = (int64_t)B * vpr;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < tu; g += stride) {
      float4 u = __ldg(Tu + (int64_t)__ldg(user + g / vpr) * vpr + (int)(g % vpr));
      ss += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
    }
  }
  float tot = block_sum(ss, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

int embed_gather(const float* Ti, const float* Tc, const float* Tp, const float* Tu, const int32_t* item,
                 const int32_t* cat, const int32_t* pos, const int32_t* user, int B, int L, int D,
                 int include_user, float* E2, float* l2_partial, int* n_partial, cudaStream_t st) {
  int64_t T = (int64_t)B * L;
  int vpr = D / 4;
  int blocks = (int)std::min<int64_t>(cdiv(T * vpr, 256), kEmbedMaxBlocks);
  blocks = std::max(blocks, 1);
  embed_gather_kernel<<<blocks, 256, 0, st>>>((const float4*)Ti, (const float4*)Tc, (const float4*)Tp,
                                              (const float4*)Tu, item, cat, pos, user, T, B, vpr, include_user,
                                              (float4*)E2, l2_partial);
  MTAM_LAUNCH_CHECK();
  *n_partial = blocks;
  return 0;
}

// Backward, element-wise tail (after dE2 = dpre * W_emb^T has been written by the GEMM):
//   dE2[t,:]  += reg * E2[t,:]                 (item | category IndexedSlices values)
//   dEp[t,:]   = dX[t,:] + reg * Tp[pos[t],:]  (position values)
//   dEu[b,:]   = reg * Tu[user[b],:]           (user values; only when include_user)
//   partial[block] = sum of squares of everything written.
__global__ void __launch_bounds__(256) embed_bwd_tail_kernel(
    const float4* __restrict__ E2, const float4* __restrict__ dX, const float4* __restrict__ Tp,
    const float4* __restrict__ Tu, const int32_t* __restrict__ pos, const int32_t* __restrict__ user, int64_t T,
    int B, int vpr, float reg, int include_user, float4* __restrict__ dE2, float4* __restrict__ dEp,
    float4* __restrict__ dEu, float* __restrict__ partial) {
  __shared__ float red[32];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float ss = 0.f;
  const int64_t total2 = T * 2 * vpr;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total2; g += stride) {
    float4 e = __ldg(E2 + g), d = dE2[g];
    d.x = fmaf(reg, e.x, d.x); d.y = fmaf(reg, e.y, d.y); d.z = fmaf(reg, e.z, d.z); d.w = fmaf(reg, e.w, d.w);
    dE2[g] = d;
    ss += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
  }
  const int64_t total1 = T * vpr;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total1; g += stride) {
    int64_t t = g / vpr;
    float4 p = __ldg(Tp + (int64_t)__ldg(pos + t) * vpr + (int)(g % vpr)), d = __ldg(dX + g);
    d.x = fmaf(reg, p.x, d.x); d.y = fmaf(reg, p.y, d.y); d.z = fmaf(reg, p.z, d.z); d.w = fmaf(reg, p.w, d.w);
    dEp[g] = d;
    ss += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
  }
  if (include_user) {
    const int64_t tu = (int64_t)B * vpr;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < tu; g += stride) {
      float4 u = __ldg(Tu + (int64_t)__ldg(user + g / vpr) * vpr + (int)(g % vpr));
      u.x *= reg; u.y *= reg; u.z *= reg; u.w *= reg;
      dEu[g] = u;
      ss += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
    }
  }
  float tot = block_sum(ss, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

int embed_bwd_tail(const float* E2, const float* dX, const float* Tp, const float* Tu, const int32_t* pos,
                   const int32_t* user, int B, int L, int D, float reg, int include_user, float* dE2, float* dEp,
                   float* dEu, float* partial, int* n_partial, cudaStream_t st) {
  int64_t T = (int64_t)B * L;
  int vpr = D / 4;
  int blocks = (int)std::min<int64_t>(cdiv(T * 2 * vpr, 256), kEmbedMaxBlocks);
  blocks = std::max(blocks, 1);
  embed_bwd_tail_kernel<<<blocks, 256, 0, st>>>((const float4*)E2, (const float4*)dX, (const float4*)Tp,
                                                (const float4*)Tu, pos, user, T, B, vpr, reg, include_user,
                                                (float4*)dE2, (float4*)dEp, (float4*)dEu, partial);
  MTAM_LAUNCH_CHECK();
  *n_partial = blocks;
  return 0;
}

// out[0] (+)= scale * sum(partial[0..n))  in fixed order, single warp: deterministic.
__global__ void finalize_sum_kernel(const float* __restrict__ partial, int n, float scale, float* out,
                                    int accumulate) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += partial[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) out[0] = accumulate ? out[0] + s * scale : s * scale;
}
int finalize_sum(const float* partial, int n, float scale, float* out, int accumulate, cudaStream_t st) {
  finalize_sum_kernel<<<1, 32, 0, st>>>(partial, n, scale, out, accumulate);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// The step's three loss scalars in one launch (they were four finalize_sum launches in a row on the step's critical
// path): l2 = 0.5 * sum(l2_partial), origin = sum(ce_partial) / global batch, loss = reg * l2 + origin
// (base_model.py:302-323).  Same summation order and the same roundings as the separate launches.
__global__ void __launch_bounds__(64) loss_scalars_kernel(const float* __restrict__ l2_partial, int n_l2,
                                                          const float* __restrict__ ce_partial, int n_ce, float inv_batch,
                                                          float reg, float* __restrict__ l2_out, float* __restrict__ origin_out,
                                                          float* __restrict__ loss_out) {
  __shared__ float s_sum[2];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* p = w ? ce_partial : l2_partial;
  const int n = w ? n_ce : n_l2;
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s += p[i];
  s = warp_sum(s);
  if (lane == 0) s_sum[w] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float l2 = __fmul_rn(s_sum[0], 0.5f), origin = __fmul_rn(s_sum[1], inv_batch);
    l2_out[0] = l2;
    origin_out[0] = origin;
    loss_out[0] = __fadd_rn(__fmul_rn(l2, reg), origin);
  }
}
int loss_scalars_sum(const float* l2_partial, int n_l2, const float* ce_partial, int n_ce, float inv_batch, float reg,
                     float* l2_out, float* origin_out, float* loss_out, cudaStream_t st) {
  loss_scalars_kernel<<<1, 64, 0, st>>>(l2_partial, n_l2, ce_partial, n_ce, inv_batch, reg, l2_out, origin_out, loss_out);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// X[t,:] = R[t,:] + Tp[pos[t],:]   (Behavior_...py:102)
__global__ void __launch_bounds__(256) add_pos_kernel(const float4* __restrict__ R, const float4* __restrict__ Tp,
                                                      const int32_t* __restrict__ pos, int64_t total, int vpr,
                                                      float4* __restrict__ X) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
    float4 r = __ldg(R + g);
    float4 p = __ldg(Tp + (int64_t)__ldg(pos + g / vpr) * vpr + (int)(g % vpr));
    X[g] = make_float4(r.x + p.x, r.y + p.y, r.z + p.z, r.w + p.w);
  }
}
int add_pos(const float* R, const float* Tp, const int32_t* pos, int64_t T, int D, float* X, cudaStream_t st) {
  int vpr = D / 4;
  int64_t total = T * vpr;
  int blocks = std::max(1, (int)std::min<int64_t>(cdiv(total, 256), kNumSMs * 8));
  add_pos_kernel<<<blocks, 256, 0, st>>>((const float4*)R, (const float4*)Tp, pos, total, vpr, (float4*)X);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// out = (mask > 0) ? x : 0     (ReLU backward of dense4emb, Behavior_...py:97-100)
__global__ void __launch_bounds__(256) relu_mask_kernel(const float4* __restrict__ x, const float4* __restrict__ mask,
                                                        int64_t n4, float4* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += stride) {
    float4 a = __ldg(x + g), m = __ldg(mask + g);
    out[g] = make_float4(m.x > 0.f ? a.x : 0.f, m.y > 0.f ? a.y : 0.f, m.z > 0.f ? a.z : 0.f, m.w > 0.f ? a.w : 0.f);
  }
}
int relu_mask(const float* x, const float* mask, int64_t n, float* out, cudaStream_t st) {
  int64_t n4 = n / 4;
  int blocks = std::max(1, (int)std::min<int64_t>(cdiv(n4, 256), kNumSMs * 8));
  relu_mask_kernel<<<blocks, 256, 0, st>>>((const float4*)x, (const float4*)mask, n4, (float4*)out);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// rows[idx[i], 0:D] = 0  for i in [0,n)   (re-zero the rows of a table-gradient region we dirtied)
__global__ void zero_rows_kernel(float4* __restrict__ dst, const int32_t* __restrict__ idx, int64_t n, int vpr) {
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < n * vpr) dst[(int64_t)idx[g / vpr] * vpr + (int)(g % vpr)] = make_float4(0.f, 0.f, 0.f, 0.f);
}
int zero_rows(float* dst, const int32_t* idx, int64_t n, int D, cudaStream_t st) {
  if (n <= 0) return 0;
  int vpr = D / 4;
  zero_rows_kernel<<<cdiv(n * vpr, 256), 256, 0, st>>>((float4*)dst, idx, n, vpr);
  MTAM_LAUNCH_CHECK();
  return 0;
}

}  // namespace mtam
