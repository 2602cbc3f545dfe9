// Gradient global norm (tf.clip_by_global_norm) and TF-style Adam over the flat arenas.
//
// Reference: cal_gradient  Model/base_model.py:290-297; tf.train.AdamOptimizer  :76
//   global_norm = sqrt(sum over gradient tensors of sum(values^2)), IndexedSlices contribute their
//   un-deduplicated values (SURVEY trap T1): the caller passes that part in norm_sq already.
//   Adam (trap T2): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
//   w -= lr_t * m / (sqrt(v) + eps).  TF's sparse apply is non-lazy, i.e. this dense update with
//   g = 0 on untouched rows.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "model_kernels.h"

namespace mtam {

__global__ void __launch_bounds__(256) sumsq_kernel(const float4* __restrict__ x, int64_t n4, const float* __restrict__ tail,
                                                    int ntail, float* __restrict__ partial) {
  __shared__ float red[32];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = __ldg(x + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < ntail) s += tail[threadIdx.x] * tail[threadIdx.x];
  float tot = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

int sumsq_num_partials(int64_t n) { return std::max(1, (int)std::min<int64_t>(cdiv(n / 4 + 1, 256 * 4), kNumSMs * 4)); }

// partial[0..*n_partial) = block sums of x[i]^2, i in [0,n).  x must be 16-byte aligned.
int sumsq_partials(const float* x, int64_t n, float* partial, int* n_partial, cudaStream_t st) {
  int blocks = sumsq_num_partials(n);
  int64_t n4 = n / 4;
  sumsq_kernel<<<blocks, 256, 0, st>>>((const float4*)x, n4, x + n4 * 4, (int)(n - n4 * 4), partial);
  MTAM_LAUNCH_CHECK();
  *n_partial = blocks;
  return 0;
}

// scalars[GLOBAL_NORM] = sqrt(*norm_sq); scalars[CLIP_SCALE] = clip / max(norm, clip)
__global__ void clip_scale_kernel(const float* norm_sq, float clip, float* gn_out, float* scale_out) {
  float gn = sqrtf(norm_sq[0]);
  gn_out[0] = gn;
  scale_out[0] = clip / fmaxf(gn, clip);
}
int clip_scale(const float* norm_sq, float clip, float* gn_out, float* scale_out, cudaStream_t st) {
  clip_scale_kernel<<<1, 1, 0, st>>>(norm_sq, clip, gn_out, scale_out);
  MTAM_LAUNCH_CHECK();
  return 0;
}

__global__ void set_scalar_kernel(float* p, float v) { p[0] = v; }
int set_scalar(float* p, float v, cudaStream_t st) {
  set_scalar_kernel<<<1, 1, 0, st>>>(p, v);
  MTAM_LAUNCH_CHECK();
  return 0;
}

__global__ void iota_kernel(int32_t* p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}
int fill_iota(int32_t* p, int64_t n, cudaStream_t st) {
  if (n <= 0) return 0;
  iota_kernel<<<cdiv(n, 256), 256, 0, st>>>(p, n);
  MTAM_LAUNCH_CHECK();
  return 0;
}

__global__ void set_scalar_u32_kernel(uint32_t* p, uint32_t v) { p[0] = v; }
int set_scalar_u32(uint32_t* p, uint32_t v, cudaStream_t st) {
  set_scalar_u32_kernel<<<1, 1, 0, st>>>(p, v);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// tf.train.GradientDescentOptimizer: w -= lr * clipped gradient (rows without a gradient hold 0 in the arena)
__global__ void __launch_bounds__(256) sgd_kernel(float4* __restrict__ w, const float4* __restrict__ g, int64_t n4,
                                                  const float* __restrict__ scale_p, const float* __restrict__ lr_p) {
  const float a = scale_p[0] * lr_p[0];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 gg = __ldg(g + i), ww = w[i];
    ww.x -= a * gg.x; ww.y -= a * gg.y; ww.z -= a * gg.z; ww.w -= a * gg.w;
    w[i] = ww;
  }
}
int sgd_apply(float* w, const float* g, int64_t n, const float* scale_p, const float* lr, cudaStream_t st) {
  int64_t n4 = n / 4;
  int blocks = std::max(1, (int)std::min<int64_t>(cdiv(n4, 256), kNumSMs * 8));
  sgd_kernel<<<blocks, 256, 0, st>>>((float4*)w, (const float4*)g, n4, scale_p, lr);
  MTAM_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) adam_kernel(float4* __restrict__ w, float4* __restrict__ m, float4* __restrict__ v,
                                                   const float4* __restrict__ g, int64_t n4, const float* __restrict__ scale_p,
                                                   const float* __restrict__ lr_t_p, float b1, float b2, float eps) {
  const float sc = scale_p[0];
  const float lr_t = lr_t_p[0];
  const float ob1 = 1.f - b1, ob2 = 1.f - b2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 gg = __ldg(g + i), mm = m[i], vv = v[i], ww = w[i];
#define ADAM1(c)                                   \
  {                                                \
    float gs = gg.c * sc;                          \
    mm.c = b1 * mm.c + ob1 * gs;                   \
    vv.c = b2 * vv.c + ob2 * gs * gs;              \
    ww.c = ww.c - lr_t * mm.c / (sqrtf(vv.c) + eps); \
  }
    ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
    m[i] = mm; v[i] = vv; w[i] = ww;
  }
}

// ---- the user table's update, split in two (MTAM kinds, Adam) -------------------------------------------------------------
// TF's Adam is not lazy (trap T2): every row of every table moves every step.  But a row of the user table that the
// batch does not name has a ZERO gradient, so its update -- m = b1 m, v = b2 v, w -= lr_t m / (sqrt(v) + eps) --
// depends on neither this step's backward pass nor the clip scale, only on lr_t: it can run from the first moment of
// the step, beside the forward and backward pass (which read only the batch's user rows), and that is 86 % of the
// parameters at cfg3.  The batch's rows are marked in a bitmap, skipped here, and updated with their gradient by
// adam_rows_listed_kernel once the norm is known.  `gzero` is a kernel ARGUMENT that holds 0: the expressions -- and so
// the bits -- are those of adam_kernel with g = 0.
__global__ void mark_rows_kernel(const int32_t* __restrict__ idx, int n, unsigned* __restrict__ marks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicOr(marks + ((unsigned)idx[i] >> 5), 1u << (idx[i] & 31));
}
__global__ void clear_marks_kernel(const int32_t* __restrict__ idx, int n, unsigned* __restrict__ marks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) marks[(unsigned)idx[i] >> 5] = 0u;      // every mark of the word came from this list
}
int mark_rows(const int32_t* idx, int n, unsigned* marks, cudaStream_t st) {
  if (n <= 0) return 0;
  mark_rows_kernel<<<cdiv(n, 256), 256, 0, st>>>(idx, n, marks);
  MTAM_LAUNCH_CHECK();
  return 0;
}
int clear_marks(const int32_t* idx, int n, unsigned* marks, cudaStream_t st) {
  if (n <= 0) return 0;
  clear_marks_kernel<<<cdiv(n, 256), 256, 0, st>>>(idx, n, marks);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// Small CTAs that come and go (128 threads, 4 float4 each): they fill whatever registers and thread slots the step's own
// kernels leave free on an SM and never hold one for long.
constexpr int ANG_THREADS = 128, ANG_UNROLL = 4;
__global__ void __launch_bounds__(ANG_THREADS) adam_rows_nograd_kernel(
    float4* __restrict__ w, float4* __restrict__ m, float4* __restrict__ v, int64_t n4, int lg_vpr,
    const unsigned* __restrict__ marks, const float* __restrict__ lr_t_p, float b1, float b2, float eps, float gzero) {
  const float lr_t = lr_t_p[0];
  const float ob1 = 1.f - b1, ob2 = 1.f - b2;
  const int64_t base = ((int64_t)blockIdx.x * ANG_UNROLL) * ANG_THREADS + threadIdx.x;
  float4 mm[ANG_UNROLL], vv[ANG_UNROLL], ww[ANG_UNROLL];
  bool on[ANG_UNROLL];
#pragma unroll
  for (int u = 0; u < ANG_UNROLL; ++u) {
    const int64_t i = base + (int64_t)u * ANG_THREADS;
    const int64_t row = i >> lg_vpr;
    on[u] = i < n4 && !((__ldg(marks + (row >> 5)) >> (row & 31)) & 1u);
    if (on[u]) { mm[u] = m[i]; vv[u] = v[i]; ww[u] = w[i]; }
  }
#pragma unroll
  for (int u = 0; u < ANG_UNROLL; ++u) {
    if (!on[u]) continue;
    const int64_t i = base + (int64_t)u * ANG_THREADS;
#define ADAM0(c)                                                 \
  {                                                              \
    float gs = gzero;                                            \
    mm[u].c = b1 * mm[u].c + ob1 * gs;                           \
    vv[u].c = b2 * vv[u].c + ob2 * gs * gs;                      \
    ww[u].c = ww[u].c - lr_t * mm[u].c / (sqrtf(vv[u].c) + eps); \
  }
    ADAM0(x) ADAM0(y) ADAM0(z) ADAM0(w)
#undef ADAM0
    m[i] = mm[u]; v[i] = vv[u]; w[i] = ww[u];
  }
}
int adam_rows_nograd(float* w, float* m, float* v, int64_t rows, int D, const unsigned* marks, const float* lr_t, float b1,
                     float b2, float eps, cudaStream_t st) {
  if (rows <= 0) return 0;
  const int vpr = D / 4;
  int lg = 0;
  while ((1 << lg) < vpr) ++lg;
  if ((1 << lg) != vpr) return set_error(-1, "adam_rows_nograd: num_units=%d", D);
  const int64_t n4 = rows * vpr;
  const int64_t blocks = (n4 + ANG_THREADS * ANG_UNROLL - 1) / (ANG_THREADS * ANG_UNROLL);
  adam_rows_nograd_kernel<<<(unsigned)blocks, ANG_THREADS, 0, st>>>((float4*)w, (float4*)m, (float4*)v, n4, lg, marks, lr_t,
                                                                    b1, b2, eps, 0.f);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// the rows the batch names, once each: entry j of the SORTED id list updates its row unless entry j-1 names it too
__global__ void __launch_bounds__(256) adam_rows_listed_kernel(float4* __restrict__ w, float4* __restrict__ m,
                                                               float4* __restrict__ v, const float4* __restrict__ g,
                                                               const int32_t* __restrict__ ids_sorted, int n, int vpr,
                                                               const float* __restrict__ scale_p,
                                                               const float* __restrict__ lr_t_p, float b1, float b2, float eps) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int j = (int)(t / vpr);
  if (j >= n) return;
  const int32_t row = ids_sorted[j];
  if (j > 0 && ids_sorted[j - 1] == row) return;
  const int64_t i = (int64_t)row * vpr + (int)(t % vpr);
  const float sc = scale_p[0], lr_t = lr_t_p[0];
  const float ob1 = 1.f - b1, ob2 = 1.f - b2;
  float4 gg = __ldg(g + i), mm = m[i], vv = v[i], ww = w[i];
#define ADAM1(c)                                   \
  {                                                \
    float gs = gg.c * sc;                          \
    mm.c = b1 * mm.c + ob1 * gs;                   \
    vv.c = b2 * vv.c + ob2 * gs * gs;              \
    ww.c = ww.c - lr_t * mm.c / (sqrtf(vv.c) + eps); \
  }
  ADAM1(x) ADAM1(y) ADAM1(z) ADAM1(w)
#undef ADAM1
  m[i] = mm; v[i] = vv; w[i] = ww;
}
int adam_rows_listed(float* w, float* m, float* v, const float* g, const int32_t* ids_sorted, int n, int D,
                     const float* scale_p, const float* lr_t, float b1, float b2, float eps, cudaStream_t st) {
  if (n <= 0) return 0;
  const int vpr = D / 4;
  adam_rows_listed_kernel<<<cdiv((int64_t)n * vpr, 256), 256, 0, st>>>((float4*)w, (float4*)m, (float4*)v, (const float4*)g,
                                                                       ids_sorted, n, vpr, scale_p, lr_t, b1, b2, eps);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// arenas are padded to a multiple of 4 floats by the planner
int adam_apply(float* w, float* m, float* v, const float* g, int64_t n, const float* scale_p, const float* lr_t, float b1,
               float b2, float eps, cudaStream_t st) {
  int64_t n4 = n / 4;
  int blocks = std::max(1, (int)std::min<int64_t>(cdiv(n4, 256), kNumSMs * 8));
  adam_kernel<<<blocks, 256, 0, st>>>((float4*)w, (float4*)m, (float4*)v, (const float4*)g, n4, scale_p, lr_t, b1, b2, eps);
  MTAM_LAUNCH_CHECK();
  return 0;
}

}  // namespace mtam
