// BPR-MF (Model/BPRMF.py:41-59): one random negative item shared by the whole batch.
//   x[i,j] = b_pos[i] - b_neg + <u_j, i_pos_j - i_neg>      ([B,1] + [B] broadcasts to [B,B], :50)
//   loss   = 5e-5 * (l2(u) + l2(i_pos) + l2(i_neg)) - mean(log(sigmoid(x)))   (:52-59)
// All gradients are sparse (IndexedSlices); their values are left in the caller's buffers for the
// deterministic scatter-add, their un-deduplicated squared norm is accumulated for the global-norm clip.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "model_kernels.h"

namespace mtam {

constexpr float kBprReg = 5e-5f;   // hard-coded in the reference (BPRMF.py:57), not FLAGS.regulation_rate

// one warp per sequence j: gathers rows, dot[j], bpos[j]; pred[j] = u_j; l2 partial per block
__global__ void __launch_bounds__(128) bpr_gather_kernel(const float* __restrict__ Tu, const float* __restrict__ Ti,
                                                         const float* __restrict__ Tb, const int32_t* __restrict__ user,
                                                         const int32_t* __restrict__ target, int neg, int B, int D,
                                                         float* __restrict__ U, float* __restrict__ IP,
                                                         float* __restrict__ dot, float* __restrict__ bpos,
                                                         float* __restrict__ l2_partial) {
  __shared__ float red[32];
  int j = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  float ss = 0.f;
  if (j < B) {
    const float* u = Tu + (int64_t)user[j] * D;
    const float* p = Ti + (int64_t)target[j] * D;
    const float* n = Ti + (int64_t)neg * D;
    float d = 0.f;
    for (int k = lane; k < D; k += 32) {
      float uv = u[k], pv = p[k];
      U[(int64_t)j * D + k] = uv;
      IP[(int64_t)j * D + k] = pv;
      d = fmaf(uv, pv - n[k], d);
      ss += uv * uv + pv * pv;
      if (j == 0) ss += n[k] * n[k];
    }
    d = warp_sum(d);
    if (lane == 0) { dot[j] = d; bpos[j] = Tb[target[j]]; }
  }
  float tot = block_sum(ss, red);
  if (threadIdx.x == 0) l2_partial[blockIdx.x] = tot;
}

__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

// block i: rowsum[i] = sum_j c[i,j], loss partial; c = -(1 - sigmoid(x)) / B^2
__global__ void __launch_bounds__(256) bpr_row_kernel(const float* __restrict__ bpos, const float* __restrict__ dot,
                                                      const float* __restrict__ Tb, int neg, int B, float* __restrict__ rowsum,
                                                      float* __restrict__ loss_partial) {
  __shared__ float red[32];
  int i = blockIdx.x;
  float base = bpos[i] - Tb[neg], ls = 0.f, cs = 0.f;
  for (int j = threadIdx.x; j < B; j += 256) {
    float x = base + dot[j];
    ls -= log_sigmoid(x);
    cs -= (1.f - sigmoidf_(x));
  }
  float l = block_sum(ls, red), c = block_sum(cs, red);
  if (threadIdx.x == 0) { loss_partial[i] = l; rowsum[i] = c / ((float)B * (float)B); }
}
// block j: colsum[j] = sum_i c[i,j]
__global__ void __launch_bounds__(256) bpr_col_kernel(const float* __restrict__ bpos, const float* __restrict__ dot,
                                                      const float* __restrict__ Tb, int neg, int B, float* __restrict__ colsum) {
  __shared__ float red[32];
  int j = blockIdx.x;
  float d = dot[j] - Tb[neg], cs = 0.f;
  for (int i = threadIdx.x; i < B; i += 256) cs -= (1.f - sigmoidf_(bpos[i] + d));
  float c = block_sum(cs, red);
  if (threadIdx.x == 0) colsum[j] = c / ((float)B * (float)B);
}

// gradient values of the sparse pieces + their squared norm (per-block partial)
//   dU[j] = colsum[j] (ip - in) + reg u ; dIP[j] = colsum[j] u + reg ip ; dINpart[j] = -colsum[j] u
__global__ void __launch_bounds__(128) bpr_grad_kernel(const float* __restrict__ U, const float* __restrict__ IP,
                                                       const float* __restrict__ Ti, int neg,
                                                       const float* __restrict__ colsum, int B, int D,
                                                       float* __restrict__ dU, float* __restrict__ dIP,
                                                       float* __restrict__ dINpart, float* __restrict__ sq_partial) {
  __shared__ float red[32];
  int j = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  float ss = 0.f;
  if (j < B) {
    const float c = colsum[j];
    const float* n = Ti + (int64_t)neg * D;
    for (int k = lane; k < D; k += 32) {
      float u = U[(int64_t)j * D + k], p = IP[(int64_t)j * D + k];
      float du = fmaf(c, p - n[k], kBprReg * u), dp = fmaf(c, u, kBprReg * p);
      dU[(int64_t)j * D + k] = du;
      dIP[(int64_t)j * D + k] = dp;
      dINpart[(int64_t)j * D + k] = -c * u;
      ss += du * du + dp * dp;
    }
  }
  float tot = block_sum(ss, red);
  if (threadIdx.x == 0) sq_partial[blockIdx.x] = tot;
}

// single block: dIN = reg * in + sum_j dINpart[j] (fixed order); db_neg = -sum_i rowsum[i];
// tail[0] = |dIN|^2 + sum_i rowsum[i]^2 + db_neg^2
__global__ void __launch_bounds__(256) bpr_tail_kernel(const float* __restrict__ dINpart, const float* __restrict__ Ti,
                                                       int neg, const float* __restrict__ rowsum, int B, int D,
                                                       float* __restrict__ dIN, float* __restrict__ dbneg,
                                                       float* __restrict__ tail) {
  __shared__ float red[32];
  float ss = 0.f;
  for (int k = threadIdx.x; k < D; k += 256) {
    float s = 0.f;
    for (int j = 0; j < B; ++j) s += dINpart[(int64_t)j * D + k];
    s = fmaf(kBprReg, Ti[(int64_t)neg * D + k], s);
    dIN[k] = s;
    ss += s * s;
  }
  float rs = 0.f;
  for (int i = threadIdx.x; i < B; i += 256) { float r = rowsum[i]; rs += r; ss += r * r; }
  float tot_r = block_sum(rs, red);
  float tot = block_sum(ss, red);
  if (threadIdx.x == 0) { dbneg[0] = -tot_r; tail[0] = tot + tot_r * tot_r; }
}

int bpr_forward(const BprArgs& a, int* n_l2, int* n_loss, cudaStream_t st) {
  int nb = cdiv(a.B, 4);
  bpr_gather_kernel<<<nb, 128, 0, st>>>(a.Tu, a.Ti, a.Tb, a.user, a.target, a.neg, a.B, a.D, a.U, a.IP, a.dot, a.bpos,
                                        a.l2_partial);
  bpr_row_kernel<<<a.B, 256, 0, st>>>(a.bpos, a.dot, a.Tb, a.neg, a.B, a.rowsum, a.loss_partial);
  MTAM_LAUNCHES(1);
  MTAM_LAUNCH_CHECK();
  *n_l2 = nb;
  *n_loss = a.B;
  return 0;
}
int bpr_backward(const BprArgs& a, int* n_sq, cudaStream_t st) {
  int nb = cdiv(a.B, 4);
  bpr_col_kernel<<<a.B, 256, 0, st>>>(a.bpos, a.dot, a.Tb, a.neg, a.B, a.colsum);
  bpr_grad_kernel<<<nb, 128, 0, st>>>(a.U, a.IP, a.Ti, a.neg, a.colsum, a.B, a.D, a.dU, a.dIP, a.dINpart, a.sq_partial);
  bpr_tail_kernel<<<1, 256, 0, st>>>(a.dINpart, a.Ti, a.neg, a.rowsum, a.B, a.D, a.dIN, a.dbneg, a.sq_partial + nb);
  MTAM_LAUNCHES(2);
  MTAM_LAUNCH_CHECK();
  *n_sq = nb + 1;
  return 0;
}

}  // namespace mtam
