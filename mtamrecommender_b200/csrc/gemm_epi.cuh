// Fused GEMM epilogue shared by the FFMA and the tcgen05 GEMM kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mtam {

struct EpiDev {
  const float* bias;
  const float* mask_pos;
  const float* add;
  int ld_mask, ld_add, relu, accumulate;
  float alpha;
};

__device__ __forceinline__ float apply_epi(float v, int m, int n, const EpiDev& e, const float* C, int ldc) {
  v *= e.alpha;
  if (e.bias) v += e.bias[n];
  if (e.relu) v = fmaxf(v, 0.f);
  if (e.mask_pos) v = (e.mask_pos[(int64_t)m * e.ld_mask + n] > 0.f) ? v : 0.f;
  if (e.add) v += e.add[(int64_t)m * e.ld_add + n];
  if (e.accumulate) v += C[(int64_t)m * ldc + n];
  return v;
}

// C = epi(sum_z partial[z])  (fixed order: deterministic)
int splitk_reduce(const float* partial, int S, int M, int N, float* C, int ldc, const EpiDev& epi, cudaStream_t st);

}  // namespace mtam
