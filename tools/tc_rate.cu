// Developer probe: cycles per tcgen05.mma (kind::tf32) for SS / TS operand sources and several N, one CTA per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I mtamrecommender_b200/csrc tools/tc_rate.cu -o /tmp/tc_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace mtam::tc;

template <int N, int TS>
__global__ void __launch_bounds__(128) rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  float* sm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (128 * 32 + 256 * 32); i += 128) sm[i] = 0.001f * (i % 97);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x < 32 && elect_one()) {     // converged warp + elect.sync: no per-MMA election loop (tc_common.cuh)
    const uint32_t a = smem_u32(sm), b = smem_u32(sm + 128 * 32);
    constexpr uint32_t id = idesc_tf32(128, N, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        if (TS) mma_tf32_ts(tm, tm + 256 + ks * 8, desc_kmajor(b, ks), id, true);
        else mma_tf32(tm, desc_kmajor(a, ks), desc_kmajor(b, ks), id, true);
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N, int TS>
void run(const char* name, long long* d) {
  const int iters = 2000;
  size_t smem = (128 * 32 + 256 * 32) * 4 + 1024;
  cudaFuncSetAttribute(rate_kernel<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  rate_kernel<N, TS><<<148, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-28s N=%3d  %7.1f cycles / MMA (M=128,K=8)   %s\n", name, N, (double)h / (iters * 4.0), cudaGetErrorString(e));
}
int main() {
  long long* d;
  cudaMalloc(&d, 8);
  run<64, 0>("SS (A,B from smem)", d);
  run<64, 1>("TS (A from TMEM)", d);
  run<128, 0>("SS (A,B from smem)", d);
  run<128, 1>("TS (A from TMEM)", d);
  run<256, 0>("SS (A,B from smem)", d);
  run<256, 1>("TS (A from TMEM)", d);
  return 0;
}
