// Stand-alone probe of the tcgen05 primitives in tc_common.cuh (compiled and run on the GPU box).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../mtamrecommender_b200/csrc/tc_common.cuh"
using namespace mtam::tc;

__global__ void __launch_bounds__(128) probe(float* out, uint32_t* info, int variant) {
  extern __shared__ uint8_t raw[];
  float* sm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  float* A = sm;            // 128 x 32
  float* B = A + 128 * 32;  // 128 x 32
  __shared__ uint64_t mbar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&slot, 128);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tacc = slot;
  if (tid == 0) { info[0] = tacc; info[1] = smem_u32(A); info[2] = smem_u32(B); }
  // A[m][k] = (m+1) if k == 0 else 0 ; B[n][k] = (n+1) if k == 0 else 0   -> D[m][n] = (m+1)(n+1)
  for (int q = tid; q < 128 * 8; q += 128) {
    int row = q >> 3, chunk = q & 7;
    float4 va = make_float4(0, 0, 0, 0), vb = va;
    if (variant == 0) { if (chunk == 0) { va.x = row + 1; vb.x = row + 1; } }
    else { va = make_float4(1, 1, 1, 1); vb = va; }
    int off = row * 32 + ((chunk ^ (row & 7)) << 2);
    *reinterpret_cast<float4*>(A + off) = va;
    *reinterpret_cast<float4*>(B + off) = vb;
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    uint32_t idesc = idesc_tf32(128, 128, 0, 0);
    info[3] = idesc;
    uint64_t da = desc_kmajor(smem_u32(A), 0), db = desc_kmajor(smem_u32(B), 0);
    info[4] = (uint32_t)da; info[5] = (uint32_t)(da >> 32);
    mma_tf32(tacc, da, db, idesc, false);
    mma_commit(&mbar);
  }
  mbar_wait(&mbar, 0);
  tc_fence_after();
  for (int c = 0; c < 128; c += 16) {
    float v[16];
    tmem_ld16(tacc + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 128 + c + j] = v[j];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tacc, 128);
}

int main() {
  float* out; uint32_t* info;
  cudaMalloc(&out, 128 * 128 * 4); cudaMalloc(&info, 64);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(out, 0xff, 128 * 128 * 4);
    probe<<<1, 128, 40000>>>(out, info, variant);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d: %s\n", variant, cudaGetErrorString(e));
    static float h[128 * 128]; uint32_t hi[16];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hi, info, 64, cudaMemcpyDeviceToHost);
    printf(" tmem base %08x smemA %08x smemB %08x idesc %08x desc lo %08x hi %08x\n", hi[0], hi[1], hi[2], hi[3], hi[4], hi[5]);
    printf(" D[0][0..3] = %g %g %g %g ; D[1][0..1] = %g %g ; D[127][127] = %g ; D[64][3] = %g\n", h[0], h[1], h[2], h[3], h[128], h[129], h[127 * 128 + 127], h[64 * 128 + 3]);
  }
  return 0;
}
