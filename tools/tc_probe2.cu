// Probe of the MN-major descriptor view (B stored [k][n], n contiguous).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../mtamrecommender_b200/csrc/tc_common.cuh"
using namespace mtam::tc;

__global__ void __launch_bounds__(128) probe(float* out, int variant, uint32_t lbo, uint32_t sbo) {
  extern __shared__ uint8_t raw[];
  float* sm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  float* A = sm;            // K-major 128 x 32
  float* B = A + 128 * 32;  // MN-major: 4 blocks (n-blocks of 32) x [32 k-rows x 32 floats]
  __shared__ uint64_t mbar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&mbar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&slot, 128);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tacc = slot;
  for (int i = tid; i < 2 * 128 * 32; i += 128) sm[i] = 0.f;
  __syncthreads();
  // A[m][k]: k = kk only -> (m+1);  B[k][n]: k = kk only -> (n+1)   (kk = variant, 0..7)
  const int kk = variant;
  for (int m = tid; m < 128; m += 128) {
    int chunk = kk >> 2, off = m * 32 + ((chunk ^ (m & 7)) << 2) + (kk & 3);
    A[off] = (float)(m + 1);
  }
  for (int n = tid; n < 128; n += 128) {
    int nb = n >> 5, nn = n & 31, chunk = nn >> 2, row = kk;   // row = k index
    int off = nb * 1024 + row * 32 + (chunk_pos_mn(row, chunk) << 2) + (nn & 3);
    B[off] = (float)(n + 1);
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    uint32_t idesc = idesc_tf32(128, 128, 0, 1);
    uint64_t da = desc_kmajor(smem_u32(A), 0), db = smem_desc(smem_u32(B), lbo, sbo, 1);
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
  float* out;
  cudaMalloc(&out, 128 * 128 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  uint32_t conv[4][2] = {{4096, 512}, {512, 4096}, {4096, 1024}, {4096, 4096}};
  for (int c = 0; c < 4; ++c)
    for (int variant = 0; variant < 8; variant += 3) {
      cudaMemset(out, 0xff, 128 * 128 * 4);
      probe<<<1, 128, 40000>>>(out, variant, conv[c][0], conv[c][1]);
      cudaError_t e = cudaDeviceSynchronize();
      static float h[128 * 128];
      cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) if (h[m * 128 + n] != (float)((m + 1) * (n + 1))) ++bad;
      printf("lbo=%u sbo=%u k=%d: %s bad=%d  D[0][0..3]=%g %g %g %g D[0][32..34]=%g %g %g D[1][5]=%g\n", conv[c][0], conv[c][1], variant,
             cudaGetErrorString(e), bad, h[0], h[1], h[2], h[3], h[32], h[33], h[34], h[128 + 5]);
    }
  return 0;
}
