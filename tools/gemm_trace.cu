// Developer probe: pipeline timeline of the warp-specialised GEMM (events of CTA 0), kv shape of cfg3.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DMTAM_WS_TRACE \
//   -I mtamrecommender_b200/csrc tools/gemm_trace.cu mtamrecommender_b200/csrc/{gemm.cu,util.cu} -o tools/gemm_trace.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../mtamrecommender_b200/csrc/tc_gemm_ws.cu"

int main(int argc, char** argv) {
  using namespace mtam;
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  int M = 51200, N = 768, K = 64, ta = 0, tb = 0;
  if (which == 1) { M = 64; N = 768; K = 51200; ta = 1; tb = 0; }
  if (which == 2) { M = 51200; N = 64; K = 768; ta = 0; tb = 1; }
  float *A, *B, *C, *bias; void* ws;
  size_t wsb = gemm_ws_splitk_workspace_bytes(M, N, K) + 1024;
  cudaMalloc(&A, (size_t)M * K * 4); cudaMalloc(&B, (size_t)K * N * 4); cudaMalloc(&C, (size_t)M * N * 4);
  cudaMalloc(&bias, N * 4); cudaMalloc(&ws, wsb);
  cudaMemset(A, 0, (size_t)M * K * 4); cudaMemset(B, 0, (size_t)K * N * 4); cudaMemset(bias, 0, N * 4);
  GemmEpilogue e; e.bias = bias; e.relu = 1;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    gemm_tf32x3_ws(ta, tb, M, N, K, A, ta ? M : K, B, tb ? K : N, C, N, e, ws, wsb, 0);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("rep %d: %.1f us %s\n", rep, ms * 1e3, cudaGetErrorString(err));
  }
  long long h[12 * 64];
  cudaMemcpyFromSymbol(h, g_ws_trace, sizeof(h));
  const char* names[] = {"A  empty ok", "A  arrived", "B  arrived", "MMA inputs ok", "MMA issued", "EPI acc_full", "EPI released", "EPI stored", "EPI fast?", "EPI bias ok", "EPI iter0", "EPI loop end"};
  long long t0 = h[0 * 64 + 8];
  printf("%-14s", "chunk/unit");
  for (int i = 8; i < 20; ++i) printf("%8d", i);
  printf("\n");
  for (int s = 0; s < 12; ++s) {
    printf("%-14s", names[s]);
    for (int i = 8; i < 20; ++i) printf("%8lld", h[s * 64 + i] ? h[s * 64 + i] - t0 : -1);
    printf("\n");
  }
  return 0;
}
