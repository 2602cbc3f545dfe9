// usage: ce_trace.bin [0 = forward | 1 = dpred pass | 2 = dTable pass] [terms: 3 | 1]
// Developer probe: pipeline timeline of the tensor-core CE kernels (events of CTA (0,0), cycles relative to the first).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DMTAM_CE_TRACE \
//   -I mtamrecommender_b200/csrc tools/ce_trace.cu mtamrecommender_b200/csrc/{ce.cu,util.cu} -o tools/ce_trace.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../mtamrecommender_b200/csrc/ce_tc.cu"

int main(int argc, char** argv) {
  using namespace mtam;
  const int B = 1024, V = 100003, D = 64;
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int terms = argc > 2 ? atoi(argv[2]) : 3;
  std::vector<float> hp((size_t)B * D), ht((size_t)V * D);
  for (auto& x : hp) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
  for (auto& x : ht) x = (rand() / (float)RAND_MAX - 0.5f) * 0.6f;
  std::vector<int> htg(B);
  for (auto& x : htg) x = rand() % V;
  float *pred, *table, *tlogit, *lse, *lo, *bp, *dT, *dp;
  int* tg;
  void* ws;
  size_t wsb = ce_workspace_bytes(B, D, V);
  cudaMalloc(&pred, hp.size() * 4); cudaMalloc(&table, ht.size() * 4); cudaMalloc(&tg, B * 4);
  cudaMalloc(&tlogit, B * 4); cudaMalloc(&lse, B * 4); cudaMalloc(&lo, B * 4); cudaMalloc(&bp, 4096);
  cudaMalloc(&dT, ht.size() * 4); cudaMalloc(&dp, hp.size() * 4); cudaMalloc(&ws, wsb);
  cudaMemcpy(pred, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(table, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(tg, htg.data(), B * 4, cudaMemcpyHostToDevice);
  int np = 0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    if (mode == 0) ce_forward_tc(D, pred, table, tg, B, V, ws, tlogit, lse, lo, bp, &np, 0, terms);
    else {
      if (rep == 0) ce_forward_tc(D, pred, table, tg, B, V, ws, tlogit, lse, lo, bp, &np, 0);
      ce_backward_tc(D, pred, table, tg, lse, B, V, 1.f / B, ws, dT, dp, 0, mode == 2 ? 2 : (mode == 1 ? 1 : 3), terms);
    }
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("rep %d: %.1f us  %s\n", rep, ms * 1e3, cudaGetErrorString(e));
  }
  long long h[16 * 64];
  cudaMemcpyFromSymbol(h, g_ce_trace, sizeof(h));
  long long t0 = h[0];
  const char* names[] = {"prod xk_full", "mma  S wait ok", "mma  S issued", "epi0 s_full ok", "epi0 arrived", "mma  g_full ok",
                         "mma  PV issued", "prod xm_full", "epi1 s_full ok", "epi1 arrived", "", "", "", "", "epi0 S loaded", "epi0 G stored"};
  const int first = mode == 2 ? 0 : 20;     // the dTable pass has 16 tiles per CTA: show them all, from the CTA's start
  const long long base = mode == 2 ? h[10 * 64] : h[0 * 64 + 20];
  printf("%-16s", "tile");
  for (int i = 0; i < (mode == 2 ? 16 : 12); ++i) printf("%8d", i + first);
  printf("\n");
  for (int s = 0; s < 10; ++s) {
    printf("%-16s", names[s]);
    for (int i = first; i < first + (mode == 2 ? 16 : 12); ++i) printf("%8lld", h[s * 64 + i] ? h[s * 64 + i] - base : -1);
    printf("\n");
  }
  printf("CTA set up %lld, Q in TMEM %lld, accumulator complete %lld, stored %lld (cycles from set-up)\n", h[10 * 64] - base,
         h[11 * 64] - base, h[12 * 64] - base, h[13 * 64] - base);
  return 0;
}
