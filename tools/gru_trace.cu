// Developer probe: where one step of the T-GRU forward recurrence (mma.sync kernel) spends its cycles
// (step 10 of CTA 0, thread 0).  Built without -DMTAM_GRU_TRACE it is a plain harness for ncu.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DMTAM_GRU_TRACE \
//   -I mtamrecommender_b200/csrc tools/gru_trace.cu mtamrecommender_b200/csrc/util.cu -o tools/gru_trace.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../mtamrecommender_b200/csrc/gru.cu"

int main(int argc, char** argv) {
  using namespace mtam;
  const int tc = argc > 1 ? atoi(argv[1]) : 1;
  const int B = 1024, L = 50, D = 64;
  const size_t T = (size_t)B * L;
  auto rnd = [](size_t n, float s) { std::vector<float> v(n); for (auto& x : v) x = (rand() / (float)RAND_MAX - 0.5f) * s; return v; };
  auto up = [](const std::vector<float>& h) { float* d; cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice); return d; };
  float *X = up(rnd(T * D, 1.f)), *GX = up(rnd(T * 3 * D, 1.f)), *tl = up(rnd(T, 10.f)), *W = up(rnd(2 * D * 3 * D, 0.2f)),
        *vecs = up(rnd(8 * D, 0.5f));
  std::vector<int> hs(B);
  for (auto& x : hs) x = 2 + rand() % (L - 1);
  int* sl; cudaMalloc(&sl, B * 4); cudaMemcpy(sl, hs.data(), B * 4, cudaMemcpyHostToDevice);
  float *Hs, *RUCT, *RH, *q0;
  cudaMalloc(&Hs, (T + 1) * D * 4); cudaMalloc(&RUCT, T * 4 * D * 4); cudaMalloc(&RH, T * D * 4); cudaMalloc(&q0, B * D * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    gru_forward(D, X, GX, tl, sl, W, vecs, B, L, Hs, RUCT, RH, q0, 0, 0, tc);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("rep %d: %.1f us %s\n", rep, ms * 1e3, cudaGetErrorString(e));
  }
#ifdef MTAM_GRU_TRACE
  long long h[32];
  cudaMemcpyFromSymbol(h, g_gru_trace, sizeof(h));
  const char* nm[] = {"loop top", "next step's loads issued", "gates product done", "gates applied + stored", "sync", "candidate product done",
                      "candidate + state update", "cur = nxt (loads landed)", "sync"};
  for (int i = 1; i < 9; ++i) printf("%-28s +%lld cycles\n", nm[i], h[i] - h[i - 1]);
  printf("step total %lld cycles\n", h[8] - h[0]);
#endif
  return 0;
}
