// Shared device/host helpers for libmtam_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace mtam {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- error plumbing: every C-ABI entry point returns a status, text kept per thread -------
std::string& last_error_slot();
int set_error(int code, const char* fmt, ...);

#define MTAM_CUDA_CHECK(expr)                                                                \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return ::mtam::set_error(-2, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,              \
                               cudaGetErrorString(_e));                                      \
  } while (0)

// kernels launched by this library since load (diagnostic: bench.py reports launches per step)
long long& launch_counter();
#define MTAM_LAUNCHES(n) (::mtam::launch_counter() += (n))
#define MTAM_LAUNCH_CHECK()                \
  do {                                     \
    MTAM_LAUNCHES(1);                      \
    MTAM_CUDA_CHECK(cudaGetLastError());   \
  } while (0)

#define MTAM_TRY(expr)              \
  do {                              \
    int _s = (expr);                \
    if (_s != 0) return _s;         \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// bump allocator over a caller-provided workspace
struct Bump {
  char* base;
  size_t cap, off;
  Bump(void* p, size_t c) : base((char*)p), cap(c), off(0) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = (T*)(base ? base + off : nullptr);
    off += n * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};

// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum over aligned groups of W lanes (W power of two <= 32)
template <int W>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Predicated loads that return 0 when `pred` is false.  Written as inline PTX because the C++ form
// `pred ? __ldg(p) : 0.f` compiles to a predicated LDG into a temporary followed by a dependent MOV,
// which stalls on every load and serialises what should be a batch of independent loads.
__device__ __forceinline__ float ld_nc_pred(const float* p, bool pred) {
  float v;
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n mov.f32 %0, 0f00000000;\n @q ld.global.nc.f32 %0, [%1];\n}"
               : "=f"(v) : "l"(p), "r"((int)pred));
  return v;
}
__device__ __forceinline__ float ld_cg_pred(const float* p, bool pred) {
  float v;
  asm volatile("{\n .reg .pred q;\n setp.ne.b32 q, %2, 0;\n mov.f32 %0, 0f00000000;\n @q ld.global.cg.f32 %0, [%1];\n}"
               : "=f"(v) : "l"(p), "r"((int)pred));
  return v;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// block-wide sum for blockDim.x <= 1024 (result valid in all threads)
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[w] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  float r = (threadIdx.x < nw) ? smem32[threadIdx.x] : 0.f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) smem32[0] = r;
  __syncthreads();
  r = smem32[0];
  return r;
}

}  // namespace mtam
