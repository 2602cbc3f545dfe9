// fp32 FFMA GEMM with fused epilogues (bias / ReLU / ReLU-mask / add / accumulate) and a
// deterministic split-K path for the weight-gradient shapes (tiny M,N; K = B*L tokens).
// This is the exact-fp32 contraction path (MTAM_GEMM_FP32); the tcgen05 path lives in tc_gemm.cu.
//
// Reference call sites: tf.layers.dense / tf.matmul in
//   Embedding/Behavior_embedding_time_aware_attention.py:95-103, Model/Modules/time_aware_attention.py:249-253,
//   Model/Modules/time_aware_rnn.py:243-256 and their tf.gradients (Model/base_model.py:292).
#include <algorithm>

#include "common.cuh"
#include "gemm_epi.cuh"
#include "kernels.h"
#include "../../include/mtam.h"

namespace mtam {

constexpr int GM = 64, GN = 64, GK = 16, GT = 256;

struct GemmBatch {
  int S;                 // > 0: blockIdx.z = problem * S + split
  int64_t sA, sB, sC;    // element strides between problems
};

template <int TA, int TB, int VA, int VB>
__global__ void __launch_bounds__(GT) gemm_f32_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                      const float* __restrict__ B, int ldb, float* __restrict__ C,
                                                      int ldc, EpiDev epi, int kchunk, float* __restrict__ partial,
                                                      GemmBatch bt) {
  __shared__ __align__(16) float As[GK][GM + 4];
  __shared__ __align__(16) float Bs[GK][GN + 4];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  int zsplit = blockIdx.z;
  if (bt.S > 0) {   // batched: blockIdx.z = problem * S + split
    const int p = zsplit / bt.S;
    zsplit -= p * bt.S;
    A += p * bt.sA; B += p * bt.sB; C += p * bt.sC;
    if (partial) partial += (int64_t)p * bt.S * M * N;
  }
  const int kbeg = zsplit * kchunk;
  const int kend = min(K, kbeg + kchunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += GK) {
    // ---- A tile -> As[k][m]
    if (TA == 0) {  // A[m][k], k fastest
      int m = t >> 2, k4 = (t & 3) * 4;
      int gm = m0 + m, gk = k0 + k4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gm < M) {
        const float* p = A + (int64_t)gm * lda + gk;
        if (VA && gk + 3 < kend) {
          float4 q = *reinterpret_cast<const float4*>(p);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (gk + i < kend) v[i] = p[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[k4 + i][m] = v[i];
    } else {  // A stored [k][m], m fastest
      int k = t >> 4, m4 = (t & 15) * 4;
      int gk = k0 + k, gm = m0 + m4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gk < kend) {
        const float* p = A + (int64_t)gk * lda + gm;
        if (VA && gm + 3 < M) {
          float4 q = *reinterpret_cast<const float4*>(p);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (gm + i < M) v[i] = p[i];
        }
      }
      *reinterpret_cast<float4*>(&As[k][m4]) = make_float4(v[0], v[1], v[2], v[3]);
    }
    // ---- B tile -> Bs[k][n]
    if (TB == 0) {  // B[k][n], n fastest
      int k = t >> 4, n4 = (t & 15) * 4;
      int gk = k0 + k, gn = n0 + n4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gk < kend) {
        const float* p = B + (int64_t)gk * ldb + gn;
        if (VB && gn + 3 < N) {
          float4 q = *reinterpret_cast<const float4*>(p);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (gn + i < N) v[i] = p[i];
        }
      }
      *reinterpret_cast<float4*>(&Bs[k][n4]) = make_float4(v[0], v[1], v[2], v[3]);
    } else {  // B stored [n][k], k fastest
      int n = t >> 2, k4 = (t & 3) * 4;
      int gn = n0 + n, gk = k0 + k4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gn < N) {
        const float* p = B + (int64_t)gn * ldb + gk;
        if (VB && gk + 3 < kend) {
          float4 q = *reinterpret_cast<const float4*>(p);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (gk + i < kend) v[i] = p[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[k4 + i][n] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      if (partial)
        partial[((int64_t)zsplit * M + m) * N + n] = acc[i][j];
      else
        C[(int64_t)m * ldc + n] = apply_epi(acc[i][j], m, n, epi, C, ldc);
    }
  }
}

// A block owns kRedCols consecutive outputs; its 8 warps take the partials z = w, w+8, ... (4 loads in flight each) and
// warp 0 adds the 8 slice sums in slice order: S/8 dependent steps instead of S, and still one fixed order.
constexpr int kRedCols = 32;
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int S, int M, int N,
                                                            float* __restrict__ C, int ldc, EpiDev epi, int64_t sC) {
  __shared__ float red[8][kRedCols];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t MN = (int64_t)M * N;
  const int64_t i = (int64_t)blockIdx.x * kRedCols + tx;
  partial += (int64_t)blockIdx.y * S * MN;      // batched: one problem per blockIdx.y
  C += (int64_t)blockIdx.y * sC;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < MN) {
    int z = ty;
    for (; z + 24 < S; z += 32) {
      s0 += partial[(int64_t)z * MN + i];
      s1 += partial[(int64_t)(z + 8) * MN + i];
      s2 += partial[(int64_t)(z + 16) * MN + i];
      s3 += partial[(int64_t)(z + 24) * MN + i];
    }
    for (; z < S; z += 8) s0 += partial[(int64_t)z * MN + i];
  }
  red[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && i < MN) {
    float s = red[0][tx];
#pragma unroll
    for (int w = 1; w < 8; ++w) s += red[w][tx];
    const int m = (int)(i / N), n = (int)(i % N);
    C[(int64_t)m * ldc + n] = apply_epi(s, m, n, epi, C, ldc);
  }
}

int splitk_reduce(const float* partial, int S, int M, int N, float* C, int ldc, const EpiDev& epi, cudaStream_t st) {
  int64_t tot = (int64_t)M * N;
  splitk_reduce_kernel<<<cdiv(tot, kRedCols), 256, 0, st>>>(partial, S, M, N, C, ldc, epi, 0);
  MTAM_LAUNCH_CHECK();
  return 0;
}

static int pick_splits(int M, int N, int K) {
  int tiles = cdiv(M, GM) * cdiv(N, GN);
  if (tiles >= kNumSMs || K < 1024) return 1;
  int want = cdiv(2 * kNumSMs, tiles);
  int maxs = std::max(1, K / 256);
  return std::max(1, std::min(want, maxs));
}

size_t gemm_splitk_workspace_bytes(int M, int N, int K) {
  int S = pick_splits(M, N, K);
  return S > 1 ? (size_t)S * M * N * sizeof(float) + 256 : 0;
}

int gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
             float* C, int ldc, const GemmEpilogue& e, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  EpiDev epi{e.bias, e.mask_pos, e.add, e.ld_mask, e.ld_add, e.relu, e.accumulate, e.alpha};
  int S = pick_splits(M, N, K);
  int kchunk = K;
  float* partial = nullptr;
  if (S > 1) {
    kchunk = cdiv(cdiv(K, S), GK) * GK;
    S = cdiv(K, kchunk);
  }
  if (S > 1) {
    size_t need = (size_t)S * M * N * sizeof(float);
    if (!ws || ws_bytes < need) return set_error(MTAM_ERR_WORKSPACE, "gemm split-K workspace %zu < %zu", ws_bytes, need);
    partial = (float*)ws;
  }
  if (K <= 0) kchunk = GK;
  bool va = ((uintptr_t)A % 16 == 0) && (lda % 4 == 0);
  bool vb = ((uintptr_t)B % 16 == 0) && (ldb % 4 == 0);
  // with split-K the chunk starts are multiples of GK=16 so alignment of k offsets is preserved
  dim3 grid(cdiv(N, GN), cdiv(M, GM), S);
#define LAUNCH(TA, TB, VA, VB) \
  gemm_f32_kernel<TA, TB, VA, VB><<<grid, GT, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, epi, kchunk, partial, bt)
#define DISPATCH_V(TA, TB)                     \
  do {                                         \
    if (va && vb) LAUNCH(TA, TB, 1, 1);        \
    else if (va) LAUNCH(TA, TB, 1, 0);         \
    else if (vb) LAUNCH(TA, TB, 0, 1);         \
    else LAUNCH(TA, TB, 0, 0);                 \
  } while (0)
  GemmBatch bt{0, 0, 0, 0};
  if (!transA && !transB) DISPATCH_V(0, 0);
  else if (!transA && transB) DISPATCH_V(0, 1);
  else if (transA && !transB) DISPATCH_V(1, 0);
  else DISPATCH_V(1, 1);
  MTAM_LAUNCH_CHECK();
  if (S > 1) {
    int64_t tot = (int64_t)M * N;
    splitk_reduce_kernel<<<cdiv(tot, kRedCols), 256, 0, st>>>(partial, S, M, N, C, ldc, epi, 0);
    MTAM_LAUNCH_CHECK();
  }
  return 0;
}

// P independent products C_p = A_p^T B_p of the same shape (A_p = A + p*sA stored [K,M], B_p = B + p*sB stored
// [K,N], C_p = C + p*sC), always split along K so that the P small problems fill the machine; partials are summed
// in a fixed order.  Used for the per-hop weight gradients (M = N = num_units, K = batch).
static int batched_splits(int P, int M, int N, int K) {
  int tiles = P * cdiv(M, GM) * cdiv(N, GN);
  int S = std::max(1, std::min(cdiv(2 * kNumSMs, tiles), std::max(1, K / 64)));
  int kchunk = cdiv(cdiv(K, S), GK) * GK;
  return cdiv(K, kchunk);
}
size_t gemm_atb_batched_workspace_bytes(int P, int M, int N, int K) {
  return (size_t)P * batched_splits(P, M, N, K) * M * N * sizeof(float) + 256;
}
int gemm_atb_batched_f32(int P, int M, int N, int K, const float* A, int lda, int64_t sA, const float* B, int ldb,
                         int64_t sB, float* C, int ldc, int64_t sC, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (P <= 0 || M <= 0 || N <= 0) return 0;
  EpiDev epi{nullptr, nullptr, nullptr, 0, 0, 0, 0, 1.f};
  const int S = batched_splits(P, M, N, std::max(K, 1));
  const int kchunk = K > 0 ? cdiv(cdiv(K, S), GK) * GK : GK;
  const size_t need = (size_t)P * S * M * N * sizeof(float);
  if (!ws || ws_bytes < need) return set_error(MTAM_ERR_WORKSPACE, "batched gemm workspace %zu < %zu", ws_bytes, need);
  float* partial = (float*)ws;
  const bool va = ((uintptr_t)A % 16 == 0) && (lda % 4 == 0) && (sA % 4 == 0);
  const bool vb = ((uintptr_t)B % 16 == 0) && (ldb % 4 == 0) && (sB % 4 == 0);
  GemmBatch bt{S, sA, sB, sC};
  dim3 grid(cdiv(N, GN), cdiv(M, GM), P * S);
  DISPATCH_V(1, 0);
  MTAM_LAUNCH_CHECK();
  dim3 rgrid(cdiv((int64_t)M * N, kRedCols), P);
  splitk_reduce_kernel<<<rgrid, 256, 0, st>>>(partial, S, M, N, C, ldc, epi, sC);
  MTAM_LAUNCH_CHECK();
  return 0;
}
#undef DISPATCH_V
#undef LAUNCH

// ---- deterministic column sums: out[n] = sum_m A[m,n] * (Bmul ? Bmul[m,n] : 1) -----------------
// Two stages (per-row-block partials, then a fixed-order sum over the blocks).  The number of row blocks is
// chosen so that the grid covers the machine whatever the shape (tall-skinny [B*L, D] or short-wide).
static int colsum_rows_per_block(int M, int N) {
  int colblocks = cdiv(N, 128);
  int want_rowblocks = std::max(1, cdiv(4 * kNumSMs, colblocks));
  int rows = std::max(16, cdiv(M, want_rowblocks));
  return std::min(rows, 256);
}

__global__ void __launch_bounds__(128) colsum_partial_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
                                      int M, int N, int rows_per_block, float* __restrict__ partial) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int m = m0;
  if (Bm) {
    for (; m + 3 < m1; m += 4) {
      float a0 = A[(int64_t)m * lda + n], a1 = A[(int64_t)(m + 1) * lda + n], a2 = A[(int64_t)(m + 2) * lda + n], a3 = A[(int64_t)(m + 3) * lda + n];
      float b0 = Bm[(int64_t)m * ldb + n], b1 = Bm[(int64_t)(m + 1) * ldb + n], b2 = Bm[(int64_t)(m + 2) * ldb + n], b3 = Bm[(int64_t)(m + 3) * ldb + n];
      s0 = fmaf(a0, b0, s0); s1 = fmaf(a1, b1, s1); s2 = fmaf(a2, b2, s2); s3 = fmaf(a3, b3, s3);
    }
    for (; m < m1; ++m) s0 = fmaf(A[(int64_t)m * lda + n], Bm[(int64_t)m * ldb + n], s0);
  } else {
    for (; m + 3 < m1; m += 4) {
      float a0 = A[(int64_t)m * lda + n], a1 = A[(int64_t)(m + 1) * lda + n], a2 = A[(int64_t)(m + 2) * lda + n], a3 = A[(int64_t)(m + 3) * lda + n];
      s0 += a0; s1 += a1; s2 += a2; s3 += a3;
    }
    for (; m < m1; ++m) s0 += A[(int64_t)m * lda + n];
  }
  partial[(int64_t)blockIdx.y * N + n] = (s0 + s1) + (s2 + s3);
}
// 32 columns per block, the 8 warps take the row blocks p = w, w+8, ...; warp 0 adds the slice sums in slice order
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial, int P, int N,
                                                           float* __restrict__ out, int accumulate) {
  __shared__ float red[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (n < N) {
    int p = ty;
    for (; p + 24 < P; p += 32) {
      s0 += partial[(int64_t)p * N + n];
      s1 += partial[(int64_t)(p + 8) * N + n];
      s2 += partial[(int64_t)(p + 16) * N + n];
      s3 += partial[(int64_t)(p + 24) * N + n];
    }
    for (; p < P; p += 8) s0 += partial[(int64_t)p * N + n];
  }
  red[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && n < N) {
    float s = red[0][tx];
#pragma unroll
    for (int w = 1; w < 8; ++w) s += red[w][tx];
    out[n] = accumulate ? out[n] + s : s;
  }
}
size_t colsum_workspace_bytes(int M, int N) {
  return (size_t)cdiv(std::max(M, 1), colsum_rows_per_block(M, N)) * N * sizeof(float) + 256;
}

int colsum_f32(const float* A, int lda, const float* Bmul, int ldb, int M, int N, float* out, int accumulate,
               void* ws, size_t ws_bytes, cudaStream_t st) {
  if (N <= 0) return 0;
  int rpb = colsum_rows_per_block(M, N);
  int P = std::max(1, cdiv(M, rpb));
  if (ws_bytes < (size_t)P * N * sizeof(float)) return set_error(MTAM_ERR_WORKSPACE, "colsum workspace too small");
  dim3 grid(cdiv(N, 128), P);
  colsum_partial_kernel<<<grid, 128, 0, st>>>(A, lda, Bmul, ldb, M, N, rpb, (float*)ws);
  colsum_final_kernel<<<cdiv(N, 32), 256, 0, st>>>((const float*)ws, P, N, out, accumulate);
  MTAM_LAUNCHES(1);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// ---- several column sums over the same M rows in ONE launch (the per-sequence bias / gain / gate gradients: M = batch,
// a few hundred columns each -- as separate partial+final launch pairs they cost more in launches than in work) ----
// A 1024-thread block owns 32 columns of one job; its 32 warps take the rows m = w, w+32, ... (4 loads in flight), and
// warp 0 adds the 32 slice sums in slice order.
__global__ void __launch_bounds__(1024) colsum_multi_kernel(const ColsumBatch jb, int M) {
  __shared__ float red[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int j = 0;
#pragma unroll
  for (int q = 1; q < kColsumMaxJobs; ++q)
    if (q < jb.n_jobs && (int)blockIdx.x >= jb.first_block[q]) j = q;
  const ColsumJob job = jb.job[j];
  const int n = ((int)blockIdx.x - jb.first_block[j]) * 32 + tx;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (n < job.N) {
    const float* A = job.A + n;
    int m = ty;
    if (job.Bmul) {
      const float* Bm = job.Bmul + n;
      for (; m + 96 < M; m += 128) {
        s0 = fmaf(A[(int64_t)m * job.lda], Bm[(int64_t)m * job.ldb], s0);
        s1 = fmaf(A[(int64_t)(m + 32) * job.lda], Bm[(int64_t)(m + 32) * job.ldb], s1);
        s2 = fmaf(A[(int64_t)(m + 64) * job.lda], Bm[(int64_t)(m + 64) * job.ldb], s2);
        s3 = fmaf(A[(int64_t)(m + 96) * job.lda], Bm[(int64_t)(m + 96) * job.ldb], s3);
      }
      for (; m < M; m += 32) s0 = fmaf(A[(int64_t)m * job.lda], Bm[(int64_t)m * job.ldb], s0);
    } else {
      for (; m + 96 < M; m += 128) {
        s0 += A[(int64_t)m * job.lda];
        s1 += A[(int64_t)(m + 32) * job.lda];
        s2 += A[(int64_t)(m + 64) * job.lda];
        s3 += A[(int64_t)(m + 96) * job.lda];
      }
      for (; m < M; m += 32) s0 += A[(int64_t)m * job.lda];
    }
  }
  red[ty][tx] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (ty == 0 && n < job.N) {
    float s = red[0][tx];
#pragma unroll
    for (int w = 1; w < 32; ++w) s += red[w][tx];
    job.out[n] = s;
  }
}
int colsum_multi_f32(ColsumBatch jb, int M, cudaStream_t st) {
  if (jb.n_jobs <= 0) return 0;
  if (jb.n_jobs > kColsumMaxJobs) return set_error(MTAM_ERR_INVALID, "colsum_multi: %d jobs > %d", jb.n_jobs, kColsumMaxJobs);
  int blocks = 0;
  for (int j = 0; j < jb.n_jobs; ++j) {
    jb.first_block[j] = blocks;
    blocks += cdiv(std::max(jb.job[j].N, 0), 32);
  }
  if (blocks == 0) return 0;
  colsum_multi_kernel<<<blocks, 1024, 0, st>>>(jb, M);
  MTAM_LAUNCH_CHECK();
  return 0;
}

}  // namespace mtam
