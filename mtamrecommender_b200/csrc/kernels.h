// Internal (C++) launch API shared by the .cu files of libmtam_b200.  Every function enqueues on
// `st` and returns 0 / negative mtam_status.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace mtam {

// ---- gather_scatter.cu ----------------------------------------------------------------------
int gather_rows(const float* table, int D, const int32_t* idx, int64_t n, float* out, cudaStream_t st);
size_t scan_tmp_ints(int64_t n);
int exclusive_scan_i32(const int* in, int* out, int64_t n, int* tmp, int* total_out, cudaStream_t st);
// radix_sort.cu: stable one-sweep radix sort of (row id, token index)
size_t sort_workspace_bytes(int64_t n, int table_rows);
int sort_by_row(const int32_t* idx, int64_t n, int table_rows, void* ws, size_t ws_bytes,
                const int32_t** keys_sorted, const int32_t** perm, cudaStream_t st);
// accumulate != 0: dst[key] += sum of the key's rows; == 0: dst[key] = that sum (one plain store per touched row)
size_t seg_reduce_workspace_bytes(int64_t n, int D);
int seg_reduce_sorted(const int32_t* keys_sorted, const int32_t* perm, const float* src, int ld_src, int64_t n,
                      int D, float* dst, int ld_dst, void* ws, size_t ws_bytes, cudaStream_t st, int accumulate = 1);
size_t scatter_add_workspace_bytes(int64_t n, int table_rows, int D);
int scatter_add_rows(float* dst, int table_rows, int D, int ld_dst, const int32_t* idx, const float* rows,
                     int ld_src, int64_t n, void* ws, size_t ws_bytes, int32_t* unique_idx, int32_t* n_unique,
                     cudaStream_t st, int accumulate = 1);

// ---- gemm.cu (fp32 FFMA GEMM with fused epilogues) --------------------------------------------
// C[M,N] = epi( op(A)[M,K] * op(B)[K,N] ), row-major storage with leading dimensions.
//   transA == 0: A is [M,K] (lda >= K);  transA == 1: A is stored [K,M] (lda >= M)
//   transB == 0: B is [K,N] (ldb >= N);  transB == 1: B is stored [N,K] (ldb >= K)
struct GemmEpilogue {
  const float* bias = nullptr;     // [N] added before the activation
  int relu = 0;                    // max(.,0) after bias
  const float* mask_pos = nullptr; // same shape/ld as C: result *= (mask_pos > 0)   (ReLU backward)
  int ld_mask = 0;
  const float* add = nullptr;      // same shape as C (ld_add): result += add  (after activation)
  int ld_add = 0;
  int accumulate = 0;              // C += result instead of C = result
  float alpha = 1.f;
};
size_t gemm_splitk_workspace_bytes(int M, int N, int K);
int gemm_f32(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
             float* C, int ldc, const GemmEpilogue& epi, void* splitk_ws, size_t splitk_ws_bytes,
             cudaStream_t st);
// P small products C_p = A_p^T B_p (A_p stored [K,M], B_p stored [K,N]) in one launch, split along K
size_t gemm_atb_batched_workspace_bytes(int P, int M, int N, int K);
int gemm_atb_batched_f32(int P, int M, int N, int K, const float* A, int lda, int64_t sA, const float* B, int ldb,
                         int64_t sB, float* C, int ldc, int64_t sC, void* ws, size_t ws_bytes, cudaStream_t st);
// tc_gemm.cu: same contract on tcgen05 (3xTF32), and the mode dispatcher
int gemm_tf32x3(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                float* C, int ldc, const GemmEpilogue& epi, void* ws, size_t ws_bytes, cudaStream_t st);
// tc_gemm_ws.cu: warp-specialised persistent variant (A operand staged in TMEM), used whenever its alignment holds
bool gemm_ws_supported(int transA, int M, int N, int K, const float* A, int lda);
size_t gemm_ws_splitk_workspace_bytes(int M, int N, int K);
int gemm_tf32x3_ws(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                   float* C, int ldc, const GemmEpilogue& epi, void* ws, size_t ws_bytes, cudaStream_t st);
int gemm_any(int mode, int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
             float* C, int ldc, const GemmEpilogue& epi, void* ws, size_t ws_bytes, cudaStream_t st);
size_t gemm_any_workspace_bytes(int M, int N, int K);
// out[n] (+)= sum_m A[m,n]   (deterministic two-stage column sum); optional second operand: sum A*Bm
int colsum_f32(const float* A, int lda, const float* Bmul, int ldb, int M, int N, float* out, int accumulate,
               void* ws, size_t ws_bytes, cudaStream_t st);
size_t colsum_workspace_bytes(int M, int N);

// up to kColsumMaxJobs column sums (optionally of A*Bmul) over the same M rows in one launch
constexpr int kColsumMaxJobs = 8;
struct ColsumJob {
  const float* A; int lda;
  const float* Bmul; int ldb;
  int N;
  float* out;
};
struct ColsumBatch { ColsumJob job[kColsumMaxJobs]; int first_block[kColsumMaxJobs]; int n_jobs; };
int colsum_multi_f32(ColsumBatch jobs, int M, cudaStream_t st);

// ---- topk.cu --------------------------------------------------------------------------------
size_t score_topk_workspace_bytes(int B, int rows, int k);
int score_topk(int mode, const float* pred, int B, int D, const float* table, int row_begin, int row_end, int k, int32_t* idx_out,
               float* score_out, void* ws, size_t ws_bytes, cudaStream_t st);
int merge_topk(const int32_t* in_idx, const float* in_score, int n_lists, int B, int k, int32_t* out_idx,
               float* out_score, cudaStream_t st);

}  // namespace mtam
