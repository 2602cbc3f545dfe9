// Generic GEMM on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the 3-term split
// (fp32-class accuracy), accumulators in TMEM, operands staged in shared memory in the SWIZZLE_128B
// row128 format (tc_common.cuh), fused epilogue straight out of TMEM.
//
// C[M,N] = epi( op(A)[M,K] * op(B)[K,N] ), same contract as gemm_f32 (kernels.h).  Used for the
// K,V / Q|K|V projections, the T-GRU x-side products, the dense4emb layer and their backward passes
// (tf.layers.dense / tf.matmul call sites listed in gemm.cu).
//
// One CTA = 128 threads = one 128 x BN output tile.  Per 32-wide K chunk: all threads load + split
// the operand chunk (coalesced 128-bit loads), one elected thread issues 12 MMAs (4 k-steps x 3
// split terms), completion is signalled through an mbarrier by tcgen05.commit.  Several CTAs are
// resident per SM so one CTA's loads overlap another's MMAs.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "gemm_epi.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "../../include/mtam.h"

namespace mtam {
using namespace tc;

constexpr int TM = 128;   // UMMA M
constexpr int TKC = 32;   // floats of K per chunk (= one 128-byte swizzle row)

// One operand chunk = NBLK blocks of [R x 32] floats of a row-major matrix (block j starts at column
// c0 + 32*j); thread t owns the 16-byte pieces q = t, t+128, ...  The global loads of a whole chunk are
// issued back to back into registers (fetch) and split + stored to the swizzled hi/lo tiles later (stash),
// so that the fetch of chunk k+1 overlaps the MMAs of chunk k.
template <int R, int NBLK>
struct ChunkRegs {
  static constexpr int PER = R * 8 * NBLK / 128;   // float4 per thread
  float4 v[PER];
  __device__ __forceinline__ void fetch(const float* __restrict__ S, int ld, int r0, int c0, int rmax, int cmax, bool vec_ok) {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int q = threadIdx.x + i * 128;
      const int blk = q / (R * 8), qq = q % (R * 8);
      const int row = qq >> 3, chunk = qq & 7;
      const int gr = r0 + row, gc = c0 + 32 * blk + chunk * 4;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < rmax && gc < cmax) {
        const float* p = S + (int64_t)gr * ld + gc;
        if (vec_ok && gc + 3 < cmax) {
          x = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          x.x = __ldg(p);
          if (gc + 1 < cmax) x.y = __ldg(p + 1);
          if (gc + 2 < cmax) x.z = __ldg(p + 2);
          if (gc + 3 < cmax) x.w = __ldg(p + 3);
        }
      }
      v[i] = x;
    }
  }
  template <bool MN>
  __device__ __forceinline__ void stash(float* hi, float* lo) const {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int q = threadIdx.x + i * 128;
      const int blk = q / (R * 8), qq = q % (R * 8);
      store_chunk_split<MN>(hi + blk * R * 32, lo + blk * R * 32, qq >> 3, qq & 7, v[i]);
    }
  }
};

template <int BN, int TA, int TB>
__global__ void __launch_bounds__(128) tc_gemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                      const float* __restrict__ B, int ldb, float* __restrict__ C,
                                                      int ldc, EpiDev epi, int kchunk, float* __restrict__ partial,
                                                      int va, int vb) {
  extern __shared__ uint8_t smem_raw[];
  float* sm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* Ahi = sm;                 // 128 x 32 floats (or 4 blocks of 32 x 32 when TA)
  float* Alo = Ahi + TM * TKC;
  float* Bhi = Alo + TM * TKC;     // BN x 32 floats (or BN/32 blocks of 32 x 32 when !TB)
  float* Blo = Bhi + BN * TKC;
  __shared__ uint64_t mbar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  if (tid == 0) {
    mbar_init(&mbar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tacc = tmem_slot;
  constexpr uint32_t idesc = idesc_tf32(TM, BN, TA ? 1 : 0, TB ? 0 : 1);
  uint32_t parity = 0;
  bool first = true;
  // A: K-major -> one [128 x 32] block; MN-major (TA) -> 4 blocks of [32 k-rows x 32 m]
  ChunkRegs<TA ? 32 : TM, TA ? TM / 32 : 1> ra;
  ChunkRegs<TB ? BN : 32, TB ? 1 : BN / 32> rb;
  auto fetch = [&](int k0) {
    if (!TA) ra.fetch(A, lda, m0, k0, M, kend, va); else ra.fetch(A, lda, k0, m0, kend, M, va);
    if (TB) rb.fetch(B, ldb, n0, k0, N, kend, vb); else rb.fetch(B, ldb, k0, n0, kend, N, vb);
  };
  if (kbeg < kend) fetch(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += TKC) {
    if (k0 != kbeg) {            // the previous chunk's MMAs must be done before its tiles are overwritten
      mbar_wait(&mbar, parity);
      parity ^= 1;
    }
    ra.template stash<TA != 0>(Ahi, Alo);
    rb.template stash<TB == 0>(Bhi, Blo);
    fence_proxy_async();
    __syncthreads();
    if (warp == 0 && elect_one()) {          // converged warp + elect.sync: straight-line MMA issue (tc_common.cuh)
      tc_fence_after();
      const uint32_t ah = smem_u32(Ahi), al = smem_u32(Alo), bh = smem_u32(Bhi), bl = smem_u32(Blo);
#pragma unroll
      for (int ks = 0; ks < TKC / 8; ++ks) {
        const uint64_t dah = TA ? desc_mnmajor(ah, ks, 4096) : desc_kmajor(ah, ks);
        const uint64_t dal = TA ? desc_mnmajor(al, ks, 4096) : desc_kmajor(al, ks);
        const uint64_t dbh = TB ? desc_kmajor(bh, ks) : desc_mnmajor(bh, ks, 4096);
        const uint64_t dbl = TB ? desc_kmajor(bl, ks) : desc_mnmajor(bl, ks, 4096);
        mma_tf32(tacc, dal, dbh, idesc, !first);   // small terms first
        mma_tf32(tacc, dah, dbl, idesc, true);
        mma_tf32(tacc, dah, dbh, idesc, true);
        first = false;
      }
      mma_commit(&mbar);
    }
    if (k0 + TKC < kend) fetch(k0 + TKC);   // global loads of the next chunk fly while the tensor core works
  }
  if (kbeg < kend) mbar_wait(&mbar, parity);
  tc_fence_after();
  // ---- epilogue: TMEM -> registers (thread = row) -> shared staging -> coalesced 128-bit global stores ----
  // every MMA has completed (last mbarrier wait), so the operand tiles can be reused as the staging buffer;
  // each warp stages and writes only its own 32 rows, so a warp-level sync is enough.
  constexpr int SS = BN + 4;
  float* stage = sm + warp * 32 * SS;
#pragma unroll 1
  for (int c = 0; c < BN; c += 16) {
    float v[16];
    tmem_ld16(tacc + ((uint32_t)(warp * 32) << 16) + c, v);
    float* s = stage + lane * SS + c;
#pragma unroll
    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(s + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
  __syncwarp();
  const bool has_k = kbeg < kend;
  const bool vec_out = ((uintptr_t)C % 16 == 0) && (ldc % 4 == 0) && (N % 4 == 0) && !epi.mask_pos && !epi.add && !partial;
  constexpr int LPR = BN / 4;          // lanes per row with float4
  constexpr int RPI = 32 / LPR;        // rows per warp instruction
  for (int r0 = 0; r0 < 32; r0 += RPI) {
    const int r = r0 + lane / LPR, cc = (lane % LPR) * 4;
    const int m = m0 + warp * 32 + r, n = n0 + cc;
    if (m >= M || n >= N) continue;
    float4 v = *reinterpret_cast<const float4*>(stage + r * SS + cc);
    if (!has_k) v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (vec_out) {   // n % 4 == 0 and N % 4 == 0  ->  n + 3 < N
      v.x *= epi.alpha; v.y *= epi.alpha; v.z *= epi.alpha; v.w *= epi.alpha;
      if (epi.bias) {
        float4 b = __ldg(reinterpret_cast<const float4*>(epi.bias + n));
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      if (epi.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
      float4* dst = reinterpret_cast<float4*>(C + (int64_t)m * ldc + n);
      if (epi.accumulate) {
        float4 o = *dst;
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *dst = v;
    } else {
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n + j < N) {
          if (partial) partial[((int64_t)blockIdx.z * M + m) * N + n + j] = e[j];
          else C[(int64_t)m * ldc + n + j] = apply_epi(e[j], m, n + j, epi, C, ldc);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tacc, BN);
}

static int tc_pick_splits(int M, int N, int K, int BN) {
  int tiles = cdiv(M, TM) * cdiv(N, BN);
  if (tiles >= kNumSMs || K < 2048) return 1;
  int want = cdiv(2 * kNumSMs, tiles);
  int maxs = std::max(1, K / 512);
  return std::max(1, std::min(want, maxs));
}
static int tc_bn(int N) { return N > 64 ? 128 : 64; }

size_t tc_gemm_splitk_workspace_bytes(int M, int N, int K) {
  int S = tc_pick_splits(M, N, K, tc_bn(N));
  return S > 1 ? (size_t)S * M * N * sizeof(float) + 256 : 0;
}

template <int BN, int TA, int TB>
static int tc_launch(dim3 grid, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                     const EpiDev& epi, int kchunk, float* partial, int va, int vb, cudaStream_t st) {
  size_t smem = std::max((size_t)(2 * TM * TKC + 2 * BN * TKC), (size_t)TM * (BN + 4)) * sizeof(float) + 1024;
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_kernel<BN, TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_gemm_kernel<BN, TA, TB><<<grid, 128, smem, st>>>(M, N, K, A, lda, B, ldb, C, ldc, epi, kchunk, partial, va, vb);
  MTAM_LAUNCH_CHECK();
  return 0;
}

int gemm_tf32x3(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                int ldc, const GemmEpilogue& e, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  if (K <= 0) return gemm_f32(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, ws, ws_bytes, st);
  if (gemm_ws_supported(transA, M, N, K, A, lda))   // the pipelined, warp-specialised kernel (tc_gemm_ws.cu)
    return gemm_tf32x3_ws(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, ws, ws_bytes, st);
  // unaligned row-major A (no caller in the model): the simple kernel below keeps one accumulator chain per split, which
  // is only fp32-class for short K (the tensor core truncates when it accumulates, see tc_gemm_ws.cu)
  if (K > 1024) return gemm_f32(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, ws, ws_bytes, st);
  EpiDev epi{e.bias, e.mask_pos, e.add, e.ld_mask, e.ld_add, e.relu, e.accumulate, e.alpha};
  const int BN = tc_bn(N);
  int S = tc_pick_splits(M, N, K, BN);
  int kchunk = K;
  if (S > 1) {
    kchunk = cdiv(cdiv(K, S), TKC) * TKC;
    S = cdiv(K, kchunk);
  }
  float* partial = nullptr;
  if (S > 1) {
    size_t need = (size_t)S * M * N * sizeof(float);
    if (!ws || ws_bytes < need) return set_error(MTAM_ERR_WORKSPACE, "tc gemm split-K workspace %zu < %zu", ws_bytes, need);
    partial = (float*)ws;
  }
  const int va = ((uintptr_t)A % 16 == 0) && (lda % 4 == 0), vb = ((uintptr_t)B % 16 == 0) && (ldb % 4 == 0);
  dim3 grid(cdiv(N, BN), cdiv(M, TM), S);
#define TCL(BN_, TA_, TB_) tc_launch<BN_, TA_, TB_>(grid, M, N, K, A, lda, B, ldb, C, ldc, epi, kchunk, partial, va, vb, st)
  int r;
  if (BN == 128) {
    if (!transA && !transB) r = TCL(128, 0, 0);
    else if (!transA && transB) r = TCL(128, 0, 1);
    else if (transA && !transB) r = TCL(128, 1, 0);
    else r = TCL(128, 1, 1);
  } else {
    if (!transA && !transB) r = TCL(64, 0, 0);
    else if (!transA && transB) r = TCL(64, 0, 1);
    else if (transA && !transB) r = TCL(64, 1, 0);
    else r = TCL(64, 1, 1);
  }
#undef TCL
  MTAM_TRY(r);
  if (S > 1) MTAM_TRY(splitk_reduce(partial, S, M, N, C, ldc, epi, st));
  return 0;
}

int gemm_any(int mode, int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
             float* C, int ldc, const GemmEpilogue& e, void* ws, size_t ws_bytes, cudaStream_t st) {
  const bool tiny = (int64_t)cdiv(M, TM) * cdiv(N, 64) <= 4 && K <= 8192;   // a handful of tiles: no tensor-core win
  if ((mode == MTAM_GEMM_TF32X3 || mode == MTAM_GEMM_TF32) && !tiny) return gemm_tf32x3(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, ws, ws_bytes, st);
  return gemm_f32(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, ws, ws_bytes, st);
}
size_t gemm_any_workspace_bytes(int M, int N, int K) {
  return std::max(std::max(gemm_splitk_workspace_bytes(M, N, K), tc_gemm_splitk_workspace_bytes(M, N, K)),
                  gemm_ws_splitk_workspace_bytes(M, N, K));
}

}  // namespace mtam

// test / diagnostic entry: C = op(A) op(B) with the requested mode (0 fp32 FFMA, 1 tcgen05 3xTF32)
extern "C" int mtam_gemm(int32_t mode, int32_t transA, int32_t transB, int32_t M, int32_t N, int32_t K, const float* A,
                         int32_t lda, const float* B, int32_t ldb, float* C, int32_t ldc, const float* bias, int32_t relu,
                         int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream) {
  mtam::GemmEpilogue e;
  e.bias = bias;
  e.relu = relu;
  e.accumulate = accumulate;
  return mtam::gemm_any(mode, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, workspace, workspace_bytes,
                        (cudaStream_t)stream);
}
extern "C" size_t mtam_gemm_workspace(int32_t M, int32_t N, int32_t K) { return mtam::gemm_any_workspace_bytes(M, N, K); }
