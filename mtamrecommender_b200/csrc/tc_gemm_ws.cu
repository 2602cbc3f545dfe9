// Warp-specialised, persistent tcgen05 GEMM (kind::tf32, 3-term split): the pipelined successor of tc_gemm.cu for
// the shapes of this model, which are all HBM-bound (tall [B*L, D] activations against small weights, and their
// weight gradients with K = B*L).  Same contract as gemm_f32 (kernels.h):  C[M,N] = epi(op(A)[M,K] op(B)[K,N]).
//
// Per CTA (one per SM, looping over its (tile, K-split) units):
//   warps 0-3   epilogue     TMEM accumulator -> registers -> per-warp smem staging -> coalesced 128-bit stores (8 row
//                            groups in flight), overlapped with the next unit's MMAs (two accumulators in TMEM)
//   warps 4-7   A producers  thread = one row of the 128 x 32 A chunk: global -> hi/lo split -> tcgen05.st into a
//                            3-stage ring IN TENSOR MEMORY (the A operand of the MMA is read from TMEM, so the
//                            3xTF32 split of A costs no shared-memory bandwidth at all)
//   warps 8-11  B producers  BN x 32 chunk: global -> hi/lo split -> SWIZZLE_128B tiles in a 3-stage smem ring; two
//                            groups of two warps take alternate chunks
//   warp  12    one thread issues the 12 MMAs of a chunk (4 k-steps x {lo*hi, hi*lo, hi*hi}) and commits the
//               stage-release / accumulator-ready mbarriers
// With A in TMEM the shared-memory traffic per 32-wide chunk (B stores 32 KB + B operand reads 48 KB at BN = 128)
// stays under the 768 cycles the MMAs take, so the kernel runs at the speed of its global loads.
//
// Reference call sites: tf.layers.dense / tf.matmul and their tf.gradients, as listed in gemm.cu.
#include <algorithm>

#include "common.cuh"
#include "gemm_epi.cuh"
#include "kernels.h"
#include "tc_common.cuh"
#include "../../include/mtam.h"

namespace mtam {
using namespace tc;

namespace {

constexpr int WM = 128, WK = 32;
// The tensor core adds into its fp32 accumulator with truncation: a chain of n accumulating MMAs drifts by about
// n * 2^-24 of the accumulated magnitude, always towards zero (measured: 1e-3 relative error on a K = 51 200 weight
// gradient).  A unit's K range is therefore cut into sub-units of kSubChunks chunks (192 MMAs); each sub-unit gets a
// fresh TMEM accumulator and the epilogue warps sum the sub-units in shared memory with round-to-nearest fp32 adds.
constexpr int kSubChunks = 16;
constexpr int NSA = 3, NSB = 4;
constexpr int kBGroups = 3;                        // B-producer groups of two warps, one chunk in flight each
constexpr int kWEpi = 0, kWA = 4, kWB = 8, kWMma = kWB + 2 * kBGroups, kWThreads = (kWMma + 1) * 32;
constexpr uint32_t COL_ACC = 0, COL_A = 256;      // accumulators: 2 x BN columns; A ring: NSA x (32 hi + 32 lo)

struct WsArgs {
  int M, N, K;
  const float* A; int lda;
  const float* B; int ldb;
  float* C; int ldc;
  EpiDev epi;
  int kchunk, S, n_mt, n_nt;
  float* partial;
};

struct WsBars {
  uint64_t a_full[NSA], a_empty[NSA], b_full[NSB], b_empty[NSB], acc_full[2], acc_empty[2];
};

// developer timeline (tools/gemm_trace.cu builds this file with -DMTAM_WS_TRACE): clock64 of pipeline events of CTA 0
#ifdef MTAM_WS_TRACE
__device__ long long g_ws_trace[12 * 64];
#define WS_TRACE(slot, i)                                                             \
  do {                                                                                \
    if (blockIdx.x == 0 && (i) < 64) g_ws_trace[(slot) * 64 + (i)] = clock64();      \
  } while (0)
#else
#define WS_TRACE(slot, i) ((void)0)
#endif

// walks the (unit, 32-wide K chunk) sequence of this CTA in the order every role consumes it
struct ChunkIter {
  int u, mt, nt, k0, kend;
  bool valid;
  __device__ __forceinline__ void set_unit(const WsArgs& g, int units) {
    valid = u < units;
    if (!valid) return;
    const int z = u % g.S, t = u / g.S;
    mt = t / g.n_nt;
    nt = t % g.n_nt;
    k0 = z * g.kchunk;
    kend = min(g.K, k0 + g.kchunk);
  }
  __device__ __forceinline__ void init(const WsArgs& g, int units) {
    u = blockIdx.x;
    set_unit(g, units);
  }
  __device__ __forceinline__ void next(const WsArgs& g, int units) {
    k0 += WK;
    if (k0 >= kend) {
      u += gridDim.x;
      set_unit(g, units);
    }
  }
};

template <int BN, int TA, int TB>
__global__ void __launch_bounds__(kWThreads, 1) tc_gemm_ws_kernel(const WsArgs g) {
  extern __shared__ uint8_t smem_raw[];
  float* sm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int BT = BN * WK;                 // floats per B tile (hi or lo)
  float* Bs = sm;                             // [NSB][hi, lo][BT]
  float* Stg = Bs + NSB * 2 * BT;             // [4 warps][32][BN + 4] epilogue staging
  __shared__ WsBars bars;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NSA; ++s) { mbar_init(&bars.a_full[s], 128); mbar_init(&bars.a_empty[s], 1); }
    for (int s = 0; s < NSB; ++s) { mbar_init(&bars.b_full[s], 64); mbar_init(&bars.b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bars.acc_full[s], 1); mbar_init(&bars.acc_empty[s], 128); }
    fence_mbar_init();
  }
  if (warp == kWMma) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int units = g.n_mt * g.n_nt * g.S;

  if (warp >= kWA && warp < kWB) {
    // ================= A producers: thread = row; the loads of chunk c+1 are in flight while chunk c is stored ====
    const int r = tid - kWA * 32;                              // 0..127 = TMEM lane
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + COL_A;
    auto fetch = [&](float (&v)[WK], const ChunkIter& it) {
      const int m = it.mt * WM + r, k0 = it.k0, kend = it.kend;
      if (TA == 0) {                                           // A stored [M][K]: this thread's 32 consecutive floats
        const float* p = g.A + (int64_t)m * g.lda + k0;
        if (m < g.M && k0 + WK <= kend) {
#pragma unroll
          for (int j = 0; j < WK; j += 4) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(p + j));
            v[j] = x.x; v[j + 1] = x.y; v[j + 2] = x.z; v[j + 3] = x.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < WK; ++j) v[j] = ld_nc_pred(p + j, m < g.M && k0 + j < kend);
        }
      } else {                                                 // A stored [K][M]: coalesced across the warp's rows
        const float* p = g.A + (int64_t)k0 * g.lda + m;
#pragma unroll
        for (int j = 0; j < WK; ++j) v[j] = ld_nc_pred(p + (int64_t)j * g.lda, m < g.M && k0 + j < kend);
      }
    };
    auto stash = [&](const float (&v)[WK], int c) {
      const int sa = c % NSA, use = c / NSA;
      if (c >= NSA) mbar_wait(&bars.a_empty[sa], (use - 1) & 1);
      tc_fence_after();
      if (r == 0) WS_TRACE(0, c);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float x = v[h * 16 + j], xh = tf32_hi(x);
          hi[j] = __float_as_uint(xh);
          lo[j] = __float_as_uint(x - xh);
        }
        tmem_st16(lane_base + sa * 64 + h * 16, hi);
        tmem_st16(lane_base + sa * 64 + 32 + h * 16, lo);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars.a_full[sa]);
      if (r == 0) WS_TRACE(1, c);
    };
    ChunkIter it;
    it.init(g, units);
    float va[WK], vb2[WK];
    if (it.valid) fetch(va, it);
    for (int c = 0; it.valid; c += 2) {
      ChunkIter n1 = it;
      n1.next(g, units);
      if (n1.valid) fetch(vb2, n1);
      stash(va, c);
      if (!n1.valid) break;
      it = n1;
      it.next(g, units);
      if (it.valid) fetch(va, it);
      stash(vb2, c + 1);
    }
  } else if (warp >= kWB && warp < kWMma) {
    // ================= B producers: kBGroups groups of two warps take the chunks round-robin =================
    // fence.proxy.async (a MEMBAR) would wait for a prefetched chunk's global loads, so instead of prefetching
    // inside one thread, one group's load latency is overlapped with the other groups' split + store: kBGroups chunks
    // (16 KB each at BN = 128) are in flight per SM, enough to cover the HBM latency.
    const int grp = (warp - kWB) >> 1, pt = tid & 63;
    constexpr int R = TB ? BN : 32, NBLK = TB ? 1 : BN / 32;   // tile = NBLK blocks of [R x 32] floats
    constexpr int PER = R * 8 * NBLK / 64;
    const bool vb = ((uintptr_t)g.B % 16 == 0) && (g.ldb % 4 == 0);
    ChunkIter it;
    it.init(g, units);
    for (int s = 0; s < grp && it.valid; ++s) it.next(g, units);
    for (int c = grp; it.valid; c += kBGroups) {
      const int n0 = it.nt * BN, k0 = it.k0, kend = it.kend;
      const int r0 = TB ? n0 : k0, c0 = TB ? k0 : n0, rmax = TB ? g.N : kend, cmax = TB ? kend : g.N;
      float4 v[PER];
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int q = pt + i * 64;
        const int blk = q / (R * 8), qq = q % (R * 8);
        const int gr = r0 + (qq >> 3), gc = c0 + 32 * blk + (qq & 7) * 4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < rmax && gc < cmax) {
          const float* p = g.B + (int64_t)gr * g.ldb + gc;
          if (vb && gc + 3 < cmax) {
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
      const int sb = c % NSB, use = c / NSB;
      if (c >= NSB) mbar_wait(&bars.b_empty[sb], (use - 1) & 1);
      float* hi = Bs + sb * 2 * BT;
      float* lo = hi + BT;
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int q = pt + i * 64;
        const int blk = q / (R * 8), qq = q % (R * 8);
        store_chunk_split<TB == 0>(hi + blk * R * 32, lo + blk * R * 32, qq >> 3, qq & 7, v[i]);
      }
      fence_proxy_async();
      mbar_arrive(&bars.b_full[sb]);
      if (pt == 0) WS_TRACE(2, c);
      for (int s = 0; s < kBGroups && it.valid; ++s) it.next(g, units);
    }
  } else if (warp == kWMma) {
    // ================= MMA issuer: the warp stays converged, one elected thread issues (tc_common.cuh: elect_one) ====
    {
      constexpr uint32_t idesc = idesc_tf32(WM, BN, 0, TB ? 0 : 1);
      int c = 0, sc = 0;                       // chunk and sub-unit counters of this CTA
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int z = u % g.S;
        const int kbeg = z * g.kchunk, kend = min(g.K, kbeg + g.kchunk);
        for (int ks0 = kbeg; ks0 < kend; ks0 += kSubChunks * WK, ++sc) {
          const int ksend = min(kend, ks0 + kSubChunks * WK);
          const int ab = sc & 1;
          if (sc >= 2) mbar_wait(&bars.acc_empty[ab], ((sc >> 1) - 1) & 1);
          const uint32_t acc = tmem + COL_ACC + ab * BN;
          bool first = true;
          for (int k0 = ks0; k0 < ksend; k0 += WK, ++c) {
            const int sa = c % NSA, sb = c % NSB;
            mbar_wait(&bars.a_full[sa], (c / NSA) & 1);
            mbar_wait(&bars.b_full[sb], (c / NSB) & 1);
            tc_fence_after();
            const uint32_t ah = tmem + COL_A + sa * 64, al = ah + 32;
            const uint32_t bh = smem_u32(Bs + sb * 2 * BT), bl = bh + BT * 4;
            if (elect_one()) {
              WS_TRACE(3, c);
#pragma unroll
              for (int ks = 0; ks < WK / 8; ++ks) {
                const uint64_t dbh = TB ? desc_kmajor(bh, ks) : desc_mnmajor(bh, ks, 4096);
                const uint64_t dbl = TB ? desc_kmajor(bl, ks) : desc_mnmajor(bl, ks, 4096);
                mma_tf32_ts(acc, al + ks * 8, dbh, idesc, !first || ks > 0);   // small terms first
                mma_tf32_ts(acc, ah + ks * 8, dbl, idesc, true);
                mma_tf32_ts(acc, ah + ks * 8, dbh, idesc, true);
              }
              mma_commit(&bars.a_empty[sa]);
              mma_commit(&bars.b_empty[sb]);
              WS_TRACE(4, c);
            }
            first = false;
          }
          if (elect_one()) mma_commit(&bars.acc_full[ab]);
        }
      }
    }
  } else {
    // ================= epilogue: warp = 32 rows; TMEM -> registers -> per-warp smem staging -> coalesced stores ====
    constexpr int SS = BN + 4;
    float* stage = Stg + warp * 32 * SS;
    const uint32_t stage_s = smem_u32(stage);
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16) + COL_ACC;
    const EpiDev& epi = g.epi;
    const bool simple = !epi.mask_pos && !epi.add;
    const bool vec_c = ((uintptr_t)g.C % 16 == 0) && (g.ldc % 4 == 0) && (g.N % 4 == 0);
    const bool vec_p = ((uintptr_t)g.partial % 16 == 0) && (g.N % 4 == 0);
    constexpr int LPR = BN / 4;          // lanes per row with float4
    constexpr int RPI = 32 / LPR;        // rows per warp instruction
    constexpr int UN = 8;                // row groups in flight
    int uc = 0, sc = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++uc) {
      const int z = u % g.S, t = u / g.S, mt = t / g.n_nt, nt = t % g.n_nt;
      const int m0 = mt * WM + warp * 32, n0 = nt * BN;
      const int kbeg = z * g.kchunk, kend = min(g.K, kbeg + g.kchunk);
      // the unit's sub-units arrive one accumulator at a time; they are summed in this thread's row of the staging tile
      for (int ks0 = kbeg; ks0 < kend; ks0 += kSubChunks * WK, ++sc) {
        const int ab = sc & 1;
        mbar_wait(&bars.acc_full[ab], (sc >> 1) & 1);
        tc_fence_after();
        if (tid == 0) WS_TRACE(5, uc);
#pragma unroll
        for (int cc = 0; cc < BN; cc += 32) {
          uint32_t r[32];
          tmem_ld32_nowait(lane_base + ab * BN + cc, r);
          tmem_ld_wait();
          const uint32_t sdst = stage_s + (uint32_t)(lane * SS + cc) * 4u;
          if (ks0 == kbeg) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) sts128(sdst + j * 4, r[j], r[j + 1], r[j + 2], r[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 o = lds128(sdst + j * 4);
              sts128(sdst + j * 4, __float_as_uint(o.x + __uint_as_float(r[j])), __float_as_uint(o.y + __uint_as_float(r[j + 1])),
                     __float_as_uint(o.z + __uint_as_float(r[j + 2])), __float_as_uint(o.w + __uint_as_float(r[j + 3])));
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&bars.acc_empty[ab]);      // the accumulator is in shared memory now: the next sub-unit may start
        if (tid == 0) WS_TRACE(6, uc);
      }
      __syncwarp();
      const int cc = (lane % LPR) * 4, n = n0 + cc, rl = lane / LPR;
      const bool fast = (g.partial ? vec_p : (simple && vec_c));
      if (fast) {
        if (n < g.N) {                       // N % 4 == 0 -> n + 3 < N
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const bool is_c = g.partial == nullptr;
          if (is_c && epi.bias) b4 = __ldg(reinterpret_cast<const float4*>(epi.bias + n));
          const int64_t ld = is_c ? g.ldc : g.N;
          float* gp = (is_c ? g.C + (int64_t)m0 * g.ldc : g.partial + ((int64_t)z * g.M + m0) * g.N) + n + rl * ld;
          const uint32_t sp = stage_s + (uint32_t)(rl * SS + cc) * 4u;
          const int rows_valid = min(32, g.M - m0) - rl;       // this lane stores rows rl + k*RPI < min(32, M - m0)
          const float alpha = is_c ? epi.alpha : 1.f;
          const bool relu = is_c && epi.relu, accum = is_c && epi.accumulate;
#pragma unroll
          for (int it = 0; it < 32 / (RPI * UN); ++it) {
            float4 v[UN], o[UN];
#pragma unroll
            for (int q = 0; q < UN; ++q) {
              const int k = (it * UN + q) * RPI;
              v[q] = lds128(sp + (uint32_t)(k * SS) * 4u);
              o[q] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (accum && k < rows_valid) o[q] = *reinterpret_cast<const float4*>(gp + k * ld);
            }
#pragma unroll
            for (int q = 0; q < UN; ++q) {
              const int k = (it * UN + q) * RPI;
              float4 x = v[q];
              x.x = fmaf(x.x, alpha, b4.x); x.y = fmaf(x.y, alpha, b4.y);
              x.z = fmaf(x.z, alpha, b4.z); x.w = fmaf(x.w, alpha, b4.w);
              if (relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
              x.x += o[q].x; x.y += o[q].y; x.z += o[q].z; x.w += o[q].w;
              if (k < rows_valid) *reinterpret_cast<float4*>(gp + k * ld) = x;
            }
          }
        }
      } else {
        for (int r0 = 0; r0 < 32; r0 += RPI) {
          const int r = r0 + lane / LPR, m = m0 + r;
          if (m >= g.M) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n + j < g.N) {
              const float e = stage[r * SS + cc + j];
              if (g.partial) g.partial[((int64_t)z * g.M + m) * g.N + n + j] = e;
              else g.C[(int64_t)m * g.ldc + n + j] = apply_epi(e, m, n + j, epi, g.C, g.ldc);
            }
          }
        }
      }
      __syncwarp();
      if (tid == 0) WS_TRACE(7, uc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWMma) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

int ws_bn(int N) { return N > 64 ? 128 : 64; }

// K splits: enough (tile, split) units to give every SM a few, each split at least 8 chunks long
int ws_pick_splits(int M, int N, int K) {
  const int tiles = cdiv(M, WM) * cdiv(N, ws_bn(N));
  if (tiles >= kNumSMs || K < 1024) return 1;
  const int want = std::max(1, (2 * kNumSMs) / tiles);   // floor: tiles * splits fills whole rounds of 148 CTAs
  const int maxs = std::max(1, K / (8 * WK));
  return std::max(1, std::min(want, maxs));
}

template <int BN, int TA, int TB>
int ws_launch(const WsArgs& g, cudaStream_t st) {
  const size_t smem = (size_t)(NSB * 2 * BN * WK + 4 * 32 * (BN + 4)) * sizeof(float) + 1024;
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(tc_gemm_ws_kernel<BN, TA, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int units = g.n_mt * g.n_nt * g.S;
  tc_gemm_ws_kernel<BN, TA, TB><<<std::min(units, kNumSMs), kWThreads, smem, st>>>(g);
  MTAM_LAUNCH_CHECK();
  return 0;
}

}  // namespace

bool gemm_ws_supported(int transA, int M, int N, int K, const float* A, int lda) {
  if (M <= 0 || N <= 0 || K <= 0) return false;
  if (!transA && (((uintptr_t)A % 16) != 0 || (lda % 4) != 0)) return false;   // row loads are 128-bit
  return true;
}

size_t gemm_ws_splitk_workspace_bytes(int M, int N, int K) {
  const int S = ws_pick_splits(M, N, K);
  return S > 1 ? (size_t)S * M * N * sizeof(float) + 256 : 0;
}

int gemm_tf32x3_ws(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                   float* C, int ldc, const GemmEpilogue& e, void* ws, size_t ws_bytes, cudaStream_t st) {
  WsArgs g{};
  g.M = M; g.N = N; g.K = K; g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc;
  g.epi = EpiDev{e.bias, e.mask_pos, e.add, e.ld_mask, e.ld_add, e.relu, e.accumulate, e.alpha};
  const int BN = ws_bn(N);
  int S = ws_pick_splits(M, N, K);
  g.kchunk = K;
  if (S > 1) {
    g.kchunk = cdiv(cdiv(K, S), WK) * WK;
    S = cdiv(K, g.kchunk);
  }
  g.S = S;
  g.n_mt = cdiv(M, WM);
  g.n_nt = cdiv(N, BN);
  g.partial = nullptr;
  if (S > 1) {
    const size_t need = (size_t)S * M * N * sizeof(float);
    if (!ws || ws_bytes < need) return set_error(MTAM_ERR_WORKSPACE, "tc gemm split-K workspace %zu < %zu", ws_bytes, need);
    g.partial = (float*)ws;
  }
  int r;
#define WSL(BN_, TA_, TB_) ws_launch<BN_, TA_, TB_>(g, st)
  if (BN == 128) {
    if (!transA && !transB) r = WSL(128, 0, 0);
    else if (!transA && transB) r = WSL(128, 0, 1);
    else if (transA && !transB) r = WSL(128, 1, 0);
    else r = WSL(128, 1, 1);
  } else {
    if (!transA && !transB) r = WSL(64, 0, 0);
    else if (!transA && transB) r = WSL(64, 0, 1);
    else if (transA && !transB) r = WSL(64, 1, 0);
    else r = WSL(64, 1, 1);
  }
#undef WSL
  MTAM_TRY(r);
  if (S > 1) MTAM_TRY(splitk_reduce(g.partial, S, M, N, C, ldc, g.epi, st));
  return 0;
}

}  // namespace mtam
