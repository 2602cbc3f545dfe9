// Full-catalogue softmax cross-entropy on the 5th-generation tensor cores (MTAM_GEMM_TF32X3).
//
// Reference: base_model.output  Model/base_model.py:300-328
//   logits = pred x item_table^T (:316); log_softmax (:317); one-hot pick (:318-321)
//   backward: G = (softmax - onehot)/B;  dpred = G T;  dT = G^T pred
//
// One warp-specialised kernel template serves the three passes.  A CTA keeps a 128-row "Q" tile
// stationary in shared memory (3xTF32 hi/lo split, K-major SWIZZLE_128B), streams 64-row "X" tiles
// through a two-stage ring, and per X tile
//     S = Q X^T                      tcgen05.mma kind::tf32, A and B from shared memory, D in TMEM
//     G = f(S)                       4 epilogue warps: tcgen05.ld -> exp2 -> hi/lo split -> tcgen05.st
//     O += G X                       tcgen05.mma with the A operand read from TMEM, B = X (MN-major view)
// so the [B,V] logits / softmax never exist in memory (the reference materialises them three times).
//   CE_FWD: Q = pred rows, X = item rows; f = online (max, sum-exp) per row, no second product
//   CE_DP : Q = pred rows, X = item rows; O = this CTA's partial of dpred
//   CE_DT : Q = item rows, X = pred rows; O = the dense item-table gradient rows (complete)
// Warp roles: 0-3 epilogue (TMEM lane quadrant = warp), 4-7 producers (global -> split -> swizzled
// smem), 8 lane 0 issues every MMA.  S/G are double-buffered in TMEM so S(i+1) runs while G(i) is computed.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "model_kernels.h"
#include "tc_common.cuh"

namespace mtam {
using namespace tc;

namespace {

constexpr int QM = 128;          // Q tile rows = UMMA M
constexpr int BX = 64;           // X tile rows = UMMA N of S, K of O
constexpr int kEpi = 128, kProd = 128, kThreads = 288;
constexpr float kLog2e = 1.4426950408889634f;
enum { CE_FWD = 0, CE_DP = 1, CE_DT = 2 };

// 2^x on the MUFU pipe (ex2.approx: 2 ulp; -inf -> 0)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// TMEM columns
constexpr uint32_t COL_S0 = 0, COL_S1 = 64, COL_GL0 = 128, COL_GL1 = 192, COL_O = 256, TMEM_COLS = 512;

struct CeTcArgs {
  const float* pred;
  const float* table;
  const int32_t* target;
  const float* lse;
  int B, V;
  float inv_batch;
  int tiles_per_cta;        // FWD/DP: X tiles (of BX items) per CTA
  float2* ms_partial;       // FWD: [gridDim.x][B]
  float* tlogit;            // FWD: [B]
  float* dpred_partial;     // DP : [gridDim.x][B][D]
  float* dTable;            // DT : [V][D]
};

struct Bars {
  uint64_t xk_full[2], xk_empty[2], xm_full[2], xm_empty[2], s_full[2], s_empty[2], g_full[2], o_full;
};

template <int D, int MODE>
__global__ void __launch_bounds__(kThreads, 1) ce_tc_kernel(const CeTcArgs a) {
  constexpr int KC = D / 32;                      // 32-float (128-byte) chunks of D
  constexpr int QT = QM * 32, XT = BX * 32;       // floats per chunk tile
  constexpr bool PV = MODE != CE_FWD;
  constexpr int STAGE = (PV ? 4 : 2) * KC * XT;   // floats per X stage: K-major hi, lo (, MN-major hi, lo)
  extern __shared__ uint8_t smem_raw[];
  float* sm = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* Qhi = sm;
  float* Qlo = Qhi + KC * QT;
  float* Xs = Qlo + KC * QT;
  __shared__ Bars bars;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float nl2_s[2][BX];     // CE_DT: -lse*log2(e) of the X tile's pred rows (+inf past B)
  __shared__ __align__(16) int tgt_s[2][BX];       // CE_DT: their target item ids

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* Qsrc;
  const float* Xsrc;
  int q0, qmax, xmax, xt_begin, n;
  if (MODE == CE_DT) {
    Qsrc = a.table; q0 = blockIdx.x * QM; qmax = a.V;
    Xsrc = a.pred; xmax = a.B; xt_begin = 0; n = (a.B + BX - 1) / BX;
  } else {
    Qsrc = a.pred; q0 = blockIdx.y * QM; qmax = a.B;
    Xsrc = a.table; xmax = a.V; xt_begin = blockIdx.x * a.tiles_per_cta;
    n = min((a.V + BX - 1) / BX - xt_begin, a.tiles_per_cta);     // >= 1 by construction of the grid
  }

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.xk_full[s], kProd);
      mbar_init(&bars.xm_full[s], kProd);
      mbar_init(&bars.xk_empty[s], 1);
      mbar_init(&bars.xm_empty[s], 1);
      mbar_init(&bars.s_full[s], 1);
      mbar_init(&bars.s_empty[s], kEpi);
      mbar_init(&bars.g_full[s], kEpi);
    }
    mbar_init(&bars.o_full, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(&tmem_slot, TMEM_COLS);
  // stationary Q tile: split into hi/lo, K-major swizzled chunk tiles
  for (int q = tid; q < QM * (D / 4); q += kThreads) {
    const int row = q / (D / 4), c4 = q % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + row < qmax) v = __ldg(reinterpret_cast<const float4*>(Qsrc + (int64_t)(q0 + row) * D + c4 * 4));
    store_chunk_split<false>(Qhi + (c4 >> 3) * QT, Qlo + (c4 >> 3) * QT, row, c4 & 7, v);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp >= 4 && warp < 8) {
    // ================= producers: global -> registers -> split -> swizzled shared tiles =================
    const int pt = tid - 128;
    constexpr int PER = BX * (D / 4) / kProd;
    float4 r[PER];
    auto fetch = [&](int i) {
      const int x0 = (xt_begin + i) * BX;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int q = pt + j * kProd;
        const int row = q / (D / 4), c4 = q % (D / 4);
        r[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x0 + row < xmax) r[j] = __ldg(reinterpret_cast<const float4*>(Xsrc + (int64_t)(x0 + row) * D + c4 * 4));
      }
    };
    fetch(0);
    for (int i = 0; i < n; ++i) {
      const int st = i & 1, use = i >> 1;
      float* xs = Xs + st * STAGE;
      if (i >= 2) mbar_wait(&bars.xk_empty[st], (use - 1) & 1);
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int q = pt + j * kProd;
        const int row = q / (D / 4), c4 = q % (D / 4);
        store_chunk_split<false>(xs + (c4 >> 3) * XT, xs + KC * XT + (c4 >> 3) * XT, row, c4 & 7, r[j]);
      }
      fence_proxy_async();
      mbar_arrive(&bars.xk_full[st]);
      if (PV) {
        if (i >= 2) mbar_wait(&bars.xm_empty[st], (use - 1) & 1);
        float* xm = xs + 2 * KC * XT;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int q = pt + j * kProd;
          const int row = q / (D / 4), c4 = q % (D / 4);
          store_chunk_split<true>(xm + (c4 >> 3) * XT, xm + KC * XT + (c4 >> 3) * XT, row, c4 & 7, r[j]);
        }
        if (MODE == CE_DT && pt < BX) {
          const int gr = (xt_begin + i) * BX + pt;
          nl2_s[st][pt] = gr < a.B ? -__ldg(a.lse + gr) * kLog2e : -INFINITY;
          tgt_s[st][pt] = gr < a.B ? __ldg(a.target + gr) : -1;
        }
        fence_proxy_async();
        mbar_arrive(&bars.xm_full[st]);
      }
      if (i + 1 < n) fetch(i + 1);
    }
  } else if (warp == 8) {
    // ================= MMA issuer (one thread) =================
    if (lane == 0) {
      const uint32_t qh = smem_u32(Qhi), ql = smem_u32(Qlo);
      constexpr uint32_t idS = idesc_tf32(QM, BX, 0, 0);
      constexpr uint32_t idO = idesc_tf32(QM, D, 0, 1);
      for (int i = 0; i <= n; ++i) {
        if (i < n) {   // S(i) = Q X(i)^T
          const int st = i & 1, use = i >> 1;
          mbar_wait(&bars.xk_full[st], use & 1);
          if (!PV && i >= 2) mbar_wait(&bars.s_empty[st], (use - 1) & 1);
          tc_fence_after();
          const uint32_t xh = smem_u32(Xs + st * STAGE), xl = xh + KC * XT * 4;
          const uint32_t sacc = tmem + (st ? COL_S1 : COL_S0);
#pragma unroll
          for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ah = desc_kmajor(qh + kc * QT * 4, ks), al = desc_kmajor(ql + kc * QT * 4, ks);
              const uint64_t bh = desc_kmajor(xh + kc * XT * 4, ks), bl = desc_kmajor(xl + kc * XT * 4, ks);
              mma_tf32(sacc, al, bh, idS, (kc | ks) != 0);   // small terms first
              mma_tf32(sacc, ah, bl, idS, true);
              mma_tf32(sacc, ah, bh, idS, true);
            }
          }
          mma_commit(&bars.s_full[st]);
          mma_commit(&bars.xk_empty[st]);
        }
        if (PV && i >= 1) {   // O += G(j) X(j)
          const int j = i - 1, st = j & 1, use = j >> 1;
          mbar_wait(&bars.xm_full[st], use & 1);
          mbar_wait(&bars.g_full[st], use & 1);
          tc_fence_after();
          const uint32_t xmh = smem_u32(Xs + st * STAGE + 2 * KC * XT), xml = xmh + KC * XT * 4;
          const uint32_t ghi = tmem + (st ? COL_S1 : COL_S0), glo = tmem + (st ? COL_GL1 : COL_GL0);
#pragma unroll
          for (int ks = 0; ks < BX / 8; ++ks) {
            const uint64_t bh = desc_mnmajor(xmh, ks, BX * 128), bl = desc_mnmajor(xml, ks, BX * 128);
            mma_tf32_ts(tmem + COL_O, glo + ks * 8, bh, idO, (j | ks) != 0);
            mma_tf32_ts(tmem + COL_O, ghi + ks * 8, bl, idO, true);
            mma_tf32_ts(tmem + COL_O, ghi + ks * 8, bh, idO, true);
          }
          mma_commit(&bars.xm_empty[st]);
        }
      }
      if (PV) mma_commit(&bars.o_full);
    }
  } else {
    // ================= epilogue warps: thread = one Q row (TMEM lane) =================
    const int qrow = q0 + warp * 32 + lane;
    const bool qvalid = qrow < qmax;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    int tg = -1;
    float nl2 = 0.f;          // CE_DP: -lse[row]*log2(e)
    float m = -INFINITY, s = 0.f;
    if (MODE != CE_DT && qvalid) {
      tg = __ldg(a.target + qrow);
      if (MODE == CE_DP) nl2 = -__ldg(a.lse + qrow) * kLog2e;
    }
    for (int i = 0; i < n; ++i) {
      const int st = i & 1, use = i >> 1;
      const int x0 = (xt_begin + i) * BX;
      mbar_wait(&bars.s_full[st], use & 1);
      if (MODE == CE_DT) mbar_wait(&bars.xm_full[st], use & 1);
      tc_fence_after();
      const uint32_t scol = lane_base + (st ? COL_S1 : COL_S0), gcol = lane_base + (st ? COL_GL1 : COL_GL0);
#pragma unroll 1
      for (int c = 0; c < BX; c += 16) {
        float v[16];
        tmem_ld16(scol + c, v);
        const int col0 = x0 + c;
        if (MODE == CE_FWD) {
          if ((unsigned)(tg - col0) < 16u) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (col0 + j == tg) a.tlogit[qrow] = v[j];
          }
          if (col0 + 16 > xmax) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (col0 + j >= xmax) v[j] = -INFINITY;
          }
          float cm = v[0];
#pragma unroll
          for (int j = 1; j < 16; ++j) cm = fmaxf(cm, v[j]);
          if (cm > m) {
            s *= ex2((m - cm) * kLog2e);     // m = -inf: s is 0 and exp2(-inf) = 0
            m = cm;
          }
          if (m > -INFINITY) {
            const float nm2 = -m * kLog2e;
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) acc += ex2(fmaf(v[j], kLog2e, nm2));
            s += acc;
          }
        } else {
          uint32_t hi[16], lo[16];
          if (MODE == CE_DP) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float p = ex2(fmaf(v[j], kLog2e, nl2));
              if (col0 + j == tg) p -= 1.f;
              float g = (qvalid && col0 + j < xmax) ? p * a.inv_batch : 0.f;
              v[j] = g;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 l4 = *reinterpret_cast<const float4*>(&nl2_s[st][c + j]);
              const int4 t4 = *reinterpret_cast<const int4*>(&tgt_s[st][c + j]);
              const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
              const int ts[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float p = ex2(fmaf(v[j + u], kLog2e, ls[u]));   // rows past B: exp2(-inf) = 0
                if (ts[u] == qrow) p -= 1.f;
                v[j + u] = qvalid ? p * a.inv_batch : 0.f;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float h = tf32_hi(v[j]);
            hi[j] = __float_as_uint(h);
            lo[j] = __float_as_uint(v[j] - h);
          }
          tmem_st16(scol + c, hi);
          tmem_st16(gcol + c, lo);
        }
      }
      if (MODE == CE_FWD) {
        tc_fence_before();
        mbar_arrive(&bars.s_empty[st]);
      } else {
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars.g_full[st]);
      }
    }
    if (MODE == CE_FWD) {
      if (qvalid) a.ms_partial[(int64_t)blockIdx.x * a.B + qrow] = make_float2(m, s);
    } else {
      mbar_wait(&bars.o_full, 0);
      tc_fence_after();
      float* dst = nullptr;
      if (qvalid)
        dst = (MODE == CE_DP) ? a.dpred_partial + ((int64_t)blockIdx.x * a.B + qrow) * D : a.dTable + (int64_t)qrow * D;
#pragma unroll 1
      for (int c = 0; c < D; c += 16) {
        float v[16];
        tmem_ld16(lane_base + COL_O + c, v);
        if (dst) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(dst + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

template <int D, int MODE>
int ce_tc_launch(dim3 grid, const CeTcArgs& a, cudaStream_t st) {
  constexpr int KC = D / 32;
  const size_t smem = (size_t)(2 * KC * QM * 32 + 2 * (MODE == CE_FWD ? 2 : 4) * KC * BX * 32) * sizeof(float) + 1024;
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(ce_tc_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ce_tc_kernel<D, MODE><<<grid, kThreads, smem, st>>>(a);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// FWD / DP: the catalogue is cut into gridDim.x ranges of `tiles_per_cta` X tiles, one CTA per (range, row tile)
void ce_tc_partition(int B, int V, int* G, int* tiles_per_cta) {
  const int total = cdiv(V, BX), rowtiles = cdiv(B, QM);
  const int want = std::max(1, kNumSMs / rowtiles);
  *tiles_per_cta = cdiv(total, want);
  *G = cdiv(total, *tiles_per_cta);
}

}  // namespace

bool ce_tc_supported(int D) { return D == 32 || D == 64; }

int ce_tc_ranges(int B, int V) {
  int G, tpc;
  ce_tc_partition(B, V, &G, &tpc);
  return G;
}

int ce_forward_tc(int D, const float* pred, const float* table, const int32_t* target, int B, int V, void* ws,
                  float* tlogit, float* lse, float* loss_origin, float* block_partial, int* n_partial, cudaStream_t st) {
  CeTcArgs a{};
  a.pred = pred; a.table = table; a.target = target; a.B = B; a.V = V;
  int G;
  ce_tc_partition(B, V, &G, &a.tiles_per_cta);
  a.ms_partial = (float2*)ws;
  a.tlogit = tlogit;
  dim3 grid(G, cdiv(B, QM));
  if (D == 64) MTAM_TRY((ce_tc_launch<64, CE_FWD>(grid, a, st)));
  else if (D == 32) MTAM_TRY((ce_tc_launch<32, CE_FWD>(grid, a, st)));
  else return set_error(-1, "tensor-core softmax CE: num_units=%d not supported (32, 64)", D);
  return ce_finalize(a.ms_partial, G, B, tlogit, lse, loss_origin, block_partial, n_partial, st);
}

int ce_backward_tc(int D, const float* pred, const float* table, const int32_t* target, const float* lse, int B, int V,
                   float inv_batch, void* ws, float* dTable, float* dpred, cudaStream_t st) {
  CeTcArgs a{};
  a.pred = pred; a.table = table; a.target = target; a.lse = lse; a.B = B; a.V = V; a.inv_batch = inv_batch;
  int G;
  ce_tc_partition(B, V, &G, &a.tiles_per_cta);
  a.dpred_partial = (float*)((char*)ws + align_up((size_t)ce_grid(V) * B * sizeof(float2), 256));
  a.dTable = dTable;
  dim3 grid_dp(G, cdiv(B, QM)), grid_dt(cdiv(V, QM));
  if (D == 64) {
    MTAM_TRY((ce_tc_launch<64, CE_DP>(grid_dp, a, st)));
    MTAM_TRY((ce_tc_launch<64, CE_DT>(grid_dt, a, st)));
  } else if (D == 32) {
    MTAM_TRY((ce_tc_launch<32, CE_DP>(grid_dp, a, st)));
    MTAM_TRY((ce_tc_launch<32, CE_DT>(grid_dt, a, st)));
  } else {
    return set_error(-1, "tensor-core softmax CE: num_units=%d not supported (32, 64)", D);
  }
  return ce_reduce_partials(a.dpred_partial, G, (int64_t)B * D, dpred, st);
}

}  // namespace mtam
