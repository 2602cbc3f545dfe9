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
#include "../../include/mtam.h"

namespace mtam {
using namespace tc;

namespace {

constexpr int QM = 128;          // Q tile rows = UMMA M
constexpr int kProdWarps = 4, kProd = kProdWarps * 32;
constexpr float kLog2e = 1.4426950408889634f;
enum { CE_FWD = 0, CE_DP = 1, CE_DT = 2, CE_BMAX = 3 };
// X tile rows = UMMA N of S (and K of O).  The forward pass has no second product and room in TMEM for
// 128-column S buffers; a 128-wide MMA keeps the tensor pipe ahead of the issuing thread.
__host__ __device__ constexpr int ce_bx(int mode) { return (mode == CE_FWD || mode == CE_BMAX) ? 128 : 64; }
__host__ __device__ constexpr bool ce_pv(int mode) { return mode == CE_DP || mode == CE_DT; }
// Epilogue warps.  A warp reaches the TMEM lanes of quadrant (warp % 4), so they come in sets of four; set p handles the
// p-th slice of a tile's columns.  Two sets everywhere: four sets (16 warps, measured with tools/ce_trace.cu) made the
// backward passes SLOWER (dPred pass 117 -> 138 us at cfg3): the 64 x 128 exponentials of a tile keep the MUFU pipe busy
// for about 1000 cycles whatever the number of warps that issue them, and the extra warps only add barrier traffic.
__host__ __device__ constexpr int ce_epi_warps(int mode) { return 8; }
// warp roles: [0, EW) epilogue, [EW, EW + 4) producers, then the S issuer, the O issuer and the bulk-copy warp
__host__ __device__ constexpr int ce_threads(int mode) { return (ce_epi_warps(mode) + kProdWarps + 3) * 32; }
// raw fp32 staging slots (64 X rows each) filled by bulk copies.  CE_DT keeps to 2 (161 KB of shared memory in all) so
// that one of its CTAs fits on an SM beside a T-GRU / hop CTA of the backward chain it overlaps with (model.cu)
__host__ __device__ constexpr int ce_nsg(int mode) { return mode == CE_DT ? 2 : (mode == CE_DP ? 3 : 4); }

// 2^x on the MUFU pipe (ex2.approx: 2 ulp; -inf -> 0)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// developer timeline (tools/ce_trace.cu builds this file with -DMTAM_CE_TRACE): clock64 of pipeline events of CTA (0,0)
#ifdef MTAM_CE_TRACE
__device__ long long g_ce_trace[16 * 64];
#define CE_TRACE(slot, i)                                                                         \
  do {                                                                                            \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (i) < 64) g_ce_trace[(slot) * 64 + (i)] = clock64(); \
  } while (0)
#else
#define CE_TRACE(slot, i) ((void)0)
#endif

// TMEM columns: Q hi/lo (A operand of S); S double-buffered (G hi overwrites S in place); G lo double-buffered;
// O accumulator.  Forward: S buffers are 128 columns wide and nothing else follows.
constexpr uint32_t COL_QH = 0, COL_QL = 64, COL_S0 = 128, TMEM_COLS = 512;
constexpr uint32_t COL_GL0 = 256, COL_O = 384;     // PV modes (S buffers 64 wide: 128, 192; G lo: 256, 320; O: 384, 448)
// The tensor core adds into its fp32 accumulator with truncation, so a chain of n accumulating MMAs drifts by about
// n * 2^-24 of the accumulated magnitude -- always towards zero.  Over the thousand MMAs a CTA of the backward passes
// feeds into O that is 5e-5, far outside fp32-class accuracy (seen as 4e-5 relative error of dpred at 100 K items).
// O is therefore double-buffered and PROMOTED: every kPromote X tiles (96 MMAs) the epilogue warps pull the chunk's
// accumulator out of TMEM and add it to a register accumulator with ordinary round-to-nearest fp32 adds.
constexpr int kPromote = 4;
constexpr int MAXST = 4;

struct CeTcArgs {
  const float* pred;
  const float* table;
  const int32_t* target;
  const float* lse;
  int B, V;
  float inv_batch;
  int tiles_per_cta;        // FWD/DP: X tiles (of ce_bx items) per CTA
  float2* ms_partial;       // FWD: [2*gridDim.x][B]
  float* tlogit;            // FWD: [B]
  float* dpred_partial;     // DP : [gridDim.x][B][D]
  float* dTable;            // DT : [V][D]
  float* bmax;              // BMAX: [B][bmax_ld] maxima of the logits over buckets of bmax_bs (16 / 64) consecutive items
  int bmax_ld, bmax_bs;
  int terms;                // 3: error-compensated hi/lo split (MTAM_GEMM_TF32X3); 1: one MMA per product on operands
                            // rounded to tf32 (MTAM_GEMM_TF32) -- the lo tiles / TMEM columns are then left unused
};

struct Bars {
  uint64_t xk_full[MAXST], xk_empty[MAXST], xm_full[MAXST], xm_empty[MAXST], stg_full[MAXST], stg_empty[MAXST], s_full[2],
      s_empty[2], g_full[2], sg_empty[2], o_full[2], o_empty[2];
};

template <int D, int MODE>
__global__ void __launch_bounds__(ce_threads(MODE), 1) ce_tc_kernel(const CeTcArgs a) {
  constexpr int kEpiWarps = ce_epi_warps(MODE), kEpi = kEpiWarps * 32, EP = kEpiWarps / 4;
  constexpr int kWarpS = kEpiWarps + kProdWarps, kWarpPV = kWarpS + 1, kWarpLd = kWarpPV + 1;
  constexpr int BX = ce_bx(MODE);
  constexpr int KC = D / 32;                      // 32-float (128-byte) chunks of D
  constexpr int XT = BX * 32;                     // floats per chunk tile
  constexpr bool PV = ce_pv(MODE);
  constexpr int NST = 2;                          // operand stages in shared memory
  constexpr int STAGE = (PV ? 4 : 2) * KC * XT;   // floats per operand stage: K-major hi, lo (, MN-major hi, lo)
  constexpr int NSG = ce_nsg(MODE);
  constexpr int SLOT = 64 * D;                    // floats per staging slot
  constexpr uint32_t SW = BX;                     // S buffer width in TMEM columns
  extern __shared__ uint8_t smem_raw[];
  float* Xs = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* Stg = Xs + NST * STAGE;
  __shared__ Bars bars;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float nl2_s[MAXST][64];   // CE_DT: -lse*log2(e) of the X tile's pred rows (-inf past B)
  __shared__ __align__(16) int tgt_s[MAXST][64];     // CE_DT: their target item ids

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool t3 = a.terms != 1;
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
    for (int s = 0; s < MAXST; ++s) {
      mbar_init(&bars.xk_full[s], kProd);
      mbar_init(&bars.xm_full[s], kProd);
      mbar_init(&bars.xk_empty[s], 1);
      mbar_init(&bars.xm_empty[s], 1);
      mbar_init(&bars.stg_full[s], 1);
      mbar_init(&bars.stg_empty[s], kProd);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.s_full[s], 1);
      mbar_init(&bars.s_empty[s], kEpi);
      mbar_init(&bars.g_full[s], kEpi);
      mbar_init(&bars.sg_empty[s], 1);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.o_full[s], 1);
      mbar_init(&bars.o_empty[s], kEpi);
    }
    fence_mbar_init();
  }
  if (warp == kWarpS) tmem_alloc(&tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) CE_TRACE(10, 0);     // CTA set up

  // epilogue thread geometry: TMEM lane quadrant = warp % 4; the EP warps of a quadrant split the columns
  const int quad = warp & 3, half = (warp >> 2) & (EP - 1);
  // the D columns of the Q tile and of O are split over QP of the EP sets, 16 or 32 columns each
  constexpr int QP = (D / 16 < EP) ? D / 16 : EP;
  constexpr int QC = D / QP;
  const bool qo_owner = half < QP;
  const int qrow = q0 + quad * 32 + lane;
  const bool qvalid = qrow < qmax;
  const uint32_t lane_base = tmem + ((uint32_t)(quad * 32) << 16);

  if (warp < kEpiWarps && qo_owner) {
    // stationary Q tile -> TMEM (hi / lo split): thread = row, QC columns each
#pragma unroll
    for (int c = 0; c < QC; c += 16) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (qvalid) v = __ldg(reinterpret_cast<const float4*>(Qsrc + (int64_t)qrow * D + half * QC + c + j));
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float h = tf32_hi(e[u]);
          hi[j + u] = t3 ? __float_as_uint(h) : tf32_rn(e[u]);
          lo[j + u] = __float_as_uint(e[u] - h);
        }
      }
      tmem_st16(lane_base + COL_QH + half * QC + c, hi);
      if (t3) tmem_st16(lane_base + COL_QL + half * QC + c, lo);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) CE_TRACE(11, 0);     // stationary Q tile in TMEM

  if (warp >= kEpiWarps && warp < kWarpS) {
    // ================= producers: global -> registers -> split -> swizzled shared tiles =================
    // work unit = 64 X rows (a whole tile in the PV modes, half a tile in the forward pass); two units are in
    // flight in registers so that the loads of unit u+1 are issued before unit u is stored
    const int pt = tid - kEpi;
    constexpr int UPT = BX / 64;                        // units per tile
    constexpr int PER = 64 * (D / 4) / kProd;
    const int nu = n * UPT;
    float4 ra[PER];
    // unit u: wait for its bulk copy and pull this thread's pieces out of the staging slot.  The slot is handed back
    // in stash(), AFTER the split stores have consumed the loaded registers: an mbarrier arrive does not wait for
    // loads still in flight (they compile to generic LD.E.128 with their own scoreboard), so arriving right after
    // issuing them lets the loader's next bulk copy overwrite the slot under them (seen as rare wrong logits of the
    // tile two ahead, round-1 GPUTEST failure at 10 M rows).
    auto fetch = [&](float4 (&r)[PER], int u) {
      const int sl = u % NSG;
      const int x0 = xt_begin * BX + u * 64;
      mbar_wait(&bars.stg_full[sl], (u / NSG) & 1);
      const uint32_t src = smem_u32(Stg + sl * SLOT);
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int q = pt + j * kProd;
        r[j] = lds128(src + (uint32_t)q * 16u);
        if (x0 + q / (D / 4) >= xmax) r[j] = make_float4(0.f, 0.f, 0.f, 0.f);   // rows the copy did not cover
      }
    };
    auto stash = [&](const float4 (&r)[PER], int u) {
      const int i = u / UPT, hrow = (u % UPT) * 64;
      const int st = i % NST, use = i / NST;
      float* xs = Xs + st * STAGE;
      if (hrow == 0 && i >= NST) mbar_wait(&bars.xk_empty[st], (use - 1) & 1);
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int q = pt + j * kProd;
        const int row = hrow + q / (D / 4), c4 = q % (D / 4);
        if (t3) store_chunk_split<false>(xs + (c4 >> 3) * XT, xs + KC * XT + (c4 >> 3) * XT, row, c4 & 7, r[j]);
        else store_chunk_rn<false>(xs + (c4 >> 3) * XT, row, c4 & 7, r[j]);
      }
      mbar_arrive(&bars.stg_empty[u % NSG]);     // every register of r[] has been read by the stores above
      if (hrow + 64 == BX) {
        fence_proxy_async();
        mbar_arrive(&bars.xk_full[st]);
        if (pt == 0) CE_TRACE(0, i);
      }
      if (PV) {
        if (i >= NST) mbar_wait(&bars.xm_empty[st], (use - 1) & 1);
        float* xm = xs + 2 * KC * XT;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int q = pt + j * kProd;
          const int row = q / (D / 4), c4 = q % (D / 4);
          if (t3) store_chunk_split<true>(xm + (c4 >> 3) * XT, xm + KC * XT + (c4 >> 3) * XT, row, c4 & 7, r[j]);
          else store_chunk_rn<true>(xm + (c4 >> 3) * XT, row, c4 & 7, r[j]);
        }
        if (MODE == CE_DT && pt < 64) {
          const int gr = (xt_begin + i) * BX + pt;
          nl2_s[st][pt] = gr < a.B ? -__ldg(a.lse + gr) * kLog2e : -INFINITY;
          tgt_s[st][pt] = gr < a.B ? __ldg(a.target + gr) : -1;
        }
        fence_proxy_async();
        mbar_arrive(&bars.xm_full[st]);
        if (pt == 0) CE_TRACE(7, i);
      }
    };
    for (int u = 0; u < nu; ++u) {
      fetch(ra, u);
      stash(ra, u);
    }
  } else if (warp == kWarpLd) {
    // ================= loader (converged warp, one elected thread): bulk copies of 64 contiguous X rows into the ring ====
    {
      const int nu = n * (BX / 64);
      for (int u = 0; u < nu; ++u) {
        const int sl = u % NSG;
        const int x0 = xt_begin * BX + u * 64;
        if (u >= NSG) mbar_wait(&bars.stg_empty[sl], ((u / NSG) - 1) & 1);
        const int rows = min(64, xmax - x0);
        const uint32_t bytes = rows > 0 ? (uint32_t)rows * D * 4 : 0;
        if (elect_one()) {
          mbar_arrive_expect_tx(&bars.stg_full[sl], bytes);
          if (bytes) bulk_g2s(Stg + sl * SLOT, Xsrc + (int64_t)x0 * D, bytes, &bars.stg_full[sl]);
        }
      }
    }
  } else if (warp == kWarpS) {
    // ================= S issuer (the warp stays converged, one elected thread issues): S(i) = Q X(i)^T =================
    {
      constexpr uint32_t idS = idesc_tf32(QM, BX, 0, 0);
      for (int i = 0; i < n; ++i) {
        const int st = i % NST, use = i / NST, b = i & 1;
        mbar_wait(&bars.xk_full[st], use & 1);
        if (i >= 2) mbar_wait(PV ? &bars.sg_empty[b] : &bars.s_empty[b], ((i >> 1) - 1) & 1);   // buffer b consumed
        tc_fence_after();
        if (!elect_one()) continue;
        CE_TRACE(1, i);
        const uint32_t xh = desc_lo(smem_u32(Xs + st * STAGE), 16);
        const uint32_t sacc = tmem + COL_S0 + b * SW;
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t qh = tmem + COL_QH + kc * 32 + ks * 8, ql = tmem + COL_QL + kc * 32 + ks * 8;
            const uint64_t bh = desc_join(xh + ((kc * XT * 4 + ks * 32) >> 4), kDescHiK);
            const uint64_t bl = desc_join(xh + (((KC + kc) * XT * 4 + ks * 32) >> 4), kDescHiK);
            bool acc = (kc | ks) != 0;
            if (t3) {                                         // small terms first
              mma_tf32_ts(sacc, ql, bh, idS, acc);
              mma_tf32_ts(sacc, qh, bl, idS, true);
              acc = true;
            }
            mma_tf32_ts(sacc, qh, bh, idS, acc);
          }
        }
        mma_commit(&bars.s_full[b]);
        mma_commit(&bars.xk_empty[st]);
        CE_TRACE(2, i);
      }
    }
  } else if (warp == kWarpPV) {
    // ================= O issuer (converged warp, one elected thread issues): O += G(j) X(j), G read from TMEM =================
    if (PV) {
      constexpr uint32_t idO = idesc_tf32(QM, D, 0, 1);
      for (int j = 0; j < n; ++j) {
        const int st = j % NST, use = j / NST, b = j & 1;
        const int chunk = j / kPromote, ob = chunk & 1;       // O chunk of kPromote tiles -> accumulator ob
        mbar_wait(&bars.xm_full[st], use & 1);
        mbar_wait(&bars.g_full[b], (j >> 1) & 1);
        if (j % kPromote == 0 && chunk >= 2) mbar_wait(&bars.o_empty[ob], ((chunk >> 1) - 1) & 1);   // drained
        tc_fence_after();
        if (!elect_one()) continue;
        CE_TRACE(5, j);
        const uint32_t xmh = desc_lo(smem_u32(Xs + st * STAGE + 2 * KC * XT), BX * 128);
        const uint32_t ghi = tmem + COL_S0 + b * SW, glo = tmem + COL_GL0 + b * SW;
        const uint32_t oacc = tmem + COL_O + ob * 64;
#pragma unroll
        for (int ks = 0; ks < BX / 8; ++ks) {
          const uint64_t bh = desc_join(xmh + ((ks * 1024) >> 4), kDescHiMN);
          const uint64_t bl = desc_join(xmh + ((KC * XT * 4 + ks * 1024) >> 4), kDescHiMN);
          bool acc = ((j % kPromote) | ks) != 0;
          if (t3) {
            mma_tf32_ts(oacc, glo + ks * 8, bh, idO, acc);
            mma_tf32_ts(oacc, ghi + ks * 8, bl, idO, true);
            acc = true;
          }
          mma_tf32_ts(oacc, ghi + ks * 8, bh, idO, acc);
        }
        mma_commit(&bars.xm_empty[st]);
        mma_commit(&bars.sg_empty[b]);
        if (j % kPromote == kPromote - 1 || j == n - 1) mma_commit(&bars.o_full[ob]);
        CE_TRACE(6, j);
      }
    }
  } else {
    // ================= epilogue warps: thread = one Q row (TMEM lane) x 1/EP of the tile's columns =================
    constexpr int HC = BX / EP;   // columns per thread per tile
    int tg = -1;
    float nl2 = 0.f;              // CE_DP: -lse[row]*log2(e)
    float m = -INFINITY, s = 0.f;
    if ((MODE == CE_FWD || MODE == CE_DP) && qvalid) {
      tg = __ldg(a.target + qrow);
      if (MODE == CE_DP) nl2 = -__ldg(a.lse + qrow) * kLog2e;
    }
    const float gscale = qvalid ? a.inv_batch : 0.f;
    constexpr int OC = QC;        // O columns per thread (sets >= QP own none)
    float osum[PV ? OC : 1];      // promoted accumulator (see kPromote)
#pragma unroll
    for (int c = 0; c < (PV ? OC : 1); ++c) osum[c] = 0.f;
    // adds O chunk `chunk` (complete once its last tile's MMAs have run) to osum and hands the buffer back
    auto drain = [&](int chunk) {
      const int ob = chunk & 1;
      mbar_wait(&bars.o_full[ob], (chunk >> 1) & 1);
      tc_fence_after();
      if (qo_owner) {
#pragma unroll
        for (int c = 0; c < OC; c += 16) {
          float v[16];
          tmem_ld16(lane_base + COL_O + ob * 64 + half * OC + c, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) osum[(PV ? c : 0) + (PV ? j : 0)] += v[j];
        }
      }
      tc_fence_before();
      mbar_arrive(&bars.o_empty[ob]);
    };
    for (int i = 0; i < n; ++i) {
      const int st = i % NST, b = i & 1, ub = i >> 1;
      const int x0 = (xt_begin + i) * BX + half * HC;
      mbar_wait(&bars.s_full[b], ub & 1);
      if (MODE == CE_DT) mbar_wait(&bars.xm_full[st], (i / NST) & 1);
      tc_fence_after();
      if (lane == 0 && quad == 0 && half < 2) CE_TRACE(3 + 5 * half, i);
      const uint32_t scol = lane_base + COL_S0 + b * SW + half * HC, gcol = lane_base + COL_GL0 + b * SW + half * HC;
      const bool ragged = x0 + HC > xmax;          // only the catalogue's last tile (warp-uniform)
      if (!PV) {
        // pull this thread's 64 logits out of TMEM in one go and hand the S buffer straight back to the tensor core
        uint32_t r[HC];
        {
          uint32_t (&r0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[0]);
          uint32_t (&r1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[HC - 32]);
          tmem_ld32_nowait(scol, r0);
          if (HC > 32) tmem_ld32_nowait(scol + 32, r1);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(&bars.s_empty[b]);
        if (lane == 0 && quad == 0 && half < 2) CE_TRACE(4 + 5 * half, i);
        float v[HC];
#pragma unroll
        for (int j = 0; j < HC; ++j) v[j] = __uint_as_float(r[j]);
        if (MODE == CE_FWD && (unsigned)(tg - x0) < (unsigned)HC) {
#pragma unroll
          for (int j = 0; j < HC; ++j)
            if (x0 + j == tg) a.tlogit[qrow] = v[j];
        }
        if (ragged) {
#pragma unroll
          for (int j = 0; j < HC; ++j)
            if (x0 + j >= xmax) v[j] = -INFINITY;
        }
        if (MODE == CE_BMAX) {
          // top-k filter pass: only the maximum of every bucket of 16 (or 64) consecutive items leaves the SM
          float g[HC / 16];
#pragma unroll
          for (int q = 0; q < HC / 16; ++q) {
            float e0 = fmaxf(v[q * 16], v[q * 16 + 1]), e1 = fmaxf(v[q * 16 + 2], v[q * 16 + 3]);
#pragma unroll
            for (int j = 4; j < 16; j += 4) {
              e0 = fmaxf(e0, fmaxf(v[q * 16 + j], v[q * 16 + j + 1]));
              e1 = fmaxf(e1, fmaxf(v[q * 16 + j + 2], v[q * 16 + j + 3]));
            }
            g[q] = fmaxf(e0, e1);
          }
          if (qvalid) {
            float* dst = a.bmax + (int64_t)qrow * a.bmax_ld;
            if (a.bmax_bs == 16) *reinterpret_cast<float4*>(dst + (x0 >> 4)) = make_float4(g[0], g[1], g[2], g[3]);
            else dst[x0 >> 6] = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
          }
          continue;
        }
        float c0 = v[0], c1 = v[1], c2 = v[2], c3 = v[3];
#pragma unroll
        for (int j = 4; j < HC; j += 4) {
          c0 = fmaxf(c0, v[j]); c1 = fmaxf(c1, v[j + 1]); c2 = fmaxf(c2, v[j + 2]); c3 = fmaxf(c3, v[j + 3]);
        }
        const float mn = fmaxf(fmaxf(m, fmaxf(c0, c1)), fmaxf(c2, c3));
        if (mn > -INFINITY) {          // false only while every column so far was masked
          s *= ex2((m - mn) * kLog2e); // m = -inf: s is 0 and 2^-inf = 0
          m = mn;
          const float nm2 = -mn * kLog2e;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int j = 0; j < HC; j += 4) {
            a0 += ex2(fmaf(v[j], kLog2e, nm2));
            a1 += ex2(fmaf(v[j + 1], kLog2e, nm2));
            a2 += ex2(fmaf(v[j + 2], kLog2e, nm2));
            a3 += ex2(fmaf(v[j + 3], kLog2e, nm2));
          }
          s += (a0 + a1) + (a2 + a3);
        }
        continue;
      }
#pragma unroll
      for (int c = 0; c < HC; c += 16) {
        float v[16];
        tmem_ld16(scol + c, v);
        const int col0 = x0 + c;
        {
          uint32_t hi[16], lo[16];
          if (MODE == CE_DP) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = ex2(fmaf(v[j], kLog2e, nl2));
            if ((unsigned)(tg - col0) < 16u) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + j == tg) v[j] -= 1.f;
            }
            if (ragged) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (col0 + j >= xmax) v[j] = 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 l4 = *reinterpret_cast<const float4*>(&nl2_s[st][half * HC + c + j]);
              const int4 t4 = *reinterpret_cast<const int4*>(&tgt_s[st][half * HC + c + j]);
              const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
              const int ts[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float p = ex2(fmaf(v[j + u], kLog2e, ls[u]));   // pred rows past B: 2^-inf = 0
                if (ts[u] == qrow) p -= 1.f;
                v[j + u] = p;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float g = v[j] * gscale;          // rows past the Q range: 0
            const float h = tf32_hi(g);
            hi[j] = t3 ? __float_as_uint(h) : tf32_rn(g);
            lo[j] = __float_as_uint(g - h);
          }
          tmem_st16(scol + c, hi);
          if (t3) tmem_st16(gcol + c, lo);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars.g_full[b]);
      if (lane == 0 && quad == 0 && half < 2) CE_TRACE(4 + 5 * half, i);
      // the chunk that ended with tile i-1 has had this tile's epilogue time to finish its MMAs: promote it
      if (PV && i >= 1 && (i - 1) % kPromote == kPromote - 1) drain((i - 1) / kPromote);
    }
    if (MODE == CE_BMAX) {
    } else if (MODE == CE_FWD) {
      if (qvalid) a.ms_partial[(int64_t)(blockIdx.x * EP + half) * a.B + qrow] = make_float2(m, s);
    } else {
      if (PV) drain((n - 1) / kPromote);     // the last chunk (every earlier one was promoted inside the loop)
      if (tid == 0) CE_TRACE(12, 0);   // accumulator complete
      if (qvalid && qo_owner) {
        float* dst = ((MODE == CE_DP) ? a.dpred_partial + ((int64_t)blockIdx.x * a.B + qrow) * D : a.dTable + (int64_t)qrow * D) + half * OC;
#pragma unroll
        for (int c = 0; c < (PV ? OC : 0); c += 4)
          *reinterpret_cast<float4*>(dst + c) = make_float4(osum[c], osum[c + 1], osum[c + 2], osum[c + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) CE_TRACE(13, 0);       // results stored
  if (warp == kWarpS) {
    __syncwarp();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

template <int D, int MODE>
int ce_tc_launch(dim3 grid, const CeTcArgs& a, cudaStream_t st) {
  constexpr int KC = D / 32;
  const size_t smem = (size_t)(2 * (ce_pv(MODE) ? 4 : 2) * KC * ce_bx(MODE) * 32 + ce_nsg(MODE) * 64 * D) * sizeof(float) + 1024;
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(ce_tc_kernel<D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ce_tc_kernel<D, MODE><<<grid, ce_threads(MODE), smem, st>>>(a);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// FWD / DP: the catalogue is cut into gridDim.x ranges of `tiles_per_cta` X tiles, one CTA per (range, row tile)
void ce_tc_partition(int mode, int B, int V, int* G, int* tiles_per_cta) {
  const int total = cdiv(V, ce_bx(mode)), rowtiles = cdiv(B, QM);
  const int want = std::max(1, kNumSMs / rowtiles);
  *tiles_per_cta = cdiv(total, want);
  *G = cdiv(total, *tiles_per_cta);
}

}  // namespace

bool ce_tc_supported(int D) { return D == 32 || D == 64; }

int ce_tc_ranges(int B, int V) {
  int G0, G1, tpc;
  ce_tc_partition(CE_FWD, B, V, &G0, &tpc);
  ce_tc_partition(CE_DP, B, V, &G1, &tpc);
  return std::max(G0, G1);
}

int ce_forward_tc(int D, const float* pred, const float* table, const int32_t* target, int B, int V, void* ws,
                  float* tlogit, float* lse, float* loss_origin, float* block_partial, int* n_partial, cudaStream_t st,
                  int terms) {
  CeTcArgs a{};
  a.terms = terms;
  a.pred = pred; a.table = table; a.target = target; a.B = B; a.V = V;
  int G;
  ce_tc_partition(CE_FWD, B, V, &G, &a.tiles_per_cta);
  a.ms_partial = (float2*)ws;
  a.tlogit = tlogit;
  dim3 grid(G, cdiv(B, QM));
  if (D == 64) MTAM_TRY((ce_tc_launch<64, CE_FWD>(grid, a, st)));
  else if (D == 32) MTAM_TRY((ce_tc_launch<32, CE_FWD>(grid, a, st)));
  else return set_error(-1, "tensor-core softmax CE: num_units=%d not supported (32, 64)", D);
  return ce_finalize(a.ms_partial, 2 * G, B, tlogit, lse, loss_origin, block_partial, n_partial, st);
}

// parts: 1 = dpred only, 2 = the dense item-table gradient only (independent of part 1; may run on another stream), 3 = both
int ce_backward_tc(int D, const float* pred, const float* table, const int32_t* target, const float* lse, int B, int V,
                   float inv_batch, void* ws, float* dTable, float* dpred, cudaStream_t st, int parts, int terms) {
  CeTcArgs a{};
  a.terms = terms;
  a.pred = pred; a.table = table; a.target = target; a.lse = lse; a.B = B; a.V = V; a.inv_batch = inv_batch;
  int G;
  ce_tc_partition(CE_DP, B, V, &G, &a.tiles_per_cta);
  a.dpred_partial = (float*)((char*)ws + ce_ms_region_bytes(B, V));
  a.dTable = dTable;
  dim3 grid_dp(G, cdiv(B, QM)), grid_dt(cdiv(V, QM));
  if (D == 64) {
    if (parts & 1) MTAM_TRY((ce_tc_launch<64, CE_DP>(grid_dp, a, st)));
    if (parts & 2) MTAM_TRY((ce_tc_launch<64, CE_DT>(grid_dt, a, st)));
  } else if (D == 32) {
    if (parts & 1) MTAM_TRY((ce_tc_launch<32, CE_DP>(grid_dp, a, st)));
    if (parts & 2) MTAM_TRY((ce_tc_launch<32, CE_DT>(grid_dt, a, st)));
  } else {
    return set_error(-1, "tensor-core softmax CE: num_units=%d not supported (32, 64)", D);
  }
  if (parts & 1) return ce_reduce_partials(a.dpred_partial, G, (int64_t)B * D, dpred, st);
  return 0;
}

// Top-k filter pass (topk.cu): bmax[b][j] = max over the items of bucket j of <pred[b], table[j*bs + .]>, buckets of
// bs = 16 or 64 consecutive rows of `table` (V rows), ld = 128/bs * ceil(V/128) floats per pred row.
int ce_bucket_max_tc(int D, const float* pred, int B, const float* table, int V, int bs, float* bmax, int ld,
                     cudaStream_t st) {
  if (bs != 16 && bs != 64) return set_error(MTAM_ERR_INVALID, "bucket maxima: bucket size %d not in {16, 64}", bs);
  CeTcArgs a{};
  a.terms = 3;
  a.pred = pred; a.table = table; a.B = B; a.V = V; a.bmax = bmax; a.bmax_ld = ld; a.bmax_bs = bs;
  int G;
  ce_tc_partition(CE_BMAX, B, V, &G, &a.tiles_per_cta);
  dim3 grid(G, cdiv(B, QM));
  if (D == 64) return ce_tc_launch<64, CE_BMAX>(grid, a, st);
  if (D == 32) return ce_tc_launch<32, CE_BMAX>(grid, a, st);
  return set_error(MTAM_ERR_INVALID, "tensor-core scoring: num_units=%d not supported (32, 64)", D);
}

}  // namespace mtam
