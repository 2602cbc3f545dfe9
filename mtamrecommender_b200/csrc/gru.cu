// Time-aware GRU ("new" cell) recurrence, forward and backward, as persistent CTA-per-row-tile
// kernels: the recurrent [D,3D] weights stay in shared memory for all L steps, the state stays
// on chip, and the x-side products [B*L,D]x[D,3D] are hoisted out of the loop into one GEMM.
//
// Reference: TimeAwareGRUCell_decay_new.call  Model/Modules/time_aware_rnn.py:186-269
//            dynamic_rnn(sequence_length = seq_len-1)  Model/Modules/gru.py:69-77, MTAMRec_model.py:68-79
//   a  = relu(x*kw1 + kb1 + h*hw1)                 (:228)
//   s  = relu(tw1*dt + tb1)                        (:236)
//   T  = sigmoid(kw2*a + tw12*s + tb12)            (:237)
//   r,u = split(sigmoid([x,h] Wg + bg))            (:243-248)
//   c  = tanh([x, r*h] Wc + bc)                    (:250-256)
//   h' = u*h + (1-u)*c*T                           (:268)
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "model_kernels.h"

namespace mtam {

constexpr int RB = 8;  // batch rows per CTA

// developer timeline (tools/gru_trace.cu builds this file with -DMTAM_GRU_TRACE): clock64 of the phases of step 10, CTA 0
#ifdef MTAM_GRU_TRACE
__device__ long long g_gru_trace[32];
#define GRU_TRACE(slot, t)                                                                       \
  do {                                                                                           \
    if (blockIdx.x == 0 && threadIdx.x == 0 && (t) == 10) g_gru_trace[slot] = clock64();         \
  } while (0)
#else
#define GRU_TRACE(slot, t) ((void)0)
#endif

// vecs layout [8][D]: kw1, kb1, hw1, tw1, tb1, kw2, tw12, tb12   (GRU_LIVE_VECS order)
template <int D>
__global__ void __launch_bounds__(4 * D) gru_fwd_kernel(const float* __restrict__ X, const float* __restrict__ GX,
                                                        const float* __restrict__ timelast,
                                                        const int32_t* __restrict__ seq_len,
                                                        const float* __restrict__ Wgru, const float* __restrict__ vecs,
                                                        int B, int L, float* __restrict__ Hs, float* __restrict__ RUCT,
                                                        float* __restrict__ RH, float* __restrict__ q0, int plain) {
  extern __shared__ __align__(16) float sm[];
  float* Wh = sm;                   // [D][3D]  rows D..2D of W_gru (the h-side)
  float* hT = Wh + 3 * D * D;       // [D][RB]
  float* rhT = hT + D * RB;         // [D][RB]
  float* uS = rhT + D * RB;         // [D][RB]
  __shared__ int steps[RB];
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * RB;
  for (int i = tid; i < 3 * D * D; i += 4 * D) Wh[i] = Wgru[D * 3 * D + i];
  for (int i = tid; i < D * RB; i += 4 * D) hT[i] = 0.f;
  if (tid < RB) steps[tid] = (b0 + tid < B) ? min(max(seq_len[b0 + tid] - 1, 0), L) : 0;
  __syncthreads();
  int tmax = 0;
#pragma unroll
  for (int r = 0; r < RB; ++r) tmax = max(tmax, steps[r]);

  const int n1 = tid % (2 * D), rg1 = tid / (2 * D);  // phase 1: column of [r|u], rows rg1*4..+3
  const int n2 = tid % D, rg2 = tid / D;              // phase 2: column of c,      rows rg2*2..+1
  const float kw1 = vecs[0 * D + n2], kb1 = vecs[1 * D + n2], hw1 = vecs[2 * D + n2], tw1 = vecs[3 * D + n2],
              tb1 = vecs[4 * D + n2], kw2 = vecs[5 * D + n2], tw12 = vecs[6 * D + n2], tb12 = vecs[7 * D + n2];

  // this thread's two weight columns never change over the L steps: keep them in registers (the shared-memory copy
  // is only the staging area), so a step's inner loops read nothing but the broadcast state vectors
  constexpr bool REGW = D <= 64;           // at num_units 128 the 256 values per thread would spill
  constexpr int RW = REGW ? D : 1;
  float w1r[RW], w2r[RW];
  if (REGW) {
#pragma unroll
    for (int k = 0; k < RW; ++k) {
      w1r[k] = Wh[k * 3 * D + n1];
      w2r[k] = Wh[k * 3 * D + 2 * D + n2];
    }
  }

  // x-side inputs of a step do not depend on the recurrence: those of step t+1 are loaded during step t
  struct StepIn { float g1[4], g2[2], xv[2], dl[2]; };
  auto fetch = [&](int t, StepIn& v) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = rg1 * 4 + i;
      const bool live = t < steps[row];
      v.g1[i] = ld_nc_pred(GX + ((int64_t)(b0 + row) * L + (live ? t : 0)) * (3 * D) + n1, live);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = rg2 * 2 + i;
      const bool live = t < steps[row];
      const int64_t tok = (int64_t)(b0 + row) * L + (live ? t : 0);
      v.g2[i] = ld_nc_pred(GX + tok * (3 * D) + 2 * D + n2, live);
      v.xv[i] = ld_nc_pred(X + tok * D + n2, live);
      v.dl[i] = ld_nc_pred(timelast + tok, live);
    }
  };
  StepIn cur, nxt;
  fetch(0, cur);
  for (int t = 0; t < tmax; ++t) {
    fetch(t + 1, nxt);          // rows whose sequence has ended (and t + 1 == tmax) load nothing
    // ---- phase 1: r,u ----
    float g1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) g1[i] = cur.g1[i];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll REGW ? D : 8
    for (int k = 0; k < D; ++k) {
      const float w = REGW ? w1r[REGW ? k : 0] : Wh[k * 3 * D + n1];
      float4 h4 = *reinterpret_cast<const float4*>(&hT[k * RB + rg1 * 4]);
      acc[0] = fmaf(h4.x, w, acc[0]); acc[1] = fmaf(h4.y, w, acc[1]);
      acc[2] = fmaf(h4.z, w, acc[2]); acc[3] = fmaf(h4.w, w, acc[3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int row = rg1 * 4 + i;
      if (t < steps[row]) {
        int64_t tok = (int64_t)(b0 + row) * L + t;
        float v = sigmoidf_(acc[i] + g1[i]);
        RUCT[tok * (4 * D) + n1] = v;  // r at [0,D), u at [D,2D)
        if (n1 < D) {
          float rh = v * hT[n1 * RB + row];
          rhT[n1 * RB + row] = rh;
          RH[tok * D + n1] = rh;
        } else {
          uS[(n1 - D) * RB + row] = v;
        }
      }
    }
    __syncthreads();
    // ---- phase 2: candidate, time gate, state update ----
    float g2[2], xv[2], dl[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) { g2[i] = cur.g2[i]; xv[i] = cur.xv[i]; dl[i] = cur.dl[i]; }
    float acc2[2] = {0.f, 0.f};
#pragma unroll REGW ? D : 8
    for (int k = 0; k < D; ++k) {
      const float w = REGW ? w2r[REGW ? k : 0] : Wh[k * 3 * D + 2 * D + n2];
      float2 r2 = *reinterpret_cast<const float2*>(&rhT[k * RB + rg2 * 2]);
      acc2[0] = fmaf(r2.x, w, acc2[0]);
      acc2[1] = fmaf(r2.y, w, acc2[1]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int row = rg2 * 2 + i;
      if (t < steps[row]) {
        int64_t tok = (int64_t)(b0 + row) * L + t;
        float c = tanhf(acc2[i] + g2[i]);
        float hold = hT[n2 * RB + row];
        float a = fmaxf(fmaf(xv[i], kw1, kb1) + hold * hw1, 0.f);
        float s = fmaxf(fmaf(tw1, dl[i], tb1), 0.f);
        // plain: tf GRUCell (GRU.gru_net, gru.py:60-67) -- no time gate.  T = 1 is also what the backward pass needs:
        // with it every gradient of the gate's parameters is exactly zero (dT * T * (1 - T) = 0)
        float Tg = plain ? 1.f : sigmoidf_(kw2 * a + tw12 * s + tb12);
        float u = uS[n2 * RB + row];
        float hn = u * hold + (1.f - u) * c * Tg;
        RUCT[tok * (4 * D) + 2 * D + n2] = c;
        RUCT[tok * (4 * D) + 3 * D + n2] = Tg;
        Hs[(tok + 1) * D + n2] = hn;  // Hs has one leading zero row
        hT[n2 * RB + row] = hn;       // only this thread touches this element in phase 2
      }
    }
    cur = nxt;
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int row = rg2 * 2 + i;
    if (b0 + row < B) q0[(int64_t)(b0 + row) * D + n2] = hT[n2 * RB + row];
  }
}

// =============================================================================================
// Forward recurrence with the h-side products on the tensor cores (mma.sync m16n8k8, 3xTF32), for the tensor-core
// arithmetic modes at num_units 32 / 64.
//
// The FFMA kernel above is bound by shared-memory RETURN bandwidth, not by latency or FMA issue: every thread pulls
// the whole state through broadcast loads (4*D threads x D x 16 B = 262 KB for the gates + 131 KB for the candidate per
// step per CTA, against 128 B/clk: ~3000 of a step's ~4400 cycles; splitting each column over two threads left the time
// unchanged for that reason).  Operand fragments do not have that problem: per step the gates are the product
//     [r|u]^T [2D x RB] = Wg_h^T [2D x D]  x  h^T [D x RB]          (M = 2D gate columns, N = RB = 8 batch rows, K = D)
// so the constant weights are the A operand and live in registers as fragments for all L steps (hi and lo of the 3xTF32
// split: 64 registers per 16-column tile), the state is the B operand (two conflict-free 4-byte loads per lane per
// k-step, split on the fly), and a warp owns one 16-column tile: D/8 k-steps x 3 MMAs in three independent
// accumulator chains.  The candidate is the same product with r*h as the B operand (the first D/16 warps).
// The C fragment leaves each thread with 2 columns x 2 rows, on which it applies the gates.
// =============================================================================================
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tf32_top(float x) { return __float_as_uint(x) & 0xFFFFE000u; }
// The gates of the tensor-core kernel.  After the products moved to mma.sync, two thirds of a step's ~900 instructions
// per thread were the twelve transcendental calls (expf + IEEE division: ~40 instructions each, one dependent chain per
// warp).  MUFU-based forms: ex2.approx on x * log2(e) and an approximate reciprocal, ~1e-6 relative -- the accuracy
// class of the 3xTF32 products beside them (the exact-fp32 mode keeps expf / tanhf and true division).
__device__ __forceinline__ float sigmoid_mufu(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_mufu(float x) { return 1.f - __fdividef(2.f, __expf(2.f * x) + 1.f); }

template <int D>
__global__ void __launch_bounds__(4 * D) gru_fwd_mma_kernel(const float* __restrict__ X, const float* __restrict__ GX,
                                                            const float* __restrict__ timelast,
                                                            const int32_t* __restrict__ seq_len,
                                                            const float* __restrict__ Wgru, const float* __restrict__ vecs,
                                                            int B, int L, float* __restrict__ Hs, float* __restrict__ RUCT,
                                                            float* __restrict__ RH, float* __restrict__ q0, int plain) {
  static_assert(RB == 8, "the batch rows of a CTA are the N = 8 of the MMA");
  constexpr int NT = 4 * D, KS = D / 8, W2 = D / 16;      // threads (2D/16 warps); k-steps; warps that own a candidate tile
  __shared__ __align__(16) float hT[D * RB], rhT[D * RB], uS[D * RB];   // [k][row]
  __shared__ int steps[RB];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int b0 = blockIdx.x * RB;
  for (int i = tid; i < D * RB; i += NT) hT[i] = 0.f;
  if (tid < RB) steps[tid] = (b0 + tid < B) ? min(max(seq_len[b0 + tid] - 1, 0), L) : 0;
  __syncthreads();
  int tmax = 0;
#pragma unroll
  for (int r = 0; r < RB; ++r) tmax = max(tmax, steps[r]);
  // this thread's outputs: columns cA, cA + 8 of [r|u] (and of c for the first W2 warps) x rows r0, r0 + 1
  const int cA = 16 * warp + gid, r0 = 2 * tig;
  const bool cand = warp < W2;                              // warp-uniform
  const int st0 = steps[r0], st1 = steps[r0 + 1];
  const int64_t tk0 = (int64_t)(b0 + r0) * L, tk1 = tk0 + L;
  // A fragments: A[m][k] = W_gru[D + k][col0 + m]  (a0: m = gid, k = tig; a1: m = gid + 8; a2: k = tig + 4; a3: both)
  const float* Wh = Wgru + (int64_t)D * 3 * D;
  uint32_t ah[KS][4], al[KS][4], ch[KS][4], cl[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = ks * 8 + tig + (j >> 1) * 4, m = (j & 1) * 8;
      const float x = __ldg(Wh + (int64_t)k * 3 * D + cA + m);
      ah[ks][j] = tf32_top(x);
      al[ks][j] = __float_as_uint(x - __uint_as_float(ah[ks][j]));
      const float y = cand ? __ldg(Wh + (int64_t)k * 3 * D + 2 * D + cA + m) : 0.f;
      ch[ks][j] = tf32_top(y);
      cl[ks][j] = __float_as_uint(y - __uint_as_float(ch[ks][j]));
    }
  }
  float kw1[2], kb1[2], hw1[2], tw1[2], tb1[2], kw2[2], tw12[2], tb12[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = (cand ? cA : 0) + 8 * j;
    kw1[j] = vecs[0 * D + c]; kb1[j] = vecs[1 * D + c]; hw1[j] = vecs[2 * D + c]; tw1[j] = vecs[3 * D + c];
    tb1[j] = vecs[4 * D + c]; kw2[j] = vecs[5 * D + c]; tw12[j] = vecs[6 * D + c]; tb12[j] = vecs[7 * D + c];
  }
  // x-side inputs of a step do not depend on the recurrence: those of step t+1 are loaded during step t.
  // index [j][i]: column cA + 8j, row r0 + i
  struct StepIn { float g1[2][2], g2[2][2], xv[2][2], dl[2]; };
  // element offsets of this thread's (row, column cA) in the [token, .] arrays: a step adds t times the row length
  // (32-bit: B * L * 4D < 2^31 is checked by the launcher)
  uint32_t og0 = (uint32_t)tk0 * (3 * D) + cA, og1 = (uint32_t)tk1 * (3 * D) + cA;     // GX
  uint32_t ox0 = (uint32_t)tk0 * D + cA, ox1 = (uint32_t)tk1 * D + cA;                 // X, RH, Hs
  uint32_t or0 = (uint32_t)tk0 * (4 * D) + cA, or1 = (uint32_t)tk1 * (4 * D) + cA;     // RUCT
  uint32_t ot0 = (uint32_t)tk0, ot1 = (uint32_t)tk1;                                   // timelast
  auto fetch = [&](int t, StepIn& v) {
    const bool l0 = t < st0, l1 = t < st1;
    const float* g0 = GX + (og0 + (uint32_t)(l0 ? t : 0) * (3 * D));
    const float* g1p = GX + (og1 + (uint32_t)(l1 ? t : 0) * (3 * D));
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      v.g1[j][0] = ld_nc_pred(g0 + 8 * j, l0);
      v.g1[j][1] = ld_nc_pred(g1p + 8 * j, l1);
    }
    if (cand) {
      const float* x0 = X + (ox0 + (uint32_t)(l0 ? t : 0) * D);
      const float* x1 = X + (ox1 + (uint32_t)(l1 ? t : 0) * D);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        v.g2[j][0] = ld_nc_pred(g0 + 2 * D + 8 * j, l0);
        v.g2[j][1] = ld_nc_pred(g1p + 2 * D + 8 * j, l1);
        v.xv[j][0] = ld_nc_pred(x0 + 8 * j, l0);
        v.xv[j][1] = ld_nc_pred(x1 + 8 * j, l1);
      }
      v.dl[0] = ld_nc_pred(timelast + (ot0 + (uint32_t)(l0 ? t : 0)), l0);
      v.dl[1] = ld_nc_pred(timelast + (ot1 + (uint32_t)(l1 ? t : 0)), l1);
    }
  };
  // D^T tile += A (hi/lo fragments) x B^T, B = src[k][row] in shared memory; three independent accumulator chains
  auto product = [&](const uint32_t (&fh)[KS][4], const uint32_t (&fl)[KS][4], const float* src, float (&out)[4]) {
    float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const float x0 = src[(ks * 8 + tig) * RB + gid], x1 = src[(ks * 8 + tig + 4) * RB + gid];
      const uint32_t h0 = tf32_top(x0), h1 = tf32_top(x1);
      const uint32_t l0 = __float_as_uint(x0 - __uint_as_float(h0)), l1 = __float_as_uint(x1 - __uint_as_float(h1));
      mma_tf32_16x8x8(d0, fl[ks], h0, h1);
      mma_tf32_16x8x8(d1, fh[ks], l0, l1);
      mma_tf32_16x8x8(d2, fh[ks], h0, h1);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = (d0[i] + d1[i]) + d2[i];      // small terms first
  };
  StepIn cur, nxt;
  fetch(0, cur);
  for (int t = 0; t < tmax; ++t) {
    GRU_TRACE(0, t);
    fetch(t + 1, nxt);          // rows whose sequence has ended (and t + 1 == tmax) load nothing
    GRU_TRACE(1, t);
    const bool l0 = t < st0, l1 = t < st1;
    // this step's rows of the saved-activation arrays
    float* const ru0 = RUCT + (or0 + (uint32_t)t * (4 * D));
    float* const ru1 = RUCT + (or1 + (uint32_t)t * (4 * D));
    // ---- phase 1: r,u ----   acc: [0] (cA, r0), [1] (cA, r0+1), [2] (cA+8, r0), [3] (cA+8, r0+1)
    float acc[4];
    product(ah, al, hT, acc);
    GRU_TRACE(2, t);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int col = cA + 8 * j;
      const float va = sigmoid_mufu(acc[2 * j] + cur.g1[j][0]), vb = sigmoid_mufu(acc[2 * j + 1] + cur.g1[j][1]);
      // rows that have ended leave rh / u entries nobody reads (phase 2 discards those rows)
      if (col < D) {                          // warp-uniform (a tile lies on one side of D)
        const float2 h2 = *reinterpret_cast<const float2*>(&hT[col * RB + r0]);
        const float2 rh = make_float2(va * h2.x, vb * h2.y);
        *reinterpret_cast<float2*>(&rhT[col * RB + r0]) = rh;
        if (l0) RH[ox0 + (uint32_t)t * D + 8 * j] = rh.x;
        if (l1) RH[ox1 + (uint32_t)t * D + 8 * j] = rh.y;
      } else {
        *reinterpret_cast<float2*>(&uS[(col - D) * RB + r0]) = make_float2(va, vb);
      }
      if (l0) ru0[8 * j] = va;      // r at [0,D), u at [D,2D)
      if (l1) ru1[8 * j] = vb;
    }
    GRU_TRACE(3, t);
    __syncthreads();
    GRU_TRACE(4, t);
    // ---- phase 2: candidate, time gate, state update (the warps that own a candidate tile) ----
    if (cand) {
      float acc2[4];
      product(ch, cl, rhT, acc2);
      GRU_TRACE(5, t);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = cA + 8 * j;
        const float2 hold2 = *reinterpret_cast<const float2*>(&hT[col * RB + r0]);
        const float2 u2 = *reinterpret_cast<const float2*>(&uS[col * RB + r0]);
        const float hold[2] = {hold2.x, hold2.y}, u[2] = {u2.x, u2.y};
        float hn[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const bool live = i ? l1 : l0;
          const float c = tanh_mufu(acc2[2 * j + i] + cur.g2[j][i]);
          const float a = fmaxf(fmaf(cur.xv[j][i], kw1[j], kb1[j]) + hold[i] * hw1[j], 0.f);
          const float sg = fmaxf(fmaf(tw1[j], cur.dl[i], tb1[j]), 0.f);
          // plain: tf GRUCell (GRU.gru_net, gru.py:60-67) -- no time gate (T = 1: every gradient of its parameters is 0)
          const float Tg = plain ? 1.f : sigmoid_mufu(kw2[j] * a + tw12[j] * sg + tb12[j]);
          hn[i] = live ? u[i] * hold[i] + (1.f - u[i]) * c * Tg : hold[i];
          if (live) {
            float* s4 = (i ? ru1 : ru0) + 2 * D + 8 * j;
            s4[0] = c;
            s4[D] = Tg;
            Hs[(i ? ox1 : ox0) + (uint32_t)(t + 1) * D + 8 * j] = hn[i];      // Hs has one leading zero row
          }
        }
        *reinterpret_cast<float2*>(&hT[col * RB + r0]) = make_float2(hn[0], hn[1]);   // only this thread's elements
      }
    }
    GRU_TRACE(6, t);
    cur = nxt;
    GRU_TRACE(7, t);
    __syncthreads();
    GRU_TRACE(8, t);
  }
  if (cand) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i)
        if (b0 + r0 + i < B) q0[(int64_t)(b0 + r0 + i) * D + cA + 8 * j] = hT[(cA + 8 * j) * RB + r0 + i];
  }
}

// Reverse-time pass.  dGX[t] = [d r_pre | d u_pre | d c_pre] (zero past length; caller memsets),
// dX[t] += element-wise path, vec_partial[block][8][D] = per-CTA sums of the 8 vector-parameter
// gradients.  dq0[b] is d loss / d short_term_intent (state after the last real item); dOut (may be null) is
// d loss / d output[t] for every step (outputs past seq_len-1 are zeros and carry no gradient).
template <int D>
__global__ void __launch_bounds__(4 * D) gru_bwd_kernel(
    const float* __restrict__ X, const float* __restrict__ timelast, const int32_t* __restrict__ seq_len,
    const float* __restrict__ Wgru, const float* __restrict__ vecs, const float* __restrict__ Hs,
    const float* __restrict__ RUCT, const float* __restrict__ dq0, const float* __restrict__ dOut, int B, int L,
    float* __restrict__ dGX, float* __restrict__ dX, float* __restrict__ vec_partial) {
  extern __shared__ __align__(16) float sm[];
  float* WhT = sm;                 // [3D][D]   WhT[n][k] = W_gru[D+k][n]
  float* dpcS = WhT + 3 * D * D;   // [D][RB]
  float* dpgS = dpcS + D * RB;     // [2D][RB]
  __shared__ int steps[RB];
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * RB;
  for (int i = tid; i < 3 * D * D; i += 4 * D) {
    int k = i / (3 * D), n = i % (3 * D);
    WhT[n * D + k] = Wgru[D * 3 * D + i];
  }
  if (tid < RB) steps[tid] = (b0 + tid < B) ? min(max(seq_len[b0 + tid] - 1, 0), L) : 0;
  __syncthreads();
  int tmax = 0;
#pragma unroll
  for (int r = 0; r < RB; ++r) tmax = max(tmax, steps[r]);

  const int n = tid % D, rg = tid / D;  // rows rg*2, rg*2+1
  const float kw1 = vecs[0 * D + n], kb1 = vecs[1 * D + n], hw1 = vecs[2 * D + n], tw1 = vecs[3 * D + n],
              tb1 = vecs[4 * D + n], kw2 = vecs[5 * D + n], tw12 = vecs[6 * D + n], tb12 = vecs[7 * D + n];
  float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // kw1,kb1,hw1,tw1,tb1,kw2,tw12,tb12
  float dh[2] = {0.f, 0.f};

  // (as in the forward kernel: both rows of a thread are computed together and branch-free -- a row that has ended
  // sees zeros everywhere, so only the stores are predicated --, the dot products run 8 independent accumulators, and
  // the rows' step counts, base pointers and d loss / d q0 live in registers)
  int st[2];
  float dq[2];
  float* dgx[2];
  float* dxp[2];
  int64_t tok0[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = rg * 2 + i;
    st[i] = steps[row];
    tok0[i] = (int64_t)(b0 + row) * L;
    dq[i] = (b0 + row < B) ? __ldg(dq0 + (int64_t)(b0 + row) * D + n) : 0.f;
    dgx[i] = dGX + tok0[i] * (3 * D) + n;
    dxp[i] = dX + tok0[i] * D + n;
  }
  // the step's saved activations are independent of the recurrence: the loads of step t-1 are issued at the top of
  // step t and land while its two matrix-vector loops run
  struct StepIn { float r, u, c, Tg, hold, x, dl, dx, dout; };
  auto fetch = [&](int t, StepIn (&v)[2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool live = t >= 0 && t < st[i];
      const int64_t tok = tok0[i] + (live ? t : 0);
      const float* s4 = RUCT + tok * (4 * D);
      v[i].r = ld_nc_pred(s4 + n, live); v[i].u = ld_nc_pred(s4 + D + n, live);
      v[i].c = ld_nc_pred(s4 + 2 * D + n, live); v[i].Tg = ld_nc_pred(s4 + 3 * D + n, live);
      v[i].hold = ld_nc_pred(Hs + tok * D + n, live);      // h_{t-1}: Hs is shifted by one (leading zero row)
      v[i].x = ld_nc_pred(X + tok * D + n, live);
      v[i].dl = ld_nc_pred(timelast + tok, live);
      v[i].dx = ld_cg_pred(dX + tok * D + n, live);         // dX[t] is updated in place: read it a step ahead too
      // d loss / d output[t] when the whole output sequence is consumed (MTAM_via_T_GRU: it is the hops' memory)
      v[i].dout = dOut ? ld_nc_pred(dOut + tok * D + n, live) : 0.f;
    }
  };
  // part of this thread's weight column lives in registers for all L steps (D <= 64): the candidate path
  // (rows 2D..3D of WhT) and the first D rows of the gate path
  constexpr bool REGW = D <= 64;
  constexpr int RW = REGW ? D : 1;
  float wc[RW], wg[RW];
  if (REGW) {
#pragma unroll
    for (int m = 0; m < RW; ++m) {
      wc[m] = WhT[(2 * D + m) * D + n];
      wg[m] = WhT[m * D + n];
    }
  }
  StepIn cur[2], nxt[2];
  fetch(tmax - 1, cur);
  for (int t = tmax - 1; t >= 0; --t) {
    float dhacc[2], rr[2], hh[2], dpc[2], dupre[2];
    fetch(t - 1, nxt);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool live = t < st[i];
      const float r = cur[i].r, u = cur[i].u, c = cur[i].c, Tg = cur[i].Tg, hold = cur[i].hold, x = cur[i].x,
                  dl = cur[i].dl;
      // a row that has not started yet (t >= its step count): dh is still 0 and every saved value was loaded as 0
      const float d = dh[i] + ((t == st[i] - 1) ? dq[i] : 0.f) + cur[i].dout;
      const float du = d * (hold - c * Tg), dc = d * (1.f - u) * Tg, dT = d * (1.f - u) * c;
      dhacc[i] = d * u;
      dpc[i] = dc * (1.f - c * c);
      const float dpT = dT * Tg * (1.f - Tg);
      const float apre = fmaf(x, kw1, kb1) + hold * hw1, spre = fmaf(tw1, dl, tb1);
      const float a = fmaxf(apre, 0.f), s = fmaxf(spre, 0.f);
      const float dpa = (apre > 0.f) ? dpT * kw2 : 0.f;
      const float dps = (spre > 0.f) ? dpT * tw12 : 0.f;
      dhacc[i] = fmaf(dpa, hw1, dhacc[i]);
      g[0] = fmaf(dpa, x, g[0]); g[1] += dpa; g[2] = fmaf(dpa, hold, g[2]);
      g[3] = fmaf(dps, dl, g[3]); g[4] += dps;
      g[5] = fmaf(dpT, a, g[5]); g[6] = fmaf(dpT, s, g[6]); g[7] += dpT;
      dupre[i] = du * u * (1.f - u);
      rr[i] = r; hh[i] = hold;
      if (live) {
        dgx[i][t * (3 * D) + 2 * D] = dpc[i];
        dxp[i][t * D] = fmaf(dpa, kw1, cur[i].dx);   // each element is read (a step earlier) and written exactly once
        dgx[i][t * (3 * D) + D] = dupre[i];
      }
    }
    *reinterpret_cast<float2*>(&dpcS[n * RB + rg * 2]) = make_float2(dpc[0], dpc[1]);
    *reinterpret_cast<float2*>(&dpgS[(D + n) * RB + rg * 2]) = make_float2(dupre[0], dupre[1]);
    __syncthreads();
    // d(r*h)[n] = sum_m dpc[m] * Wc_h[n][m]
    float acc[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
#pragma unroll REGW ? D / 4 : 2
    for (int m = 0; m < D; m += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w = REGW ? wc[REGW ? m + j : 0] : WhT[(2 * D + m + j) * D + n];
        const float2 d2 = *reinterpret_cast<const float2*>(&dpcS[(m + j) * RB + rg * 2]);
        acc[j][0] = fmaf(d2.x, w, acc[j][0]);
        acc[j][1] = fmaf(d2.y, w, acc[j][1]);
      }
    }
    {
      float drpre[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float drh = (acc[0][i] + acc[1][i]) + (acc[2][i] + acc[3][i]);
        dhacc[i] = fmaf(drh, rr[i], dhacc[i]);
        drpre[i] = drh * hh[i] * rr[i] * (1.f - rr[i]);          // 0 for a row that has not started (r = h = 0)
        if (t < st[i]) dgx[i][t * (3 * D)] = drpre[i];
      }
      *reinterpret_cast<float2*>(&dpgS[n * RB + rg * 2]) = make_float2(drpre[0], drpre[1]);
    }
    __syncthreads();
    // dh_prev[n] += sum_m dpg[m] * Wg_h[n][m],  m over 2D
    float acc2[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    if (REGW) {
#pragma unroll
      for (int m = 0; m < RW; m += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 d2 = *reinterpret_cast<const float2*>(&dpgS[(m + j) * RB + rg * 2]);
          acc2[j][0] = fmaf(d2.x, wg[REGW ? m + j : 0], acc2[j][0]);
          acc2[j][1] = fmaf(d2.y, wg[REGW ? m + j : 0], acc2[j][1]);
        }
      }
    }
#pragma unroll 2
    for (int m = REGW ? D : 0; m < 2 * D; m += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w = WhT[(m + j) * D + n];
        const float2 d2 = *reinterpret_cast<const float2*>(&dpgS[(m + j) * RB + rg * 2]);
        acc2[j][0] = fmaf(d2.x, w, acc2[j][0]);
        acc2[j][1] = fmaf(d2.y, w, acc2[j][1]);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (t < st[i]) dh[i] = dhacc[i] + ((acc2[0][i] + acc2[1][i]) + (acc2[2][i] + acc2[3][i]));
      cur[i] = nxt[i];
    }
    __syncthreads();
  }
  // per-CTA reduction of the vector-parameter gradients over the 4 row groups (fixed order)
  float* red = dpgS;  // reuse: [4][8][D] floats = 32*D <= 2D*RB*... (2D*8 = 16D) -> use WhT instead
  red = WhT;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) red[(rg * 8 + j) * D + n] = g[j];
  __syncthreads();
  if (rg == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = red[(0 * 8 + j) * D + n] + red[(1 * 8 + j) * D + n] + red[(2 * 8 + j) * D + n] +
                red[(3 * 8 + j) * D + n];
      vec_partial[((int64_t)blockIdx.x * 8 + j) * D + n] = s;
    }
  }
}

// The same reverse-time pass with its two matrix-vector products per step on mma.sync (3xTF32), for the tensor-core
// arithmetic modes at num_units 32 / 64 (see gru_fwd_mma_kernel).  The element-wise work keeps the FFMA kernel's thread
// mapping (column n = tid % D, rows 2 rg, 2 rg + 1); the products' halves travel through `part`.
template <int D>
__global__ void __launch_bounds__(4 * D) gru_bwd_mma_kernel(
    const float* __restrict__ X, const float* __restrict__ timelast, const int32_t* __restrict__ seq_len,
    const float* __restrict__ Wgru, const float* __restrict__ vecs, const float* __restrict__ Hs,
    const float* __restrict__ RUCT, const float* __restrict__ dq0, const float* __restrict__ dOut, int B, int L,
    float* __restrict__ dGX, float* __restrict__ dX, float* __restrict__ vec_partial) {
  static_assert(RB == 8, "the batch rows of a CTA are the N = 8 of the MMA");
  __shared__ __align__(16) float dpcS[D * RB];        // [m][row]: B operand of the candidate product
  __shared__ __align__(16) float dpgS[2 * D * RB];    // [m][row]: B operand of the gate product
  __shared__ __align__(16) float part[2][D * RB];     // [k half][n][row]: the two halves of a product's k range
  __shared__ __align__(16) float red[4 * 8 * D];      // the final reduction of the vector-parameter gradients
  __shared__ int steps[RB];
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * RB;
  if (tid < RB) steps[tid] = (b0 + tid < B) ? min(max(seq_len[b0 + tid] - 1, 0), L) : 0;
  __syncthreads();
  int tmax = 0;
#pragma unroll
  for (int r = 0; r < RB; ++r) tmax = max(tmax, steps[r]);

  const int n = tid % D, rg = tid / D;  // rows rg*2, rg*2+1
  const float kw1 = vecs[0 * D + n], kb1 = vecs[1 * D + n], hw1 = vecs[2 * D + n], tw1 = vecs[3 * D + n],
              tb1 = vecs[4 * D + n], kw2 = vecs[5 * D + n], tw12 = vecs[6 * D + n], tb12 = vecs[7 * D + n];
  float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // kw1,kb1,hw1,tw1,tb1,kw2,tw12,tb12
  float dh[2] = {0.f, 0.f};

  // (as in the forward kernel: both rows of a thread are computed together and branch-free -- a row that has ended
  // sees zeros everywhere, so only the stores are predicated --, the dot products run 8 independent accumulators, and
  // the rows' step counts, base pointers and d loss / d q0 live in registers)
  int st[2];
  float dq[2];
  float* dgx[2];
  float* dxp[2];
  int64_t tok0[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = rg * 2 + i;
    st[i] = steps[row];
    tok0[i] = (int64_t)(b0 + row) * L;
    dq[i] = (b0 + row < B) ? __ldg(dq0 + (int64_t)(b0 + row) * D + n) : 0.f;
    dgx[i] = dGX + tok0[i] * (3 * D) + n;
    dxp[i] = dX + tok0[i] * D + n;
  }
  // the step's saved activations are independent of the recurrence: the loads of step t-1 are issued at the top of
  // step t and land while its two matrix-vector loops run
  struct StepIn { float r, u, c, Tg, hold, x, dl, dx, dout; };
  auto fetch = [&](int t, StepIn (&v)[2]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool live = t >= 0 && t < st[i];
      const int64_t tok = tok0[i] + (live ? t : 0);
      const float* s4 = RUCT + tok * (4 * D);
      v[i].r = ld_nc_pred(s4 + n, live); v[i].u = ld_nc_pred(s4 + D + n, live);
      v[i].c = ld_nc_pred(s4 + 2 * D + n, live); v[i].Tg = ld_nc_pred(s4 + 3 * D + n, live);
      v[i].hold = ld_nc_pred(Hs + tok * D + n, live);      // h_{t-1}: Hs is shifted by one (leading zero row)
      v[i].x = ld_nc_pred(X + tok * D + n, live);
      v[i].dl = ld_nc_pred(timelast + tok, live);
      v[i].dx = ld_cg_pred(dX + tok * D + n, live);         // dX[t] is updated in place: read it a step ahead too
      // d loss / d output[t] when the whole output sequence is consumed (MTAM_via_T_GRU: it is the hops' memory)
      v[i].dout = dOut ? ld_nc_pred(dOut + tok * D + n, live) : 0.f;
    }
  };
  // Both products of a step are  out[n][row] = sum_m A[n][m] * src[m][row]  with constant A: the candidate path
  // A[n][m] = W_gru[D+n][2D+m] (m < D) and the gate path A[n][m] = W_gru[D+n][m] (m < 2D).  M = D output columns = D/16
  // tiles; a tile is shared by the warp PAIR (w, w + D/16), each taking half of the k-steps, so that all warps work and the
  // fragments (3xTF32 hi / lo) of both products fit the register file: (D/16 + D/8) k-steps x 8 registers.
  constexpr int NTILE = D / 16, KC = D / 16, KG = D / 8;       // k-steps per warp: candidate product, gate product
  const int warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int tile = warp % NTILE, khalf = warp / NTILE;
  const float* Wh = Wgru + (int64_t)D * 3 * D;                 // W_gru[D + n][.]
  uint32_t fch[KC][4], fcl[KC][4], fgh[KG][4], fgl[KG][4];
#pragma unroll
  for (int ks = 0; ks < KC; ++ks)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = 16 * tile + gid + (j & 1) * 8, m = (khalf * KC + ks) * 8 + tig + (j >> 1) * 4;
      const float x = __ldg(Wh + (int64_t)nn * 3 * D + 2 * D + m);
      fch[ks][j] = tf32_top(x);
      fcl[ks][j] = __float_as_uint(x - __uint_as_float(fch[ks][j]));
    }
#pragma unroll
  for (int ks = 0; ks < KG; ++ks)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = 16 * tile + gid + (j & 1) * 8, m = (khalf * KG + ks) * 8 + tig + (j >> 1) * 4;
      const float x = __ldg(Wh + (int64_t)nn * 3 * D + m);
      fgh[ks][j] = tf32_top(x);
      fgl[ks][j] = __float_as_uint(x - __uint_as_float(fgh[ks][j]));
    }
  // this warp's half of a product -> part[khalf][n][row]  (C fragment: n = 16 tile + gid (+8), rows 2 tig, 2 tig + 1)
  auto store_part = [&](const float (&o)[4]) {
    float* pp = part[khalf] + (16 * tile + gid) * RB + 2 * tig;
    *reinterpret_cast<float2*>(pp) = make_float2(o[0], o[1]);
    *reinterpret_cast<float2*>(pp + 8 * RB) = make_float2(o[2], o[3]);
  };
  StepIn cur[2], nxt[2];
  fetch(tmax - 1, cur);
  for (int t = tmax - 1; t >= 0; --t) {
    float dhacc[2], rr[2], hh[2], dpc[2], dupre[2];
    fetch(t - 1, nxt);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool live = t < st[i];
      const float r = cur[i].r, u = cur[i].u, c = cur[i].c, Tg = cur[i].Tg, hold = cur[i].hold, x = cur[i].x,
                  dl = cur[i].dl;
      // a row that has not started yet (t >= its step count): dh is still 0 and every saved value was loaded as 0
      const float d = dh[i] + ((t == st[i] - 1) ? dq[i] : 0.f) + cur[i].dout;
      const float du = d * (hold - c * Tg), dc = d * (1.f - u) * Tg, dT = d * (1.f - u) * c;
      dhacc[i] = d * u;
      dpc[i] = dc * (1.f - c * c);
      const float dpT = dT * Tg * (1.f - Tg);
      const float apre = fmaf(x, kw1, kb1) + hold * hw1, spre = fmaf(tw1, dl, tb1);
      const float a = fmaxf(apre, 0.f), s = fmaxf(spre, 0.f);
      const float dpa = (apre > 0.f) ? dpT * kw2 : 0.f;
      const float dps = (spre > 0.f) ? dpT * tw12 : 0.f;
      dhacc[i] = fmaf(dpa, hw1, dhacc[i]);
      g[0] = fmaf(dpa, x, g[0]); g[1] += dpa; g[2] = fmaf(dpa, hold, g[2]);
      g[3] = fmaf(dps, dl, g[3]); g[4] += dps;
      g[5] = fmaf(dpT, a, g[5]); g[6] = fmaf(dpT, s, g[6]); g[7] += dpT;
      dupre[i] = du * u * (1.f - u);
      rr[i] = r; hh[i] = hold;
      if (live) {
        dgx[i][t * (3 * D) + 2 * D] = dpc[i];
        dxp[i][t * D] = fmaf(dpa, kw1, cur[i].dx);   // each element is read (a step earlier) and written exactly once
        dgx[i][t * (3 * D) + D] = dupre[i];
      }
    }
    *reinterpret_cast<float2*>(&dpcS[n * RB + rg * 2]) = make_float2(dpc[0], dpc[1]);
    *reinterpret_cast<float2*>(&dpgS[(D + n) * RB + rg * 2]) = make_float2(dupre[0], dupre[1]);
    __syncthreads();
    // d(r*h)[n] = sum_m dpc[m] * Wc_h[n][m]: each warp its half of the k range, then the halves meet in shared memory
    {
      float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f}, o[4];
#pragma unroll
      for (int ks = 0; ks < KC; ++ks) {
        const int kk = (khalf * KC + ks) * 8 + tig;
        const float x0 = dpcS[kk * RB + gid], x1 = dpcS[(kk + 4) * RB + gid];
        const uint32_t h0 = tf32_top(x0), h1 = tf32_top(x1);
        const uint32_t l0 = __float_as_uint(x0 - __uint_as_float(h0)), l1 = __float_as_uint(x1 - __uint_as_float(h1));
        mma_tf32_16x8x8(d0, fcl[ks], h0, h1);
        mma_tf32_16x8x8(d1, fch[ks], l0, l1);
        mma_tf32_16x8x8(d2, fch[ks], h0, h1);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = (d0[i] + d1[i]) + d2[i];
      store_part(o);
    }
    __syncthreads();
    const float2 pb0 = *reinterpret_cast<const float2*>(&part[0][n * RB + rg * 2]);
    const float2 pb1 = *reinterpret_cast<const float2*>(&part[1][n * RB + rg * 2]);
    const float acc_b[2] = {pb0.x + pb1.x, pb0.y + pb1.y};
    {
      float drpre[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float drh = acc_b[i];
        dhacc[i] = fmaf(drh, rr[i], dhacc[i]);
        drpre[i] = drh * hh[i] * rr[i] * (1.f - rr[i]);          // 0 for a row that has not started (r = h = 0)
        if (t < st[i]) dgx[i][t * (3 * D)] = drpre[i];
      }
      *reinterpret_cast<float2*>(&dpgS[n * RB + rg * 2]) = make_float2(drpre[0], drpre[1]);
    }
    __syncthreads();
    // dh_prev[n] += sum_m dpg[m] * Wg_h[n][m],  m over 2D
    {
      float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f}, o[4];
#pragma unroll
      for (int ks = 0; ks < KG; ++ks) {
        const int kk = (khalf * KG + ks) * 8 + tig;
        const float x0 = dpgS[kk * RB + gid], x1 = dpgS[(kk + 4) * RB + gid];
        const uint32_t h0 = tf32_top(x0), h1 = tf32_top(x1);
        const uint32_t l0 = __float_as_uint(x0 - __uint_as_float(h0)), l1 = __float_as_uint(x1 - __uint_as_float(h1));
        mma_tf32_16x8x8(d0, fgl[ks], h0, h1);
        mma_tf32_16x8x8(d1, fgh[ks], l0, l1);
        mma_tf32_16x8x8(d2, fgh[ks], h0, h1);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) o[i] = (d0[i] + d1[i]) + d2[i];
      store_part(o);       // (the candidate product's halves were read before the barrier above)
    }
    __syncthreads();
    const float2 pc0 = *reinterpret_cast<const float2*>(&part[0][n * RB + rg * 2]);
    const float2 pc1 = *reinterpret_cast<const float2*>(&part[1][n * RB + rg * 2]);
    const float acc_c[2] = {pc0.x + pc1.x, pc0.y + pc1.y};
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (t < st[i]) dh[i] = dhacc[i] + acc_c[i];
      cur[i] = nxt[i];
    }
    __syncthreads();
  }
  // per-CTA reduction of the vector-parameter gradients over the 4 row groups (fixed order)
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) red[(rg * 8 + j) * D + n] = g[j];
  __syncthreads();
  if (rg == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = red[(0 * 8 + j) * D + n] + red[(1 * 8 + j) * D + n] + red[(2 * 8 + j) * D + n] +
                red[(3 * 8 + j) * D + n];
      vec_partial[((int64_t)blockIdx.x * 8 + j) * D + n] = s;
    }
  }
}

static size_t gru_smem_bytes(int D) { return (size_t)(3 * D * D + 3 * D * RB) * sizeof(float); }
int gru_num_blocks(int B) { return cdiv(B, RB); }

template <int D>
static int gru_fwd_launch(const float* X, const float* GX, const float* timelast, const int32_t* seq_len,
                          const float* Wgru, const float* vecs, int B, int L, float* Hs, float* RUCT, float* RH,
                          float* q0, int plain, cudaStream_t st) {
  size_t smem = gru_smem_bytes(D);
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(gru_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gru_fwd_kernel<D><<<gru_num_blocks(B), 4 * D, smem, st>>>(X, GX, timelast, seq_len, Wgru, vecs, B, L, Hs, RUCT, RH, q0, plain);
  MTAM_LAUNCH_CHECK();
  return 0;
}
template <int D>
static int gru_bwd_launch(const float* X, const float* timelast, const int32_t* seq_len, const float* Wgru,
                          const float* vecs, const float* Hs, const float* RUCT, const float* dq0, const float* dOut, int B, int L,
                          float* dGX, float* dX, float* vec_partial, cudaStream_t st) {
  size_t smem = gru_smem_bytes(D);
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(gru_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gru_bwd_kernel<D><<<gru_num_blocks(B), 4 * D, smem, st>>>(X, timelast, seq_len, Wgru, vecs, Hs, RUCT, dq0, dOut, B, L, dGX,
                                                          dX, vec_partial);
  MTAM_LAUNCH_CHECK();
  return 0;
}

int gru_forward(int D, const float* X, const float* GX, const float* timelast, const int32_t* seq_len,
                const float* Wgru, const float* vecs, int B, int L, float* Hs, float* RUCT, float* RH, float* q0,
                cudaStream_t st, int plain, int tensor_cores) {
  // h-side products on mma.sync (3xTF32); num_units 128 and arrays of 2^31 elements or more: the FFMA kernel
  if (tensor_cores && (D == 64 || D == 32) && (int64_t)gru_num_blocks(B) * RB * L * 4 * D < (1ll << 31)) {
    if (D == 64) gru_fwd_mma_kernel<64><<<gru_num_blocks(B), 256, 0, st>>>(X, GX, timelast, seq_len, Wgru, vecs, B, L, Hs, RUCT, RH, q0, plain);
    else gru_fwd_mma_kernel<32><<<gru_num_blocks(B), 128, 0, st>>>(X, GX, timelast, seq_len, Wgru, vecs, B, L, Hs, RUCT, RH, q0, plain);
    MTAM_LAUNCH_CHECK();
    return 0;
  }
  switch (D) {
    case 32: return gru_fwd_launch<32>(X, GX, timelast, seq_len, Wgru, vecs, B, L, Hs, RUCT, RH, q0, plain, st);
    case 64: return gru_fwd_launch<64>(X, GX, timelast, seq_len, Wgru, vecs, B, L, Hs, RUCT, RH, q0, plain, st);
    case 128: return gru_fwd_launch<128>(X, GX, timelast, seq_len, Wgru, vecs, B, L, Hs, RUCT, RH, q0, plain, st);
  }
  return set_error(-1, "T-GRU: num_units=%d not supported (32, 64, 128)", D);
}
int gru_backward(int D, const float* X, const float* timelast, const int32_t* seq_len, const float* Wgru,
                 const float* vecs, const float* Hs, const float* RUCT, const float* dq0, const float* dOut, int B, int L, float* dGX,
                 float* dX, float* vec_partial, cudaStream_t st, int tensor_cores) {
  if (tensor_cores && (D == 64 || D == 32)) {
    if (D == 64) gru_bwd_mma_kernel<64><<<gru_num_blocks(B), 256, 0, st>>>(X, timelast, seq_len, Wgru, vecs, Hs, RUCT, dq0, dOut, B, L, dGX, dX, vec_partial);
    else gru_bwd_mma_kernel<32><<<gru_num_blocks(B), 128, 0, st>>>(X, timelast, seq_len, Wgru, vecs, Hs, RUCT, dq0, dOut, B, L, dGX, dX, vec_partial);
    MTAM_LAUNCH_CHECK();
    return 0;
  }
  switch (D) {
    case 32: return gru_bwd_launch<32>(X, timelast, seq_len, Wgru, vecs, Hs, RUCT, dq0, dOut, B, L, dGX, dX, vec_partial, st);
    case 64: return gru_bwd_launch<64>(X, timelast, seq_len, Wgru, vecs, Hs, RUCT, dq0, dOut, B, L, dGX, dX, vec_partial, st);
    case 128: return gru_bwd_launch<128>(X, timelast, seq_len, Wgru, vecs, Hs, RUCT, dq0, dOut, B, L, dGX, dX, vec_partial, st);
  }
  return set_error(-1, "T-GRU: num_units=%d not supported (32, 64, 128)", D);
}

}  // namespace mtam
