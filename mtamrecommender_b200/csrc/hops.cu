// MTAM multi-hop time-aware attentive memory read (Tq = 1), forward and backward.
// One warp owns one sequence for all N hops: the query never leaves registers, keys/values are
// streamed once per hop with full-line coalesced reads (lanes across D), the time gate is computed
// in-register.  HBM/L2-bound by the K,V reads; no tensor-core reshaping.
//
// Reference: Time_Aware_Attention.vanilla_attention  Model/Modules/time_aware_attention.py:524-556
//            time_aware_multihead_attention          :215-456   (math restated in SURVEY 9.4)
//            final tf.contrib layer_norm             Model/MTAMRec_model.py:91, net_utils.py:229-232
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "model_kernels.h"

namespace mtam {

constexpr int HOP_WARPS = 4;

template <int VPL>
__device__ __forceinline__ void ldv(float (&r)[VPL], const float* p) {
  if (VPL == 4) {
    float4 t = *reinterpret_cast<const float4*>(p);
    r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
  } else if (VPL == 2) {
    float2 t = *reinterpret_cast<const float2*>(p);
    r[0] = t.x; r[1] = t.y;
  } else {
#pragma unroll
    for (int v = 0; v < VPL; ++v) r[v] = p[v];
  }
}
template <int VPL>
__device__ __forceinline__ void stv(float* p, const float (&r)[VPL]) {
  if (VPL == 4) *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
  else if (VPL == 2) *reinterpret_cast<float2*>(p) = make_float2(r[0], r[1]);
  else {
#pragma unroll
    for (int v = 0; v < VPL; ++v) p[v] = r[v];
  }
}
__device__ __forceinline__ float group_sum_rt(float v, int lanes) {
  for (int o = lanes >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int VPL>
__global__ void __launch_bounds__(HOP_WARPS * 32) hop_fwd_kernel(HopArgs a) {
  constexpr int D = VPL * 32;
  extern __shared__ __align__(16) float sm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * HOP_WARPS + w;
  if (b >= a.B) return;
  const int L = a.L, H = a.H, N = a.N;
  float* qs = sm + (size_t)w * (D + 2 * H * L);
  float* sc = qs + D;
  const int d0 = lane * VPL;
  const int lph = 32 / H, head = lane / lph;
  const float sqrt_dh = sqrtf((float)(D / H));
  const int len = min(max(a.seq_len[b], 0), L);
  const float tq = a.target_time[b];
  const int ldkv = 2 * N * D;

  float q[VPL];
  ldv<VPL>(q, a.Qin + (int64_t)b * D + d0);
  for (int i = 0; i < N; ++i) {
    stv<VPL>(qs + d0, q);
    __syncwarp();
    // Q = relu(q Wq + bq), qt = q Wt   (:249, :320)
    float Q[VPL], qt[VPL];
    ldv<VPL>(Q, a.bq + (int64_t)i * D + d0);
#pragma unroll
    for (int v = 0; v < VPL; ++v) qt[v] = 0.f;
    const float* Wq = a.Wq + (int64_t)i * D * D;
    const float* Wt = a.Wt + (int64_t)i * D * D;
#pragma unroll 4
    for (int k = 0; k < D; ++k) {
      float qk = qs[k];
      float wq[VPL], wt[VPL];
      ldv<VPL>(wq, Wq + k * D + d0);
      ldv<VPL>(wt, Wt + k * D + d0);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        Q[v] = fmaf(qk, wq[v], Q[v]);
        qt[v] = fmaf(qk, wt[v], qt[v]);
      }
    }
#pragma unroll
    for (int v = 0; v < VPL; ++v) Q[v] = fmaxf(Q[v], 0.f);
    const int64_t ib = (int64_t)i * a.B + b;
    stv<VPL>(a.Qr + ib * D + d0, Q);
    stv<VPL>(a.Qt + ib * D + d0, qt);

    const float* w1 = a.gate + ((int64_t)i * 5 + 0) * L;
    const float* b1 = a.gate + ((int64_t)i * 5 + 1) * L;
    const float* o1 = a.gate + ((int64_t)i * 5 + 2) * L;
    const float* o2 = a.gate + ((int64_t)i * 5 + 3) * L;
    const float* ob = a.gate + ((int64_t)i * 5 + 4) * L;
    // ---- pass 1: gated scores ----
#pragma unroll 2
    for (int j = 0; j < len; ++j) {
      const int64_t tok = (int64_t)b * L + j;
      float x[VPL], kk[VPL];
      ldv<VPL>(x, a.X + tok * D + d0);
      ldv<VPL>(kk, a.KV + tok * ldkv + (int64_t)i * 2 * D + d0);
      float pz = 0.f, pa = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        pz = fmaf(qt[v], x[v], pz);
        pa = fmaf(Q[v], kk[v], pa);
      }
      float z = warp_sum(pz);
      float av = group_sum_rt(pa, lph);
      float Z = tanhf(z);                                              // :323
      float dlt = logf(fabsf(tq - __ldg(a.time_list + tok)) + 1.f);    // :339
      float Dk = tanhf(fmaf(dlt, __ldg(w1 + j), __ldg(b1 + j)));       // :343
      float G = __ldg(o1 + j) * Dk + __ldg(o2 + j) * Z + __ldg(ob + j);  // :350
      float gate = sigmoidf_(G);
      float s = (av * gate) / sqrt_dh;                                 // :381-384
      if ((lane % lph) == 0) {
        sc[head * L + j] = s;
        a.AA[(ib * H + head) * L + j] = av;
      }
      if (lane == 0) {
        a.ZZ[ib * L + j] = Z;
        a.DK[ib * L + j] = Dk;
        a.GT[ib * L + j] = gate;
      }
    }
    __syncwarp();
    // ---- softmax over the len valid keys (masked keys get exactly 0: exp(-2^32 - m) == 0) ----
    for (int h = 0; h < H; ++h) {
      float m = -INFINITY;
      for (int j = lane; j < len; j += 32) m = fmaxf(m, sc[h * L + j]);
      m = warp_max(m);
      float s = 0.f;
      for (int j = lane; j < len; j += 32) {
        float e = expf(sc[h * L + j] - m);
        sc[h * L + j] = e;
        s += e;
      }
      s = warp_sum(s);
      for (int j = lane; j < L; j += 32) {
        float p = (j < len) ? sc[h * L + j] / s : 0.f;
        sc[h * L + j] = p;
        a.PA[(ib * H + h) * L + j] = p;
      }
    }
    __syncwarp();
    // ---- pass 2: O = P V ----
    float O[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) O[v] = 0.f;
#pragma unroll 4
    for (int j = 0; j < len; ++j) {
      const int64_t tok = (int64_t)b * L + j;
      float vv[VPL];
      ldv<VPL>(vv, a.KV + tok * ldkv + (int64_t)i * 2 * D + D + d0);
      float p = sc[head * L + j];
#pragma unroll
      for (int v = 0; v < VPL; ++v) O[v] = fmaf(p, vv[v], O[v]);
    }
    // ---- residual + normalize (eps 1e-8)  :447-454, :7-34 ----
    float y[VPL], s1 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) { y[v] = O[v] + q[v]; s1 += y[v]; }
    float mean = warp_sum(s1) / D;
    float s2 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) { y[v] -= mean; s2 = fmaf(y[v], y[v], s2); }
    float var = warp_sum(s2) / D;
    float rstd = 1.f / sqrtf(var + 1e-8f);
    float gm[VPL], bt[VPL], xh[VPL];
    ldv<VPL>(gm, a.ln_gamma + (int64_t)i * D + d0);
    ldv<VPL>(bt, a.ln_beta + (int64_t)i * D + d0);
#pragma unroll
    for (int v = 0; v < VPL; ++v) { xh[v] = y[v] * rstd; q[v] = fmaf(gm[v], xh[v], bt[v]); }
    stv<VPL>(a.XH + ((int64_t)b * N + i) * D + d0, xh);
    if (lane == 0) a.RSTD[(int64_t)b * N + i] = rstd;
    stv<VPL>(a.Qin + ((int64_t)(i + 1) * a.B + b) * D + d0, q);
    __syncwarp();
  }
  // ---- final tf.contrib layer_norm (eps 1e-12) ----
  {
    float s1 = 0.f;
#pragma unroll
    for (int v = 0; v < VPL; ++v) s1 += q[v];
    float mean = warp_sum(s1) / D;
    float s2 = 0.f, y[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) { y[v] = q[v] - mean; s2 = fmaf(y[v], y[v], s2); }
    float var = warp_sum(s2) / D;
    float rstd = rsqrtf(var + 1e-12f);
    float gm[VPL], bt[VPL], xh[VPL], o[VPL];
    ldv<VPL>(gm, a.lnf_gamma + d0);
    ldv<VPL>(bt, a.lnf_beta + d0);
#pragma unroll
    for (int v = 0; v < VPL; ++v) { xh[v] = y[v] * rstd; o[v] = fmaf(gm[v], xh[v], bt[v]); }
    stv<VPL>(a.XHF + (int64_t)b * D + d0, xh);
    if (lane == 0) a.RSTDF[b] = rstd;
    stv<VPL>(a.pred + (int64_t)b * D + d0, o);
  }
}

// layer-norm backward for one row held across a warp
template <int VPL>
__device__ __forceinline__ void ln_bwd_row(const float (&dout)[VPL], const float (&gamma)[VPL],
                                           const float (&xh)[VPL], float rstd, int D, float (&dy)[VPL]) {
  float dxh[VPL], s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    dxh[v] = dout[v] * gamma[v];
    s1 += dxh[v];
    s2 = fmaf(dxh[v], xh[v], s2);
  }
  float m1 = warp_sum(s1) / D, m2 = warp_sum(s2) / D;
#pragma unroll
  for (int v = 0; v < VPL; ++v) dy[v] = (dxh[v] - m1 - xh[v] * m2) * rstd;
}

template <int VPL>
__global__ void __launch_bounds__(HOP_WARPS * 32) hop_bwd_kernel(HopArgs a, HopGradArgs g) {
  constexpr int D = VPL * 32;
  extern __shared__ __align__(16) float sm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * HOP_WARPS + w;
  if (b >= a.B) return;
  const int L = a.L, H = a.H, N = a.N;
  float* qs = sm + (size_t)w * (2 * D + 2 * H * L + 32);
  float* qs2 = qs + D;
  float* ps = qs2 + D;      // [H][L] probabilities
  float* dps = ps + H * L;  // [H][L] dP
  float* dsum = dps + H * L;  // [H]
  const int d0 = lane * VPL;
  const int lph = 32 / H, head = lane / lph;
  const float sqrt_dh = sqrtf((float)(D / H));
  const int len = min(max(a.seq_len[b], 0), L);
  const float tq = a.target_time[b];
  const int ldkv = 2 * N * D;

  float dq[VPL];
  {
    float dp[VPL], gm[VPL], xh[VPL];
    ldv<VPL>(dp, g.dpred + (int64_t)b * D + d0);
    ldv<VPL>(gm, a.lnf_gamma + d0);
    ldv<VPL>(xh, a.XHF + (int64_t)b * D + d0);
    ln_bwd_row<VPL>(dp, gm, xh, a.RSTDF[b], D, dq);
  }
  for (int i = N - 1; i >= 0; --i) {
    const int64_t ib = (int64_t)i * a.B + b;
    stv<VPL>(g.DOUT + ((int64_t)b * N + i) * D + d0, dq);  // for d gamma_i / d beta_i column sums
    float dy[VPL];
    {
      float gm[VPL], xh[VPL];
      ldv<VPL>(gm, a.ln_gamma + (int64_t)i * D + d0);
      ldv<VPL>(xh, a.XH + ((int64_t)b * N + i) * D + d0);
      ln_bwd_row<VPL>(dq, gm, xh, a.RSTD[(int64_t)b * N + i], D, dy);
    }
    float Q[VPL], qt[VPL];
    ldv<VPL>(Q, a.Qr + ib * D + d0);
    ldv<VPL>(qt, a.Qt + ib * D + d0);
    for (int j = lane; j < H * L; j += 32) ps[j] = a.PA[ib * H * L + j];
    // ---- pass A: dP[h][j] = dO_h . V_hj ----
#pragma unroll 2
    for (int j = 0; j < len; ++j) {
      const int64_t tok = (int64_t)b * L + j;
      float vv[VPL];
      ldv<VPL>(vv, a.KV + tok * ldkv + (int64_t)i * 2 * D + D + d0);
      float s = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) s = fmaf(dy[v], vv[v], s);
      s = group_sum_rt(s, lph);
      if ((lane % lph) == 0) dps[head * L + j] = s;
    }
    __syncwarp();
    for (int h = 0; h < H; ++h) {
      float s = 0.f;
      for (int j = lane; j < len; j += 32) s = fmaf(ps[h * L + j], dps[h * L + j], s);
      s = warp_sum(s);
      if (lane == 0) dsum[h] = s;
    }
    __syncwarp();
    const float* w1 = a.gate + ((int64_t)i * 5 + 0) * L;
    const float* o1 = a.gate + ((int64_t)i * 5 + 2) * L;
    const float* o2 = a.gate + ((int64_t)i * 5 + 3) * L;
    float* gb = g.GB + (int64_t)b * (5 * N * L) + (int64_t)i * 5 * L;
    (void)w1;
    float dQ[VPL], dqt[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) { dQ[v] = 0.f; dqt[v] = 0.f; }
    const float my_dsum = dsum[head];
    // ---- pass B ----
#pragma unroll 2
    for (int j = 0; j < len; ++j) {
      const int64_t tok = (int64_t)b * L + j;
      float x[VPL], kk[VPL], vv[VPL];
      ldv<VPL>(x, a.X + tok * D + d0);
      ldv<VPL>(kk, a.KV + tok * ldkv + (int64_t)i * 2 * D + d0);
      ldv<VPL>(vv, a.KV + tok * ldkv + (int64_t)i * 2 * D + D + d0);
      const float gate = __ldg(a.GT + ib * L + j), Z = __ldg(a.ZZ + ib * L + j), Dk = __ldg(a.DK + ib * L + j);
      float dgate = 0.f;
      for (int h = 0; h < H; ++h) {
        float dS = ps[h * L + j] * (dps[h * L + j] - dsum[h]);
        dgate = fmaf(dS, __ldg(a.AA + (ib * H + h) * L + j), dgate);
      }
      dgate /= sqrt_dh;
      const float dG = dgate * gate * (1.f - gate);
      const float dDk = dG * __ldg(o1 + j);
      const float dpre = dDk * (1.f - Dk * Dk);
      if (lane == 0) {
        float dlt = logf(fabsf(tq - __ldg(a.time_list + tok)) + 1.f);
        gb[0 * L + j] = dpre * dlt;  // _time_input_w1
        gb[1 * L + j] = dpre;        // _time_input_b1
        gb[2 * L + j] = dG * Dk;     // time_output_w1
        gb[3 * L + j] = dG * Z;      // time_output_w2
        gb[4 * L + j] = dG;          // time_output_b
      }
      const float dM = dG * __ldg(o2 + j) * (1.f - Z * Z);
      const float p = ps[head * L + j];
      const float dS = p * (dps[head * L + j] - my_dsum);
      const float dA = dS * gate / sqrt_dh;
      float dK[VPL], dV[VPL], dx[VPL];
      ldv<VPL>(dx, g.dX + tok * D + d0);
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        dQ[v] = fmaf(dA, kk[v], dQ[v]);
        dK[v] = (kk[v] > 0.f) ? dA * Q[v] : 0.f;
        dV[v] = (vv[v] > 0.f) ? p * dy[v] : 0.f;
        dqt[v] = fmaf(dM, x[v], dqt[v]);
        dx[v] = fmaf(dM, qt[v], dx[v]);
      }
      stv<VPL>(g.dKV + tok * ldkv + (int64_t)i * 2 * D + d0, dK);
      stv<VPL>(g.dKV + tok * ldkv + (int64_t)i * 2 * D + D + d0, dV);
      stv<VPL>(g.dX + tok * D + d0, dx);
    }
    float dQp[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) dQp[v] = (Q[v] > 0.f) ? dQ[v] : 0.f;
    stv<VPL>(g.DQP + ((int64_t)b * N + i) * D + d0, dQp);
    stv<VPL>(g.DQT + ((int64_t)b * N + i) * D + d0, dqt);
    // dq = dy (residual) + dQpre Wq^T + dqt Wt^T   (transposed copies prepared by the host side)
    stv<VPL>(qs + d0, dQp);
    stv<VPL>(qs2 + d0, dqt);
    __syncwarp();
    const float* WqT = g.WqT + (int64_t)i * D * D;
    const float* WtT = g.WtT + (int64_t)i * D * D;
#pragma unroll
    for (int v = 0; v < VPL; ++v) dq[v] = dy[v];
#pragma unroll 4
    for (int k = 0; k < D; ++k) {
      float s1 = qs[k], s2 = qs2[k];
      float wq[VPL], wt[VPL];
      ldv<VPL>(wq, WqT + k * D + d0);
      ldv<VPL>(wt, WtT + k * D + d0);
#pragma unroll
      for (int v = 0; v < VPL; ++v) dq[v] = fmaf(s1, wq[v], fmaf(s2, wt[v], dq[v]));
    }
    __syncwarp();
  }
  stv<VPL>(g.dq0 + (int64_t)b * D + d0, dq);
}

// =============================================================================================
// CTA-per-sequence variant (the default): the warp-per-sequence kernels above expose every global
// load's latency to a single warp (ncu: 52 % long-scoreboard stalls, 7 warps per SM at B = 1024).
// Here one 128-thread CTA owns a sequence: X (once) and the hop's K, V rows are staged in shared
// memory with cp.async while the query projections run, the score / dP dot products are split into
// 16-float chunks over (chunk, key) work items (lanes across keys, conflict-free 128-bit reads), the
// gate transcendentals are evaluated once per key instead of once per lane, and dX is accumulated in
// shared memory across the hops.  Needs head width % 16 == 0 and the tiles to fit in shared memory;
// otherwise the launchers fall back to the kernels above.  Same saved activations, same outputs.
// =============================================================================================
constexpr int HC_T = 128;

__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// rows [0,len) of D floats, global row stride ldg -> shared rows of stride RS (both 16-byte aligned)
template <int D>
__device__ __forceinline__ void load_rows_async(float* s, int RS, const float* g, int64_t ldg, int len) {
  for (int e = threadIdx.x; e < len * (D / 4); e += HC_T) {
    const int j = e / (D / 4), c = e % (D / 4);
    cp_async16(s + j * RS + c * 4, g + j * ldg + c * 4);
  }
}
__device__ __forceinline__ float dot16(const float* __restrict__ row, const float4 (&v)[4]) {
  const float4* r = reinterpret_cast<const float4*>(row);
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int u = 0; u < 4; u += 2) {
    const float4 a = r[u], b = r[u + 1];
    s0 = fmaf(a.x, v[u].x, s0); s0 = fmaf(a.y, v[u].y, s0); s0 = fmaf(a.z, v[u].z, s0); s0 = fmaf(a.w, v[u].w, s0);
    s1 = fmaf(b.x, v[u + 1].x, s1); s1 = fmaf(b.y, v[u + 1].y, s1); s1 = fmaf(b.z, v[u + 1].z, s1); s1 = fmaf(b.w, v[u + 1].w, s1);
  }
  return s0 + s1;
}

struct HopCtaSmem {
  int RS, X, K, V, dX, vec, mv, part, part2, sc, dps, dA, dM, dsum, bs, total;   // offsets in floats
};
__host__ __device__ inline HopCtaSmem hop_cta_layout(int D, int H, int L, int N, bool bwd) {
  HopCtaSmem o;
  const int C = D / 16, Lp = (L + 3) & ~3;
  o.RS = D + 4;
  int p = 0;
  o.X = p; p += L * o.RS;
  o.K = p; p += L * o.RS;
  o.V = p; p += L * o.RS;
  o.dX = p; p += bwd ? N * (Lp + D) : 0;   // bwd: per hop dM[L] and qt[D] -- dX = sum_i dM_i (x) qt_i is formed once at the end
  o.vec = p; p += 6 * D;          // fwd: q, Q, qt;  bwd: dq, dy, Q, qt, dQpre, dqt
  o.mv = p; p += 2 * HC_T;
  o.part = p; p += C * Lp;
  o.part2 = p; p += C * Lp;
  o.sc = p; p += H * Lp;          // fwd: scores / probabilities;  bwd: probabilities
  o.dps = p; p += bwd ? H * Lp : 0;
  o.dA = p; p += bwd ? H * Lp : 0;
  o.dM = p; p += bwd ? Lp : 0;
  o.dsum = p; p += 32;
  o.bs = p; p += bwd ? (HC_T / 32) * (D / 4) * 8 : 0;   // per-warp column sums of this sequence's dK | dV rows
  o.total = p;
  return o;
}

template <int D>
__global__ void __launch_bounds__(HC_T) hop_fwd_cta_kernel(HopArgs a) {
  constexpr int C = D / 16, PARTS = HC_T / D, KP = D / PARTS, VPL = D / 32;
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int L = a.L, H = a.H, N = a.N, dh = D / H, Lp = (L + 3) & ~3;
  const HopCtaSmem o = hop_cta_layout(D, H, L, N, false);
  const int RS = o.RS;
  float *Xs = sm + o.X, *Ks = sm + o.K, *Vs = sm + o.V, *qv = sm + o.vec, *Qv = qv + D, *qtv = Qv + D, *mv = sm + o.mv,
        *pa = sm + o.part, *pz = sm + o.part2, *sc = sm + o.sc;
  const float sqrt_dh = sqrtf((float)dh);
  const int len = min(max(a.seq_len[b], 0), L);
  const float tq = a.target_time[b];
  const int ldkv = 2 * N * D;
  const int64_t tok0 = (int64_t)b * L;
  load_rows_async<D>(Xs, RS, a.X + tok0 * D, D, len);
  if (t < D) qv[t] = a.Qin[(int64_t)b * D + t];
  for (int i = 0; i < N; ++i) {
    load_rows_async<D>(Ks, RS, a.KV + tok0 * ldkv + (int64_t)i * 2 * D, ldkv, len);
    load_rows_async<D>(Vs, RS, a.KV + tok0 * ldkv + (int64_t)i * 2 * D + D, ldkv, len);
    cp_async_commit();
    __syncthreads();
    const int64_t ib = (int64_t)i * a.B + b;
    {  // Q = relu(q Wq + bq), qt = q Wt   (:249, :320): thread = (output d, k range)
      const int d = t % D, part = t / D;
      const float* Wq = a.Wq + (int64_t)i * D * D + (int64_t)part * KP * D + d;
      const float* Wt = a.Wt + (int64_t)i * D * D + (int64_t)part * KP * D + d;
      float aq = 0.f, at = 0.f;
      // the weight columns come from L2: 2 x WB loads in flight per thread, then their FMAs (same order as a plain loop)
      constexpr int WB = KP < 16 ? KP : 16;
#pragma unroll
      for (int k0 = 0; k0 < KP; k0 += WB) {
        float wq[WB], wt[WB];
#pragma unroll
        for (int j = 0; j < WB; ++j) { wq[j] = __ldg(Wq + (k0 + j) * D); wt[j] = __ldg(Wt + (k0 + j) * D); }
#pragma unroll
        for (int j = 0; j < WB; ++j) {
          const float qk = qv[part * KP + k0 + j];
          aq = fmaf(qk, wq[j], aq);
          at = fmaf(qk, wt[j], at);
        }
      }
      mv[t] = aq;
      mv[HC_T + t] = at;
    }
    __syncthreads();
    if (t < D) {
      float s1 = __ldg(a.bq + (int64_t)i * D + t), s2 = 0.f;
#pragma unroll
      for (int p = 0; p < PARTS; ++p) { s1 += mv[p * D + t]; s2 += mv[HC_T + p * D + t]; }
      s1 = fmaxf(s1, 0.f);
      Qv[t] = s1; qtv[t] = s2;
      a.Qr[ib * D + t] = s1;
      a.Qt[ib * D + t] = s2;
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- partial dot products over 16-float chunks: work item = (chunk c, block of 32 keys) ----
    const int JB = (len + 31) >> 5;
    for (int id = w; id < C * JB; id += HC_T / 32) {
      const int c = id % C, j = (id / C) * 32 + lane;
      float4 qc[4], tc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        qc[u] = *reinterpret_cast<const float4*>(Qv + c * 16 + u * 4);
        tc[u] = *reinterpret_cast<const float4*>(qtv + c * 16 + u * 4);
      }
      if (j < len) {
        pa[c * Lp + j] = dot16(Ks + j * RS + c * 16, qc);
        pz[c * Lp + j] = dot16(Xs + j * RS + c * 16, tc);
      }
    }
    __syncthreads();
    // ---- gate and scores: thread = key ----
    const float* w1 = a.gate + ((int64_t)i * 5 + 0) * L;
    const float* b1 = a.gate + ((int64_t)i * 5 + 1) * L;
    const float* o1 = a.gate + ((int64_t)i * 5 + 2) * L;
    const float* o2 = a.gate + ((int64_t)i * 5 + 3) * L;
    const float* ob = a.gate + ((int64_t)i * 5 + 4) * L;
    for (int j = t; j < len; j += HC_T) {
      float z = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) z += pz[c * Lp + j];
      const float Z = tanhf(z);                                                 // :323
      const float dlt = logf(fabsf(tq - __ldg(a.time_list + tok0 + j)) + 1.f);  // :339
      const float Dk = tanhf(fmaf(dlt, __ldg(w1 + j), __ldg(b1 + j)));          // :343
      const float G = __ldg(o1 + j) * Dk + __ldg(o2 + j) * Z + __ldg(ob + j);   // :350
      const float gate = sigmoidf_(G);
      a.ZZ[ib * L + j] = Z;
      a.DK[ib * L + j] = Dk;
      a.GT[ib * L + j] = gate;
      const int cph = dh / 16;
      for (int h = 0; h < H; ++h) {
        float av = 0.f;
        for (int c = h * cph; c < (h + 1) * cph; ++c) av += pa[c * Lp + j];
        a.AA[(ib * H + h) * L + j] = av;
        sc[h * Lp + j] = (av * gate) / sqrt_dh;                                  // :381-384
      }
    }
    __syncthreads();
    // ---- softmax over the len valid keys (masked keys get exactly 0: exp(-2^32 - m) == 0) ----
    for (int h = w; h < H; h += HC_T / 32) {
      float m = -INFINITY;
      for (int j = lane; j < len; j += 32) m = fmaxf(m, sc[h * Lp + j]);
      m = warp_max(m);
      float s = 0.f;
      for (int j = lane; j < len; j += 32) {
        const float e = expf(sc[h * Lp + j] - m);
        sc[h * Lp + j] = e;
        s += e;
      }
      s = warp_sum(s);
      for (int j = lane; j < L; j += 32) {
        const float p = (j < len) ? sc[h * Lp + j] / s : 0.f;
        if (j < len) sc[h * Lp + j] = p;
        a.PA[(ib * H + h) * L + j] = p;
      }
    }
    __syncthreads();
    {  // ---- O = P V: thread = (d, key subset) ----
      const int d = t % D, part = t / D, h = d / dh;
      float acc = 0.f;
      for (int j = part; j < len; j += PARTS) acc = fmaf(sc[h * Lp + j], Vs[j * RS + d], acc);
      mv[t] = acc;
    }
    __syncthreads();
    if (w == 0) {  // ---- residual + normalize (eps 1e-8)  :447-454, :7-34 ----
      const int d0 = lane * VPL;
      float y[VPL], s1 = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) {
        float ov = 0.f;
#pragma unroll
        for (int p = 0; p < PARTS; ++p) ov += mv[p * D + d0 + v];
        y[v] = ov + qv[d0 + v];
        s1 += y[v];
      }
      const float mean = warp_sum(s1) / D;
      float s2 = 0.f;
#pragma unroll
      for (int v = 0; v < VPL; ++v) { y[v] -= mean; s2 = fmaf(y[v], y[v], s2); }
      const float var = warp_sum(s2) / D;
      const float rstd = 1.f / sqrtf(var + 1e-8f);
      float gm[VPL], bt[VPL], xh[VPL], qn[VPL];
      ldv<VPL>(gm, a.ln_gamma + (int64_t)i * D + d0);
      ldv<VPL>(bt, a.ln_beta + (int64_t)i * D + d0);
#pragma unroll
      for (int v = 0; v < VPL; ++v) { xh[v] = y[v] * rstd; qn[v] = fmaf(gm[v], xh[v], bt[v]); qv[d0 + v] = qn[v]; }
      stv<VPL>(a.XH + ((int64_t)b * N + i) * D + d0, xh);
      if (lane == 0) a.RSTD[(int64_t)b * N + i] = rstd;
      stv<VPL>(a.Qin + ((int64_t)(i + 1) * a.B + b) * D + d0, qn);
    }
    __syncthreads();
  }
  if (w == 0) {  // ---- final tf.contrib layer_norm (eps 1e-12) ----
    const int d0 = lane * VPL;
    float s1 = 0.f, q[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) { q[v] = qv[d0 + v]; s1 += q[v]; }
    const float mean = warp_sum(s1) / D;
    float s2 = 0.f, y[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) { y[v] = q[v] - mean; s2 = fmaf(y[v], y[v], s2); }
    const float var = warp_sum(s2) / D;
    const float rstd = rsqrtf(var + 1e-12f);
    float gm[VPL], bt[VPL], xh[VPL], ov[VPL];
    ldv<VPL>(gm, a.lnf_gamma + d0);
    ldv<VPL>(bt, a.lnf_beta + d0);
#pragma unroll
    for (int v = 0; v < VPL; ++v) { xh[v] = y[v] * rstd; ov[v] = fmaf(gm[v], xh[v], bt[v]); }
    stv<VPL>(a.XHF + (int64_t)b * D + d0, xh);
    if (lane == 0) a.RSTDF[b] = rstd;
    stv<VPL>(a.pred + (int64_t)b * D + d0, ov);
  }
}

template <int D>
__global__ void __launch_bounds__(HC_T) hop_bwd_cta_kernel(HopArgs a, HopGradArgs g) {
  constexpr int C = D / 16, PARTS = HC_T / D, KP = D / PARTS, VPL = D / 32;
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x, t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int L = a.L, H = a.H, N = a.N, dh = D / H, Lp = (L + 3) & ~3;
  const HopCtaSmem o = hop_cta_layout(D, H, L, N, true);
  const int RS = o.RS;
  float *Xs = sm + o.X, *Ks = sm + o.K, *Vs = sm + o.V, *dMs = sm + o.dX, *qts = dMs + N * Lp, *dqv = sm + o.vec, *dyv = dqv + D, *Qv = dyv + D,
        *qtv = Qv + D, *dQpv = qtv + D, *dqtv = dQpv + D, *mv = sm + o.mv, *part = sm + o.part, *ps = sm + o.sc,
        *dps = sm + o.dps, *dA = sm + o.dA, *dMv = sm + o.dM, *dsum = sm + o.dsum;
  const float sqrt_dh = sqrtf((float)dh);
  const int len = min(max(a.seq_len[b], 0), L);
  const float tq = a.target_time[b];
  const int ldkv = 2 * N * D;
  const int64_t tok0 = (int64_t)b * L;
  load_rows_async<D>(Xs, RS, a.X + tok0 * D, D, len);
  if (w == 0) {   // backward of the final layer norm
    const int d0 = lane * VPL;
    float dp[VPL], gm[VPL], xh[VPL], dq[VPL];
    ldv<VPL>(dp, g.dpred + (int64_t)b * D + d0);
    ldv<VPL>(gm, a.lnf_gamma + d0);
    ldv<VPL>(xh, a.XHF + (int64_t)b * D + d0);
    ln_bwd_row<VPL>(dp, gm, xh, a.RSTDF[b], D, dq);
    stv<VPL>(dqv + d0, dq);
  }
  for (int i = N - 1; i >= 0; --i) {
    const int64_t ib = (int64_t)i * a.B + b;
    __syncthreads();     // dqv complete; K/V tiles of the previous hop no longer read
    load_rows_async<D>(Ks, RS, a.KV + tok0 * ldkv + (int64_t)i * 2 * D, ldkv, len);
    load_rows_async<D>(Vs, RS, a.KV + tok0 * ldkv + (int64_t)i * 2 * D + D, ldkv, len);
    cp_async_commit();
    if (w == 0) {
      const int d0 = lane * VPL;
      float dq[VPL], gm[VPL], xh[VPL], dy[VPL];
      ldv<VPL>(dq, dqv + d0);
      stv<VPL>(g.DOUT + ((int64_t)b * N + i) * D + d0, dq);   // for d gamma_i / d beta_i column sums
      ldv<VPL>(gm, a.ln_gamma + (int64_t)i * D + d0);
      ldv<VPL>(xh, a.XH + ((int64_t)b * N + i) * D + d0);
      ln_bwd_row<VPL>(dq, gm, xh, a.RSTD[(int64_t)b * N + i], D, dy);
      stv<VPL>(dyv + d0, dy);
    } else if (w == 1) {
      for (int d = lane; d < D; d += 32) { Qv[d] = a.Qr[ib * D + d]; qtv[d] = a.Qt[ib * D + d]; }
    } else {
      for (int j = t - 64; j < H * L; j += HC_T - 64) ps[(j / L) * Lp + (j % L)] = a.PA[ib * H * L + j];
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- pass A: dP[h][j] = dO_h . V_hj  (partial dots over 16-float chunks) ----
    const int JB = (len + 31) >> 5;
    for (int id = w; id < C * JB; id += HC_T / 32) {
      const int c = id % C, j = (id / C) * 32 + lane;
      float4 yc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) yc[u] = *reinterpret_cast<const float4*>(dyv + c * 16 + u * 4);
      if (j < len) part[c * Lp + j] = dot16(Vs + j * RS + c * 16, yc);
    }
    __syncthreads();
    const int cph = dh / 16;
    for (int j = t; j < len; j += HC_T)
      for (int h = 0; h < H; ++h) {
        float s = 0.f;
        for (int c = h * cph; c < (h + 1) * cph; ++c) s += part[c * Lp + j];
        dps[h * Lp + j] = s;
      }
    __syncthreads();
    for (int h = w; h < H; h += HC_T / 32) {
      float s = 0.f;
      for (int j = lane; j < len; j += 32) s = fmaf(ps[h * Lp + j], dps[h * Lp + j], s);
      s = warp_sum(s);
      if (lane == 0) dsum[h] = s;
    }
    __syncthreads();
    // ---- per key: softmax / gate backward, gate-parameter gradients ----
    const float* o1 = a.gate + ((int64_t)i * 5 + 2) * L;
    const float* o2 = a.gate + ((int64_t)i * 5 + 3) * L;
    float* gb = g.GB + (int64_t)b * (5 * N * L) + (int64_t)i * 5 * L;
    for (int j = t; j < len; j += HC_T) {
      const float gate = __ldg(a.GT + ib * L + j), Z = __ldg(a.ZZ + ib * L + j), Dk = __ldg(a.DK + ib * L + j);
      float dgate = 0.f;
      for (int h = 0; h < H; ++h) {
        const float dS = ps[h * Lp + j] * (dps[h * Lp + j] - dsum[h]);
        dgate = fmaf(dS, __ldg(a.AA + (ib * H + h) * L + j), dgate);
        dA[h * Lp + j] = dS * gate / sqrt_dh;
      }
      dgate /= sqrt_dh;
      const float dG = dgate * gate * (1.f - gate);
      const float dDk = dG * __ldg(o1 + j);
      const float dpre = dDk * (1.f - Dk * Dk);
      const float dlt = logf(fabsf(tq - __ldg(a.time_list + tok0 + j)) + 1.f);
      gb[0 * L + j] = dpre * dlt;  // _time_input_w1
      gb[1 * L + j] = dpre;        // _time_input_b1
      gb[2 * L + j] = dG * Dk;     // time_output_w1
      gb[3 * L + j] = dG * Z;      // time_output_w2
      gb[4 * L + j] = dG;          // time_output_b
      dMv[j] = dG * __ldg(o2 + j) * (1.f - Z * Z);
    }
    __syncthreads();
    {  // dQ[d] = sum_j dA[h][j] K[j][d],  dqt[d] = sum_j dM[j] X[j][d]: thread = (d, key subset)
      const int d = t % D, pt = t / D, h = d / dh;
      float aq = 0.f, at = 0.f;
      for (int j = pt; j < len; j += PARTS) {
        aq = fmaf(dA[h * Lp + j], Ks[j * RS + d], aq);
        at = fmaf(dMv[j], Xs[j * RS + d], at);
      }
      mv[t] = aq;
      mv[HC_T + t] = at;
    }
    // dK, dV (relu-masked, written once; masked keys get zeros); this hop's rank-1 share of dX is kept as (dM, qt)
    if (t < D) qts[i * D + t] = qtv[t];
    for (int j = t; j < len; j += HC_T) dMs[i * Lp + j] = dMv[j];
    float4 sk = make_float4(0.f, 0.f, 0.f, 0.f), sv = sk;   // this thread's share of the column sums (bias gradient)
    for (int e = t; e < L * (D / 4); e += HC_T) {
      const int j = e / (D / 4), d = (e % (D / 4)) * 4, h = d / dh;
      float4 dk = make_float4(0.f, 0.f, 0.f, 0.f), dv = dk;
      if (j < len) {
        const float4 kk = *reinterpret_cast<const float4*>(Ks + j * RS + d);
        const float4 vv = *reinterpret_cast<const float4*>(Vs + j * RS + d);
        const float4 q4 = *reinterpret_cast<const float4*>(Qv + d);
        const float4 y4 = *reinterpret_cast<const float4*>(dyv + d);
        const float da = dA[h * Lp + j], p = ps[h * Lp + j];
        dk.x = kk.x > 0.f ? da * q4.x : 0.f; dk.y = kk.y > 0.f ? da * q4.y : 0.f;
        dk.z = kk.z > 0.f ? da * q4.z : 0.f; dk.w = kk.w > 0.f ? da * q4.w : 0.f;
        dv.x = vv.x > 0.f ? p * y4.x : 0.f; dv.y = vv.y > 0.f ? p * y4.y : 0.f;
        dv.z = vv.z > 0.f ? p * y4.z : 0.f; dv.w = vv.w > 0.f ? p * y4.w : 0.f;
      }
      float* dst = g.dKV + (tok0 + j) * ldkv + (int64_t)i * 2 * D + d;
      __stcs(reinterpret_cast<float4*>(dst), dk);
      __stcs(reinterpret_cast<float4*>(dst + D), dv);
      sk.x += dk.x; sk.y += dk.y; sk.z += dk.z; sk.w += dk.w;
      sv.x += dv.x; sv.y += dv.y; sv.z += dv.z; sv.w += dv.w;
    }
    {  // a thread always lands on the same 4 columns (HC_T is a multiple of D/4): lanes, then warps, in a fixed order
      float* bsw = sm + o.bs;
#pragma unroll
      for (int off = D / 4; off < 32; off <<= 1) {
        sk.x += __shfl_xor_sync(0xffffffffu, sk.x, off); sk.y += __shfl_xor_sync(0xffffffffu, sk.y, off);
        sk.z += __shfl_xor_sync(0xffffffffu, sk.z, off); sk.w += __shfl_xor_sync(0xffffffffu, sk.w, off);
        sv.x += __shfl_xor_sync(0xffffffffu, sv.x, off); sv.y += __shfl_xor_sync(0xffffffffu, sv.y, off);
        sv.z += __shfl_xor_sync(0xffffffffu, sv.z, off); sv.w += __shfl_xor_sync(0xffffffffu, sv.w, off);
      }
      if (lane < D / 4) {
        float4* dstb = reinterpret_cast<float4*>(bsw + ((w * (D / 4)) + (t % (D / 4))) * 8);
        dstb[0] = sk; dstb[1] = sv;
      }
    }
    __syncthreads();
    for (int c = t; c < 2 * D; c += HC_T) {
      const int half = c / D, dc = c % D;
      float s = 0.f;
#pragma unroll
      for (int ww = 0; ww < HC_T / 32; ++ww) s += sm[o.bs + ((ww * (D / 4)) + dc / 4) * 8 + half * 4 + (dc & 3)];
      g.BKV[((int64_t)b * N + i) * 2 * D + c] = s;
    }
    if (t < D) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int p = 0; p < PARTS; ++p) { s1 += mv[p * D + t]; s2 += mv[HC_T + p * D + t]; }
      s1 = Qv[t] > 0.f ? s1 : 0.f;
      dQpv[t] = s1; dqtv[t] = s2;
      g.DQP[((int64_t)b * N + i) * D + t] = s1;
      g.DQT[((int64_t)b * N + i) * D + t] = s2;
    }
    __syncthreads();
    {  // dq = dy (residual) + dQpre Wq^T + dqt Wt^T   (transposed copies prepared by the host side)
      const int d = t % D, pt = t / D;
      const float* WqT = g.WqT + (int64_t)i * D * D + (int64_t)pt * KP * D + d;
      const float* WtT = g.WtT + (int64_t)i * D * D + (int64_t)pt * KP * D + d;
      float acc = 0.f;
      constexpr int WB = KP < 16 ? KP : 16;
#pragma unroll
      for (int k0 = 0; k0 < KP; k0 += WB) {
        float wq[WB], wt[WB];
#pragma unroll
        for (int j = 0; j < WB; ++j) { wq[j] = __ldg(WqT + (k0 + j) * D); wt[j] = __ldg(WtT + (k0 + j) * D); }
#pragma unroll
        for (int j = 0; j < WB; ++j)
          acc = fmaf(dQpv[pt * KP + k0 + j], wq[j], fmaf(dqtv[pt * KP + k0 + j], wt[j], acc));
      }
      mv[t] = acc;
    }
    __syncthreads();
    if (t < D) {
      float s = dyv[t];
#pragma unroll
      for (int p = 0; p < PARTS; ++p) s += mv[p * D + t];
      dqv[t] = s;
    }
  }
  __syncthreads();
  if (t < D) g.dq0[(int64_t)b * D + t] = dqv[t];
  for (int e = t; e < len * (D / 4); e += HC_T) {
    const int j = e / (D / 4), d = (e % (D / 4)) * 4;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = N - 1; i >= 0; --i) {     // the order the hops were walked in
      const float dm = dMs[i * Lp + j];
      const float4 t4 = *reinterpret_cast<const float4*>(qts + i * D + d);
      x.x = fmaf(dm, t4.x, x.x); x.y = fmaf(dm, t4.y, x.y); x.z = fmaf(dm, t4.z, x.z); x.w = fmaf(dm, t4.w, x.w);
    }
    *reinterpret_cast<float4*>(g.dX + (tok0 + j) * D + d) = x;
  }
}

static bool hop_cta_ok(int D, int H, int L, int N, bool bwd, size_t* smem) {
  if (D % H != 0 || (D / H) % 16 != 0) return false;
  *smem = (size_t)hop_cta_layout(D, H, L, N, bwd).total * sizeof(float);
  return *smem <= 220 * 1024;
}
// hop_backward writes every dKV element itself when it takes the CTA-per-sequence path
bool hop_backward_writes_all_dkv(int D, int H, int L, int N) {
  size_t smem;
  return hop_cta_ok(D, H, L, N, true, &smem);
}

size_t hop_smem_bytes(int D, int H, int L, bool bwd) {
  size_t per_warp = bwd ? (size_t)(2 * D + 2 * H * L + 32) : (size_t)(D + 2 * H * L);
  return per_warp * HOP_WARPS * sizeof(float);
}

int hop_forward(const HopArgs& a, cudaStream_t st) {
  if (a.H < 1 || a.H > 32 || (32 % a.H) != 0 || (a.D % a.H) != 0)
    return set_error(-1, "attention: num_heads=%d must divide 32 and num_units", a.H);
  size_t csmem;
  if (hop_cta_ok(a.D, a.H, a.L, a.N, false, &csmem)) {
#define HOP_FWD_CTA(DD)                                                                                               \
  do {                                                                                                                \
    MTAM_CUDA_CHECK(cudaFuncSetAttribute(hop_fwd_cta_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem)); \
    hop_fwd_cta_kernel<DD><<<a.B, HC_T, csmem, st>>>(a);                                                              \
  } while (0)
    switch (a.D) {
      case 32: HOP_FWD_CTA(32); break;
      case 64: HOP_FWD_CTA(64); break;
      case 128: HOP_FWD_CTA(128); break;
      default: return set_error(-1, "attention: num_units=%d not supported (32, 64, 128)", a.D);
    }
#undef HOP_FWD_CTA
    MTAM_LAUNCH_CHECK();
    return 0;
  }
  size_t smem = hop_smem_bytes(a.D, a.H, a.L, false);
  int blocks = cdiv(a.B, HOP_WARPS);
#define HOP_FWD(V)                                                                                            \
  do {                                                                                                        \
    MTAM_CUDA_CHECK(cudaFuncSetAttribute(hop_fwd_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    hop_fwd_kernel<V><<<blocks, HOP_WARPS * 32, smem, st>>>(a);                                               \
  } while (0)
  switch (a.D) {
    case 32: HOP_FWD(1); break;
    case 64: HOP_FWD(2); break;
    case 128: HOP_FWD(4); break;
    default: return set_error(-1, "attention: num_units=%d not supported (32, 64, 128)", a.D);
  }
#undef HOP_FWD
  MTAM_LAUNCH_CHECK();
  return 0;
}

int hop_backward(const HopArgs& a, const HopGradArgs& g, cudaStream_t st) {
  size_t csmem;
  if (hop_cta_ok(a.D, a.H, a.L, a.N, true, &csmem)) {
#define HOP_BWD_CTA(DD)                                                                                               \
  do {                                                                                                                \
    MTAM_CUDA_CHECK(cudaFuncSetAttribute(hop_bwd_cta_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem)); \
    hop_bwd_cta_kernel<DD><<<a.B, HC_T, csmem, st>>>(a, g);                                                           \
  } while (0)
    switch (a.D) {
      case 32: HOP_BWD_CTA(32); break;
      case 64: HOP_BWD_CTA(64); break;
      case 128: HOP_BWD_CTA(128); break;
      default: return set_error(-1, "attention: num_units=%d not supported (32, 64, 128)", a.D);
    }
#undef HOP_BWD_CTA
    MTAM_LAUNCH_CHECK();
    return 0;
  }
  size_t smem = hop_smem_bytes(a.D, a.H, a.L, true);
  int blocks = cdiv(a.B, HOP_WARPS);
#define HOP_BWD(V)                                                                                            \
  do {                                                                                                        \
    MTAM_CUDA_CHECK(cudaFuncSetAttribute(hop_bwd_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    hop_bwd_kernel<V><<<blocks, HOP_WARPS * 32, smem, st>>>(a, g);                                            \
  } while (0)
  switch (a.D) {
    case 32: HOP_BWD(1); break;
    case 64: HOP_BWD(2); break;
    case 128: HOP_BWD(4); break;
    default: return set_error(-1, "attention: num_units=%d not supported (32, 64, 128)", a.D);
  }
#undef HOP_BWD
  MTAM_LAUNCH_CHECK();
  return 0;
}

// dst[n][k] = src[k][n] for `count` DxD matrices
__global__ void transpose_dd_kernel(const float* __restrict__ src, float* __restrict__ dst, int D) {
  __shared__ float tile[32][33];
  const float* s = src + (int64_t)blockIdx.z * D * D;
  float* d = dst + (int64_t)blockIdx.z * D * D;
  int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
  for (int r = 0; r < 32; r += 8)
    if (x < D && y + r < D) tile[threadIdx.y + r][threadIdx.x] = s[(y + r) * D + x];
  __syncthreads();
  x = blockIdx.y * 32 + threadIdx.x;
  y = blockIdx.x * 32 + threadIdx.y;
  for (int r = 0; r < 32; r += 8)
    if (x < D && y + r < D) d[(y + r) * D + x] = tile[threadIdx.x][threadIdx.y + r];
}
int transpose_dd(const float* src, float* dst, int D, int count, cudaStream_t st) {
  dim3 grid(cdiv(D, 32), cdiv(D, 32), count);
  transpose_dd_kernel<<<grid, dim3(32, 8), 0, st>>>(src, dst, D);
  MTAM_LAUNCH_CHECK();
  return 0;
}

}  // namespace mtam
