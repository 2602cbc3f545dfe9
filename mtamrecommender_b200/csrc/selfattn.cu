// Self-attention encoders (Tq = Tk = L): PISTRec / Time_Aware_Self_Attention_Model (time-aware
// gate), Ti_Self_Attention_Model (additive log-interval bias) and Self_Attention_Model (plain).
// One CTA owns one sequence per block: Q/K/V/keys are staged once in shared memory, the [L,L]
// score matrix of one head lives in shared memory, nothing of size [B*H,L,L] is re-read from HBM
// inside the block.  The Q|K|V projections are one [T,D]x[D,3D] GEMM per block.
//
// Reference: Time_Aware_Attention.self_attention / Tiself_attention  Model/Modules/time_aware_attention.py:459-522
//            time_aware_multihead_attention :215-456, TiSAS_multihead_attention :73-214
//            Attention.multihead_attention / self_attention  Model/Modules/multihead_attention.py:71-221
//            PISTRec_model.py:38-74, attention_baseline_models.py:33-84 (gather @ seq_len-1, layer_norm)
// Attention dropout of the plain and TiSAS variants (multihead_attention.py:179, time_aware_attention.py:198; the
// reference passes is_training=True unconditionally, so it is active in evaluation too): the softmax weights are
// multiplied by keep/(1-rate) before the PV product.  TF's Philox stream cannot be reproduced, so the keep mask comes
// from a counter-based hash of (seed, call counter, block, element) -- sa_keep() below -- which the parity tests
// restate on the host to hand the oracle the identical mask.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "selfattn.h"

namespace mtam {

static const char* kSaGateLive[5] = {"_time_input_w1", "_time_input_b1", "time_output_w1", "time_output_w2",
                                     "time_output_b"};
static bool time_aware(int kind) { return kind == MTAM_KIND_PISTREC || kind == MTAM_KIND_TA_SASREC; }
static size_t a4(size_t x) { return (x + 3) / 4 * 4; }

int sa_build_layout(const mtam_config& c, size_t& o, std::vector<ParamDesc>& P, SaLayout& s) {
  const int D = c.D, L = c.L, N = c.N;
  auto take = [&](size_t n) { size_t r = o; o = a4(o + n); return r; };
  s.W3 = take((size_t)N * D * 3 * D);
  s.b3 = take((size_t)N * 3 * D);
  if (time_aware(c.kind)) {
    s.Wt = take((size_t)N * D * D);
    s.gate = take((size_t)N * 5 * L * L);
    s.gate_dead = take((size_t)N * L * L);
  }
  s.lnb = take((size_t)N * D);
  s.lng = take((size_t)N * D);
  s.lnfb = take(D);
  s.lnfg = take(D);
  const char* dn[3] = {"dense", "dense_1", "dense_2"};
  for (int i = 0; i < N; ++i) {
    const std::string b = "UserHistoryEncoder/encoder/num_blocks_" + std::to_string(i) + "/";
    for (int k = 0; k < 3; ++k) {
      P.push_back({b + dn[k] + "/kernel", D, D, 2, 3 * D, s.W3 + (size_t)i * D * 3 * D + (size_t)k * D, 0});
      P.push_back({b + dn[k] + "/bias", 1, D, 1, D, s.b3 + (size_t)i * 3 * D + (size_t)k * D, 0});
    }
    if (time_aware(c.kind)) {
      P.push_back({b + "self_attention/_time_input_w", D, D, 2, D, s.Wt + (size_t)i * D * D, 0});
      for (int k = 0; k < 5; ++k)
        P.push_back({b + "self_attention/" + kSaGateLive[k], L, L, 2, L, s.gate + ((size_t)i * 5 + k) * L * L, 0});
      P.push_back({b + "self_attention/time_output_w3", L, L, 2, L, s.gate_dead + (size_t)i * L * L, MTAM_PARAM_DEAD});
    }
    P.push_back({b + "self_attention/ln/beta", 1, D, 1, D, s.lnb + (size_t)i * D, 0});
    P.push_back({b + "self_attention/ln/gamma", 1, D, 1, D, s.lng + (size_t)i * D, 0});
  }
  P.push_back({"UserHistoryEncoder/LayerNorm/beta", 1, D, 1, D, s.lnfb, 0});
  P.push_back({"UserHistoryEncoder/LayerNorm/gamma", 1, D, 1, D, s.lnfg, 0});
  return 0;
}

struct SaWs {
  float *ENCo, *QKV, *ET, *P, *A, *Z, *DK, *GT, *XH, *RSTD, *dEncA, *dEncB, *dQKV, *dET, *GB;
  size_t total;
};
static SaWs sa_plan(const mtam_config& c, void* base, size_t cap) {
  const int64_t B = c.max_batch, L = c.L, D = c.D, N = c.N, H = c.H, T = B * L;
  const bool ta = time_aware(c.kind);
  Bump b(base, cap);
  SaWs w{};
  w.ENCo = b.take<float>(N * T * D);
  w.QKV = b.take<float>(N * T * 3 * D);
  w.P = b.take<float>(N * B * H * L * L);
  w.XH = b.take<float>(N * T * D);
  w.RSTD = b.take<float>(N * T);
  w.dEncA = b.take<float>(T * D);
  w.dEncB = b.take<float>(T * D);
  w.dQKV = b.take<float>(T * 3 * D);
  if (ta) {
    w.ET = b.take<float>(N * T * D);
    w.A = b.take<float>(N * B * H * L * L);
    w.Z = b.take<float>(N * B * L * L);
    w.DK = b.take<float>(N * B * L * L);
    w.GT = b.take<float>(N * B * L * L);
    w.dET = b.take<float>(T * D);
    w.GB = b.take<float>(B * 5 * L * L);
  }
  w.total = b.off + 256;
  return w;
}
size_t sa_workspace_bytes(const mtam_config& c) { return sa_plan(c, nullptr, 0).total; }

// -------------------------------------------------------------------------------------------------
// keep decision of attention dropout for element e = ((b*H + h)*L + i)*L + j of block `blk` in forward call `counter`
__host__ __device__ __forceinline__ bool sa_keep(uint32_t seed, uint32_t counter, uint32_t blk, uint32_t e, uint32_t thr24) {
  uint32_t x = e * 0x9E3779B1u ^ (seed + 0x85EBCA77u * (counter * 64u + blk));
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;   // murmur3 finaliser
  return (x >> 8) >= thr24;                                                       // P(keep) = 1 - thr24 / 2^24
}

struct SaBlockArgs {
  int B, L, D, H, mode;  // mode: 0 plain, 1 time-aware gate, 2 tisas additive interval
  uint32_t drop_thr24;   // round(rate * 2^24); 0 = no dropout
  float drop_scale;      // 1 / (1 - rate)
  uint32_t drop_seed, blk;
  const uint32_t* drop_counter;
  const int32_t* seq_len;
  const float* time_list;  // [B,L]
  const float* enc;        // [T,D]  block input (raw queries = raw keys)
  const float* QKV;        // [T,3D] relu projections
  const float* ET;         // [T,D]  enc * Wt        (mode 1)
  const float* gate;       // [5][L*L]               (mode 1)
  const float* ln_gamma;   // [D]
  const float* ln_beta;    // [D]
  float* P;                // [B][H][L][L]
  float* A;                // [B][H][L][L]           (mode 1)
  float* Z; float* DK; float* GT;  // [B][L][L]      (mode 1)
  float* XH;               // [T,D]
  float* RSTD;             // [T]
  float* out;              // [T,D]  block output
  // backward
  const float* dOut;       // [T,D]
  float* dEnc;             // [T,D]  written: residual part (+ dM^T ET for mode 1)
  float* dQKV;             // [T,3D] relu-masked, pre-zeroed
  float* dET;              // [T,D]  pre-zeroed (mode 1)
  float* GB;               // [B][5][L*L] pre-zeroed (mode 1)
};

constexpr int SA_THREADS = 256;

static size_t sa_smem_floats(int L, int D, int mode, bool bwd) {
  size_t row = (size_t)L * (D + 1), sq = (size_t)L * (L + 1);
  if (!bwd) return (mode == 1 ? 5 : 4) * row + 2 * sq;
  return (mode == 1 ? 6 : 5) * row + 3 * sq;
}

__device__ __forceinline__ void sa_load_rows(float* dst, const float* __restrict__ src, int ld, int L, int D) {
  for (int i = threadIdx.x; i < L * D; i += SA_THREADS) {
    int r = i / D, c = i % D;
    dst[r * (D + 1) + c] = src[(int64_t)r * ld + c];
  }
}

__global__ void __launch_bounds__(SA_THREADS) sa_block_fwd_kernel(SaBlockArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int L = a.L, D = a.D, H = a.H, dh = D / H, RS = D + 1, SS = L + 1;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = SA_THREADS / 32;
  float* sQ = sm;
  float* sK = sQ + L * RS;
  float* sV = sK + L * RS;
  float* sE = sV + L * RS;
  float* sET = sE + L * RS;                             // mode 1 only
  float* sS = (a.mode == 1) ? sET + L * RS : sET;
  float* sG = sS + L * SS;
  const int len = min(max(a.seq_len[b], 0), L);
  const int64_t t0 = (int64_t)b * L;
  sa_load_rows(sQ, a.QKV + t0 * 3 * D, 3 * D, L, D);
  sa_load_rows(sK, a.QKV + t0 * 3 * D + D, 3 * D, L, D);
  sa_load_rows(sV, a.QKV + t0 * 3 * D + 2 * D, 3 * D, L, D);
  sa_load_rows(sE, a.enc + t0 * D, D, L, D);
  if (a.mode == 1) sa_load_rows(sET, a.ET + t0 * D, D, L, D);
  __syncthreads();
  const float inv = 1.0f;  // scores are divided by sqrt(dh) below, after gating, as the reference does
  (void)inv;
  const float sqrt_dh = sqrtf((float)dh);
  const uint32_t drop_ctr = a.drop_thr24 ? *a.drop_counter : 0u;
  // ---- gate / interval term, shared by all heads ----
  if (a.mode != 0) {
    const float* w1 = a.gate;
    for (int p = tid; p < len * len; p += SA_THREADS) {
      int i = p / len, j = p % len;
      float dlt = logf(fabsf(a.time_list[t0 + i] - a.time_list[t0 + j]) + 1.f);   // :339 / :110
      if (a.mode == 2) { sG[i * SS + j] = dlt; continue; }
      float z = 0.f;
      for (int k = 0; k < D; ++k) z = fmaf(sET[i * RS + k], sE[j * RS + k], z);
      float Z = tanhf(z);                                                           // :320-323
      int q = i * L + j;
      float Dk = tanhf(fmaf(dlt, w1[q], w1[L * L + q]));                            // :343
      float G = w1[2 * L * L + q] * Dk + w1[3 * L * L + q] * Z + w1[4 * L * L + q]; // :350
      float g = sigmoidf_(G);
      sG[i * SS + j] = g;
      int64_t o = ((int64_t)b * L + i) * L + j;
      a.Z[o] = Z; a.DK[o] = Dk; a.GT[o] = g;
    }
  }
  __syncthreads();
  for (int h = 0; h < H; ++h) {
    const int c0 = h * dh;
    for (int p = tid; p < len * len; p += SA_THREADS) {
      int i = p / len, j = p % len;
      float s = 0.f;
      for (int k = 0; k < dh; ++k) s = fmaf(sQ[i * RS + c0 + k], sK[j * RS + c0 + k], s);
      if (a.mode == 1) {
        a.A[(((int64_t)b * H + h) * L + i) * L + j] = s;
        s = s * sG[i * SS + j];                       // :381
      } else if (a.mode == 2) {
        s = s + sG[i * SS + j];                       // :139
      }
      sS[i * SS + j] = s / sqrt_dh;                   // :384 / :142 / multihead_attention.py:116
    }
    __syncthreads();
    // softmax over valid keys, query mask (rows >= len are exactly 0)
    for (int i = warp; i < L; i += nwarp) {
      float* Pg = a.P + (((int64_t)b * H + h) * L + i) * L;
      if (i >= len) {
        for (int j = lane; j < L; j += 32) { sS[i * SS + j] = 0.f; Pg[j] = 0.f; }
        continue;
      }
      float m = -INFINITY;
      for (int j = lane; j < len; j += 32) m = fmaxf(m, sS[i * SS + j]);
      m = warp_max(m);
      float s = 0.f;
      for (int j = lane; j < len; j += 32) { float e = expf(sS[i * SS + j] - m); sS[i * SS + j] = e; s += e; }
      s = warp_sum(s);
      for (int j = lane; j < L; j += 32) {
        float p = (j < len) ? sS[i * SS + j] / s : 0.f;
        Pg[j] = p;                                       // the softmax weights themselves are kept for the backward pass
        if (a.drop_thr24)
          p = sa_keep(a.drop_seed, drop_ctr, a.blk, (uint32_t)((((int64_t)b * H + h) * L + i) * L + j), a.drop_thr24)
                  ? p * a.drop_scale : 0.f;
        sS[i * SS + j] = p;
      }
    }
    __syncthreads();
    // O_h = P V_h, written over Q_h (no longer needed)
    for (int p = tid; p < L * dh; p += SA_THREADS) {
      int i = p / dh, d = p % dh;
      float o = 0.f;
      if (i < len)
        for (int j = 0; j < len; ++j) o = fmaf(sS[i * SS + j], sV[j * RS + c0 + d], o);
      sQ[i * RS + c0 + d] = o;
    }
    __syncthreads();
  }
  // residual + normalize (eps 1e-8), one warp per position
  for (int i = warp; i < L; i += nwarp) {
    float s1 = 0.f;
    for (int d = lane; d < D; d += 32) { float y = sQ[i * RS + d] + sE[i * RS + d]; sQ[i * RS + d] = y; s1 += y; }
    float mean = warp_sum(s1) / D;
    float s2 = 0.f;
    for (int d = lane; d < D; d += 32) { float y = sQ[i * RS + d] - mean; s2 = fmaf(y, y, s2); }
    float rstd = 1.f / sqrtf(warp_sum(s2) / D + 1e-8f);
    for (int d = lane; d < D; d += 32) {
      float xh = (sQ[i * RS + d] - mean) * rstd;
      a.XH[(t0 + i) * D + d] = xh;
      a.out[(t0 + i) * D + d] = fmaf(a.ln_gamma[d], xh, a.ln_beta[d]);
    }
    if (lane == 0) a.RSTD[t0 + i] = rstd;
  }
}

__global__ void __launch_bounds__(SA_THREADS) sa_block_bwd_kernel(SaBlockArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int L = a.L, D = a.D, H = a.H, dh = D / H, RS = D + 1, SS = L + 1;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = SA_THREADS / 32;
  float* sdO = sm;
  float* sQ = sdO + L * RS;
  float* sK = sQ + L * RS;
  float* sV = sK + L * RS;
  float* sE = sV + L * RS;
  float* sET = sE + L * RS;                              // mode 1 only
  float* sP = (a.mode == 1) ? sET + L * RS : sET;
  float* sdS = sP + L * SS;
  float* sDG = sdS + L * SS;
  const int len = min(max(a.seq_len[b], 0), L);
  const int64_t t0 = (int64_t)b * L;
  sa_load_rows(sQ, a.QKV + t0 * 3 * D, 3 * D, L, D);
  sa_load_rows(sK, a.QKV + t0 * 3 * D + D, 3 * D, L, D);
  sa_load_rows(sV, a.QKV + t0 * 3 * D + 2 * D, 3 * D, L, D);
  sa_load_rows(sE, a.enc + t0 * D, D, L, D);
  if (a.mode == 1) sa_load_rows(sET, a.ET + t0 * D, D, L, D);
  for (int p = tid; p < L * SS; p += SA_THREADS) sDG[p] = 0.f;
  // layer-norm backward per position -> dy (residual gradient and dO)
  for (int i = warp; i < L; i += nwarp) {
    float s1 = 0.f, s2 = 0.f;
    for (int d = lane; d < D; d += 32) {
      float dxh = a.dOut[(t0 + i) * D + d] * a.ln_gamma[d];
      s1 += dxh;
      s2 = fmaf(dxh, a.XH[(t0 + i) * D + d], s2);
    }
    float m1 = warp_sum(s1) / D, m2 = warp_sum(s2) / D, rstd = a.RSTD[t0 + i];
    for (int d = lane; d < D; d += 32) {
      float dxh = a.dOut[(t0 + i) * D + d] * a.ln_gamma[d];
      float dy = (dxh - m1 - a.XH[(t0 + i) * D + d] * m2) * rstd;
      sdO[i * RS + d] = dy;
      a.dEnc[(t0 + i) * D + d] = dy;
    }
  }
  __syncthreads();
  const float sqrt_dh = sqrtf((float)dh);
  const uint32_t drop_ctr = a.drop_thr24 ? *a.drop_counter : 0u;
  float* sPd = a.drop_thr24 ? sDG : sP;     // dropped weights (the V gradient uses them); sDG is free outside mode 1
  for (int h = 0; h < H; ++h) {
    const int c0 = h * dh;
    const float* Pg = a.P + ((int64_t)b * H + h) * L * L;
    for (int p = tid; p < len * len; p += SA_THREADS) {
      int i = p / len, j = p % len;
      const float pw = Pg[i * L + j];
      sP[i * SS + j] = pw;
      float s = 0.f;
      for (int k = 0; k < dh; ++k) s = fmaf(sdO[i * RS + c0 + k], sV[j * RS + c0 + k], s);
      if (a.drop_thr24) {                   // out = (P * m) V: dP = dPd * m, dV uses P * m
        const float m = sa_keep(a.drop_seed, drop_ctr, a.blk, (uint32_t)((((int64_t)b * H + h) * L + i) * L + j), a.drop_thr24)
                            ? a.drop_scale : 0.f;
        s *= m;
        sPd[i * SS + j] = pw * m;
      }
      sdS[i * SS + j] = s;  // dP
    }
    __syncthreads();
    for (int i = warp; i < len; i += nwarp) {
      float rs = 0.f;
      for (int j = lane; j < len; j += 32) rs = fmaf(sP[i * SS + j], sdS[i * SS + j], rs);
      rs = warp_sum(rs);
      for (int j = lane; j < len; j += 32) {
        float dS = sP[i * SS + j] * (sdS[i * SS + j] - rs);
        float dA = dS / sqrt_dh;
        if (a.mode == 1) {
          float A = a.A[(((int64_t)b * H + h) * L + i) * L + j];
          sDG[i * SS + j] += dS * A / sqrt_dh;          // gate is shared by the heads
          dA = dS * a.GT[((int64_t)b * L + i) * L + j] / sqrt_dh;
        }
        sdS[i * SS + j] = dA;
      }
    }
    __syncthreads();
    for (int p = tid; p < len * dh; p += SA_THREADS) {
      int r = p / dh, d = p % dh;
      float dv = 0.f, dq = 0.f, dk = 0.f;
      for (int x = 0; x < len; ++x) {
        dv = fmaf(sPd[x * SS + r], sdO[x * RS + c0 + d], dv);    // dV[j=r] = sum_i P[i,j] dO[i]
        dq = fmaf(sdS[r * SS + x], sK[x * RS + c0 + d], dq);     // dQ[i=r] = sum_j dA[i,j] K[j]
        dk = fmaf(sdS[x * SS + r], sQ[x * RS + c0 + d], dk);     // dK[j=r] = sum_i dA[i,j] Q[i]
      }
      float* g = a.dQKV + (t0 + r) * 3 * D + c0 + d;
      g[0] = (sQ[r * RS + c0 + d] > 0.f) ? dq : 0.f;
      g[D] = (sK[r * RS + c0 + d] > 0.f) ? dk : 0.f;
      g[2 * D] = (sV[r * RS + c0 + d] > 0.f) ? dv : 0.f;
    }
    __syncthreads();
  }
  if (a.mode == 1) {
    const float* w = a.gate;
    float* gb = a.GB + (int64_t)b * 5 * L * L;
    for (int p = tid; p < len * len; p += SA_THREADS) {
      int i = p / len, j = p % len, q = i * L + j;
      int64_t o = ((int64_t)b * L + i) * L + j;
      float g = a.GT[o], Z = a.Z[o], Dk = a.DK[o];
      float dG = sDG[i * SS + j] * g * (1.f - g);
      float dpre = dG * w[2 * L * L + q] * (1.f - Dk * Dk);
      float dlt = logf(fabsf(a.time_list[t0 + i] - a.time_list[t0 + j]) + 1.f);
      gb[q] = dpre * dlt;
      gb[L * L + q] = dpre;
      gb[2 * L * L + q] = dG * Dk;
      gb[3 * L * L + q] = dG * Z;
      gb[4 * L * L + q] = dG;
      sDG[i * SS + j] = dG * w[3 * L * L + q] * (1.f - Z * Z);   // dM
    }
    __syncthreads();
    for (int p = tid; p < len * D; p += SA_THREADS) {
      int r = p / D, d = p % D;
      float det = 0.f, de = 0.f;
      for (int x = 0; x < len; ++x) {
        det = fmaf(sDG[r * SS + x], sE[x * RS + d], det);        // dET[i=r] = sum_j dM[i,j] e[j]
        de = fmaf(sDG[x * SS + r], sET[x * RS + d], de);         // de[j=r] += sum_i dM[i,j] ET[i]
      }
      a.dET[(t0 + r) * D + d] = det;
      a.dEnc[(t0 + r) * D + d] += de;
    }
  }
}

// pred[b] = layer_norm(enc[b, seq_len[b]-1]) (eps 1e-12): one warp per sequence
__global__ void final_ln_fwd_kernel(const float* __restrict__ enc, const int32_t* __restrict__ seq_len, int B, int L,
                                    int D, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    float* __restrict__ pred, float* __restrict__ XHF, float* __restrict__ RSTDF) {
  int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  int pos = seq_len ? min(max(seq_len[b] - 1, 0), L - 1) : 0;
  const float* x = enc + ((int64_t)b * L + pos) * D;
  float s1 = 0.f;
  for (int d = lane; d < D; d += 32) s1 += x[d];
  float mean = warp_sum(s1) / D, s2 = 0.f;
  for (int d = lane; d < D; d += 32) { float y = x[d] - mean; s2 = fmaf(y, y, s2); }
  float rstd = rsqrtf(warp_sum(s2) / D + 1e-12f);
  for (int d = lane; d < D; d += 32) {
    float xh = (x[d] - mean) * rstd;
    XHF[(int64_t)b * D + d] = xh;
    pred[(int64_t)b * D + d] = fmaf(gamma[d], xh, beta[d]);
  }
  if (lane == 0) RSTDF[b] = rstd;
}
// dEnc[b, seq_len[b]-1] = LN backward of dpred[b]; dEnc pre-zeroed
__global__ void final_ln_bwd_kernel(const float* __restrict__ dpred, const float* __restrict__ XHF,
                                    const float* __restrict__ RSTDF, const int32_t* __restrict__ seq_len, int B, int L,
                                    int D, const float* __restrict__ gamma, float* __restrict__ dEnc) {
  int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  int pos = seq_len ? min(max(seq_len[b] - 1, 0), L - 1) : 0;
  float s1 = 0.f, s2 = 0.f;
  for (int d = lane; d < D; d += 32) {
    float dxh = dpred[(int64_t)b * D + d] * gamma[d];
    s1 += dxh;
    s2 = fmaf(dxh, XHF[(int64_t)b * D + d], s2);
  }
  float m1 = warp_sum(s1) / D, m2 = warp_sum(s2) / D, rstd = RSTDF[b];
  for (int d = lane; d < D; d += 32) {
    float dxh = dpred[(int64_t)b * D + d] * gamma[d];
    dEnc[((int64_t)b * L + pos) * D + d] = (dxh - m1 - XHF[(int64_t)b * D + d] * m2) * rstd;
  }
}

// plain rows: the gather kernels above with L = 1 and no length array (row b itself)
int ln_rows_forward(const float* x, int B, int D, const float* gamma, const float* beta, float* out, float* xhat, float* rstd,
                    cudaStream_t st) {
  final_ln_fwd_kernel<<<cdiv(B, 4), 128, 0, st>>>(x, nullptr, B, 1, D, gamma, beta, out, xhat, rstd);
  MTAM_LAUNCH_CHECK();
  return 0;
}
int ln_rows_backward(const float* dout, const float* xhat, const float* rstd, int B, int D, const float* gamma, float* dx,
                     cudaStream_t st) {
  final_ln_bwd_kernel<<<cdiv(B, 4), 128, 0, st>>>(dout, xhat, rstd, nullptr, B, 1, D, gamma, dx);
  MTAM_LAUNCH_CHECK();
  return 0;
}

static int sa_mode(int kind) { return time_aware(kind) ? 1 : (kind == MTAM_KIND_TISASREC ? 2 : 0); }

static int sa_check_smem(const mtam_config& c, size_t* fwd, size_t* bwd) {
  int mode = sa_mode(c.kind);
  *fwd = sa_smem_floats(c.L, c.D, mode, false) * sizeof(float);
  *bwd = sa_smem_floats(c.L, c.D, mode, true) * sizeof(float);
  if (*bwd > 227 * 1024)
    return set_error(MTAM_ERR_UNSUPPORTED, "self-attention block: L=%d, num_units=%d needs %zu bytes of shared memory (> 227 KB)",
                     c.L, c.D, *bwd);
  return 0;
}

// dropout applies to the plain (0) and TiSAS (2) blocks only: the time-aware block has none (time_aware_attention.py:440)
static void sa_dropout_args(const SaCtx& c, int mode, int blk, SaBlockArgs& a) {
  const bool on = mode != 1 && c.drop_rate > 0.f;
  a.drop_thr24 = on ? (uint32_t)lrintf(c.drop_rate * 16777216.f) : 0u;
  a.drop_scale = on ? 1.f / (1.f - c.drop_rate) : 1.f;
  a.drop_seed = c.drop_seed;
  a.blk = (uint32_t)blk;
  a.drop_counter = c.drop_counter;
}

int sa_forward(const SaCtx& c, cudaStream_t st) {
  const mtam_config& g = c.cfg;
  const int B = c.bt->B, L = g.L, D = g.D, N = g.N, H = g.H, mode = sa_mode(g.kind);
  const int64_t T = (int64_t)B * L;
  size_t sf, sb;
  MTAM_TRY(sa_check_smem(g, &sf, &sb));
  SaWs w = sa_plan(g, c.ws, c.ws_bytes);
  if (w.total > c.ws_bytes) return set_error(MTAM_ERR_WORKSPACE, "self-attention workspace too small");
  const int64_t Tm = (int64_t)g.max_batch * L;   // per-block strides use max_batch so offsets are static
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(sa_block_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sf));
  for (int i = 0; i < N; ++i) {
    const float* enc = (i == 0) ? c.X : w.ENCo + (size_t)(i - 1) * Tm * D;
    float* qkv = w.QKV + (size_t)i * Tm * 3 * D;
    GemmEpilogue e;
    e.bias = c.params + c.sl.b3 + (size_t)i * 3 * D;
    e.relu = 1;
    MTAM_TRY(gemm_any(g.gemm_mode, 0, 0, (int)T, 3 * D, D, enc, D, c.params + c.sl.W3 + (size_t)i * D * 3 * D, 3 * D, qkv, 3 * D, e,
                      c.gemm_ws, c.gemm_ws_bytes, st));
    SaBlockArgs a{};
    a.B = B; a.L = L; a.D = D; a.H = H; a.mode = mode;
    sa_dropout_args(c, mode, i, a);
    a.seq_len = c.bt->seq_length; a.time_list = c.bt->time_list; a.enc = enc; a.QKV = qkv;
    if (mode == 1) {
      float* et = w.ET + (size_t)i * Tm * D;
      GemmEpilogue e2;
      MTAM_TRY(gemm_any(g.gemm_mode, 0, 0, (int)T, D, D, enc, D, c.params + c.sl.Wt + (size_t)i * D * D, D, et, D, e2, c.gemm_ws,
                        c.gemm_ws_bytes, st));
      a.ET = et;
      a.gate = c.params + c.sl.gate + (size_t)i * 5 * L * L;
      a.A = w.A + (size_t)i * g.max_batch * H * L * L;
      a.Z = w.Z + (size_t)i * g.max_batch * L * L;
      a.DK = w.DK + (size_t)i * g.max_batch * L * L;
      a.GT = w.GT + (size_t)i * g.max_batch * L * L;
    }
    a.ln_gamma = c.params + c.sl.lng + (size_t)i * D;
    a.ln_beta = c.params + c.sl.lnb + (size_t)i * D;
    a.P = w.P + (size_t)i * g.max_batch * H * L * L;
    a.XH = w.XH + (size_t)i * Tm * D;
    a.RSTD = w.RSTD + (size_t)i * Tm;
    a.out = w.ENCo + (size_t)i * Tm * D;
    sa_block_fwd_kernel<<<B, SA_THREADS, sf, st>>>(a);
    MTAM_LAUNCH_CHECK();
  }
  const float* last = (N == 0) ? c.X : w.ENCo + (size_t)(N - 1) * Tm * D;
  final_ln_fwd_kernel<<<cdiv(B, 4), 128, 0, st>>>(last, c.bt->seq_length, B, L, D, c.params + c.sl.lnfg,
                                                  c.params + c.sl.lnfb, c.pred, c.XHF, c.RSTDF);
  MTAM_LAUNCH_CHECK();
  return 0;
}

int sa_backward(const SaCtx& c, cudaStream_t st) {
  const mtam_config& g = c.cfg;
  const int B = c.bt->B, L = g.L, D = g.D, N = g.N, H = g.H, mode = sa_mode(g.kind);
  const int64_t T = (int64_t)B * L;
  size_t sf, sb;
  MTAM_TRY(sa_check_smem(g, &sf, &sb));
  SaWs w = sa_plan(g, c.ws, c.ws_bytes);
  const int64_t Tm = (int64_t)g.max_batch * L;
  float* G = c.grads;
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(sa_block_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb));
  // final layer norm
  MTAM_TRY(colsum_f32(c.dpred, D, nullptr, 0, B, D, G + c.sl.lnfb, 0, c.colsum_ws, c.colsum_ws_bytes, st));
  MTAM_TRY(colsum_f32(c.dpred, D, c.XHF, D, B, D, G + c.sl.lnfg, 0, c.colsum_ws, c.colsum_ws_bytes, st));
  float* dNext = (N % 2 == 0) ? c.dX : w.dEncA;   // ping-pong so that block 0 writes into dX
  // choose buffers: gradient w.r.t. the output of block i lives in buf(i+1); buf(0) == dX
  auto buf = [&](int k) -> float* { return (k == 0) ? c.dX : ((k % 2) ? w.dEncA : w.dEncB); };
  (void)dNext;
  float* dOutN = buf(N);
  MTAM_CUDA_CHECK(cudaMemsetAsync(dOutN, 0, (size_t)T * D * sizeof(float), st));
  final_ln_bwd_kernel<<<cdiv(B, 4), 128, 0, st>>>(c.dpred, c.XHF, c.RSTDF, c.bt->seq_length, B, L, D,
                                                  c.params + c.sl.lnfg, dOutN);
  MTAM_LAUNCH_CHECK();
  GemmEpilogue e0, eacc;
  eacc.accumulate = 1;
  for (int i = N - 1; i >= 0; --i) {
    const float* enc = (i == 0) ? c.X : w.ENCo + (size_t)(i - 1) * Tm * D;
    const float* dOut = buf(i + 1);
    float* dEnc = buf(i);
    const float* xh = w.XH + (size_t)i * Tm * D;
    // layer-norm parameter gradients: rows past the sequence length carry zero dOut
    MTAM_TRY(colsum_f32(dOut, D, nullptr, 0, (int)T, D, G + c.sl.lnb + (size_t)i * D, 0, c.colsum_ws, c.colsum_ws_bytes, st));
    MTAM_TRY(colsum_f32(dOut, D, xh, D, (int)T, D, G + c.sl.lng + (size_t)i * D, 0, c.colsum_ws, c.colsum_ws_bytes, st));
    MTAM_CUDA_CHECK(cudaMemsetAsync(w.dQKV, 0, (size_t)T * 3 * D * sizeof(float), st));
    SaBlockArgs a{};
    a.B = B; a.L = L; a.D = D; a.H = H; a.mode = mode;
    sa_dropout_args(c, mode, i, a);
    a.seq_len = c.bt->seq_length; a.time_list = c.bt->time_list; a.enc = enc;
    a.QKV = w.QKV + (size_t)i * Tm * 3 * D;
    a.ln_gamma = c.params + c.sl.lng + (size_t)i * D;
    a.P = w.P + (size_t)i * g.max_batch * H * L * L;
    a.XH = const_cast<float*>(xh);
    a.RSTD = w.RSTD + (size_t)i * Tm;
    a.dOut = dOut; a.dEnc = dEnc; a.dQKV = w.dQKV;
    if (mode == 1) {
      MTAM_CUDA_CHECK(cudaMemsetAsync(w.dET, 0, (size_t)T * D * sizeof(float), st));
      MTAM_CUDA_CHECK(cudaMemsetAsync(w.GB, 0, (size_t)B * 5 * L * L * sizeof(float), st));
      a.ET = w.ET + (size_t)i * Tm * D;
      a.gate = c.params + c.sl.gate + (size_t)i * 5 * L * L;
      a.A = w.A + (size_t)i * g.max_batch * H * L * L;
      a.Z = w.Z + (size_t)i * g.max_batch * L * L;
      a.DK = w.DK + (size_t)i * g.max_batch * L * L;
      a.GT = w.GT + (size_t)i * g.max_batch * L * L;
      a.dET = w.dET; a.GB = w.GB;
    }
    sa_block_bwd_kernel<<<B, SA_THREADS, sb, st>>>(a);
    MTAM_LAUNCH_CHECK();
    const float* W3 = c.params + c.sl.W3 + (size_t)i * D * 3 * D;
    MTAM_TRY(colsum_f32(w.dQKV, 3 * D, nullptr, 0, (int)T, 3 * D, G + c.sl.b3 + (size_t)i * 3 * D, 0, c.colsum_ws,
                        c.colsum_ws_bytes, st));
    MTAM_TRY(gemm_any(g.gemm_mode, 1, 0, D, 3 * D, (int)T, enc, D, w.dQKV, 3 * D, G + c.sl.W3 + (size_t)i * D * 3 * D, 3 * D, e0,
                      c.gemm_ws, c.gemm_ws_bytes, st));
    MTAM_TRY(gemm_any(g.gemm_mode, 0, 1, (int)T, D, 3 * D, w.dQKV, 3 * D, W3, 3 * D, dEnc, D, eacc, c.gemm_ws, c.gemm_ws_bytes, st));
    if (mode == 1) {
      const float* Wt = c.params + c.sl.Wt + (size_t)i * D * D;
      MTAM_TRY(gemm_any(g.gemm_mode, 1, 0, D, D, (int)T, enc, D, w.dET, D, G + c.sl.Wt + (size_t)i * D * D, D, e0, c.gemm_ws,
                        c.gemm_ws_bytes, st));
      MTAM_TRY(gemm_any(g.gemm_mode, 0, 1, (int)T, D, D, w.dET, D, Wt, D, dEnc, D, eacc, c.gemm_ws, c.gemm_ws_bytes, st));
      MTAM_TRY(colsum_f32(w.GB, 5 * L * L, nullptr, 0, B, 5 * L * L, G + c.sl.gate + (size_t)i * 5 * L * L, 0,
                          c.colsum_ws, c.colsum_ws_bytes, st));
    }
  }
  return 0;
}

}  // namespace mtam
