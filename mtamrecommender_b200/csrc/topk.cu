// Full-catalogue scoring with fused top-k selection, list merge and HR/NDCG.
//
// Reference: base_model.metrics_topK  Model/base_model.py:188-213
//   item_result = tf.matmul(pred, item_table^T)  (:194-195); tf.nn.top_k(item_result, k) for
//   k in {1,5,10,30,50} (:196-200) -- one top-50 pass, the others are prefixes (top_k is sorted
//   descending, equal values -> lower index first);  calculate_topK  (:215-242).
// The [B,V] score matrix is never written: a CTA owns 64 prediction rows and a contiguous chunk
// of items, keeps the rows' running top-k in shared memory and only appends tile candidates that
// beat the current k-th score.
//
// MTAM_GEMM_TF32X3 (num_units 32 / 64): the [B,V] product runs on tcgen05 as a FILTER.  ce_bucket_max_tc
// (ce_tc.cu, the softmax forward pipeline with a max epilogue) leaves only the maximum logit of every bucket of
// 16 / 64 consecutive items; topk_refine_kernel then picks, per pred row, the k+14 buckets with the largest maxima
// (radix select, ties -> lower bucket), rescores their items with the same fp32 FMA chain as the kernel above and
// selects the top k under (score desc, index asc).  The returned scores are fp32 dot products (not tf32 values), and
// the selection is exact with respect to them unless more than 14 bucket maxima lie within the 3xTF32 rounding
// error (~1e-6 relative) of the k-th one.  An item's score is computed in one fixed order, so sharded and unsharded
// runs agree bit for bit; it can differ from the MTAM_GEMM_FP32 kernel's in the last bit (another summation order).
#include <limits.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "model_kernels.h"
#include "../../include/mtam.h"

namespace mtam {

constexpr int KT = 64;          // tile edge
constexpr int KPAD = KT + 4;
constexpr int KMAX = 64;        // largest supported k

__device__ __forceinline__ bool better(float s1, int i1, float s2, int i2) {
  return s1 > s2 || (s1 == s2 && i1 < i2);
}

template <int D>
__global__ void __launch_bounds__(256) score_topk_kernel(const float* __restrict__ pred, const float* __restrict__ table,
                                                         int B, int row_begin, int row_end, int k, int chunk_items,
                                                         float* __restrict__ cand_score, int32_t* __restrict__ cand_idx) {
  extern __shared__ __align__(16) float sm[];
  float* Ts = sm;                       // [D][KPAD]
  float* Ps = Ts + D * KPAD;            // [D][KPAD]
  float* ls = Ps + D * KPAD;            // [KT][KMAX] list scores (sorted, best first)
  int* li = (int*)(ls + KT * KMAX);     // [KT][KMAX] list indices
  float* qs = (float*)(li + KT * KMAX); // [KT][KT] candidate queue scores
  int* qi = (int*)(qs + KT * KT);       // [KT][KT]
  int* qn = qi + KT * KT;               // [KT] queue lengths
  float* thr = (float*)(qn + KT);       // [KT] current k-th score
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * KT;
  const int c0 = row_begin + blockIdx.x * chunk_items;
  const int c1 = min(row_end, c0 + chunk_items);
  for (int i = threadIdx.x; i < KT * KMAX; i += 256) { ls[i] = -INFINITY; li[i] = INT_MAX; }
  if (threadIdx.x < KT) { thr[threadIdx.x] = -INFINITY; qn[threadIdx.x] = 0; }
  // prediction tile, k-major
  for (int i = threadIdx.x; i < KT * (D / 4); i += 256) {
    int r = i % KT, c4 = (i / KT) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + r < B) v = __ldg(reinterpret_cast<const float4*>(pred + (int64_t)(m0 + r) * D + c4));
    Ps[(c4 + 0) * KPAD + r] = v.x; Ps[(c4 + 1) * KPAD + r] = v.y;
    Ps[(c4 + 2) * KPAD + r] = v.z; Ps[(c4 + 3) * KPAD + r] = v.w;
  }
  for (int v0 = c0; v0 < c1; v0 += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * (D / 4); i += 256) {
      int r = i % KT, c4 = (i / KT) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v0 + r < c1) v = __ldg(reinterpret_cast<const float4*>(table + (int64_t)(v0 + r) * D + c4));
      Ts[(c4 + 0) * KPAD + r] = v.x; Ts[(c4 + 1) * KPAD + r] = v.y;
      Ts[(c4 + 2) * KPAD + r] = v.z; Ts[(c4 + 3) * KPAD + r] = v.w;
    }
    __syncthreads();
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
    for (int kk = 0; kk < D; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(Ps + kk * KPAD + ty * 4);
      float4 b = *reinterpret_cast<const float4*>(Ts + kk * KPAD + tx * 4);
      float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const float t = thr[r];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = v0 + tx * 4 + j;
        // items arrive in increasing index inside a CTA, so an equal score can never displace
        if (col < c1 && acc[i][j] > t) {
          int slot = atomicAdd(&qn[r], 1);
          qs[r * KT + slot] = acc[i][j];
          qi[r * KT + slot] = col;
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < KT) {
      const int r = threadIdx.x;
      const int n = qn[r];
      float* S = ls + r * KMAX;
      int* I = li + r * KMAX;
      for (int c = 0; c < n; ++c) {
        float s = qs[r * KT + c];
        int id = qi[r * KT + c];
        if (!better(s, id, S[k - 1], I[k - 1])) continue;
        int p = k - 1;
        while (p > 0 && better(s, id, S[p - 1], I[p - 1])) { S[p] = S[p - 1]; I[p] = I[p - 1]; --p; }
        S[p] = s; I[p] = id;
      }
      qn[r] = 0;
      thr[r] = S[k - 1];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < KT * k; i += 256) {
    int r = i / k, c = i % k;
    if (m0 + r < B) {
      int64_t o = ((int64_t)blockIdx.x * B + (m0 + r)) * k + c;
      cand_score[o] = ls[r * KMAX + c];
      cand_idx[o] = li[r * KMAX + c];
    }
  }
}

// one CTA per row: pick the best k of n_lists*k candidates under (score desc, index asc)
__global__ void __launch_bounds__(128) merge_topk_kernel(const int32_t* __restrict__ in_idx, const float* __restrict__ in_score,
                                                         int n_lists, int B, int k, int32_t* __restrict__ out_idx,
                                                         float* __restrict__ out_score) {
  extern __shared__ __align__(16) float sm[];
  const int n = n_lists * k;
  float* S = sm;
  int* I = (int*)(sm + n);
  __shared__ float ws[4];
  __shared__ int wi[4], wp[4];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += 128) {
    int l = i / k, c = i % k;
    int64_t o = ((int64_t)l * B + b) * k + c;
    S[i] = in_score[o];
    I[i] = in_idx[o];
  }
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    float bs = -INFINITY;
    int bi = INT_MAX, bp = -1;
    for (int i = threadIdx.x; i < n; i += 128)
      if (better(S[i], I[i], bs, bi) || bp < 0) {
        if (bp < 0 || better(S[i], I[i], bs, bi)) { bs = S[i]; bi = I[i]; bp = i; }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float s2 = __shfl_xor_sync(0xffffffffu, bs, o);
      int i2 = __shfl_xor_sync(0xffffffffu, bi, o), p2 = __shfl_xor_sync(0xffffffffu, bp, o);
      if (p2 >= 0 && (bp < 0 || better(s2, i2, bs, bi) || (s2 == bs && i2 == bi && p2 < bp))) { bs = s2; bi = i2; bp = p2; }
    }
    if ((threadIdx.x & 31) == 0) { ws[threadIdx.x >> 5] = bs; wi[threadIdx.x >> 5] = bi; wp[threadIdx.x >> 5] = bp; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 4; ++w)
        if (wp[w] >= 0 && (bp < 0 || better(ws[w], wi[w], bs, bi) || (ws[w] == bs && wi[w] == bi && wp[w] < bp))) {
          bs = ws[w]; bi = wi[w]; bp = wp[w];
        }
      out_idx[(int64_t)b * k + r] = bi;
      if (out_score) out_score[(int64_t)b * k + r] = bs;
      if (bp >= 0) { S[bp] = -INFINITY; I[bp] = INT_MAX; }
    }
    __syncthreads();
  }
}

// HR@k / NDCG@k for k in {1,5,10,30,50} (calculate_topK): single block, fixed order
__global__ void __launch_bounds__(256) hr_ndcg_kernel(const int32_t* __restrict__ topk, int B, int k,
                                                      const int32_t* __restrict__ target, float* __restrict__ out10) {
  __shared__ float red[32];
  const int ks[5] = {1, 5, 10, 30, 50};
  float hit[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, nd[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int b = threadIdx.x; b < B; b += 256) {
    int t = target[b], rank = -1;
    for (int c = 0; c < k; ++c)
      if (topk[(int64_t)b * k + c] == t) { rank = c; break; }
    if (rank >= 0) {
      float g = logf(2.f) / logf((float)rank + 2.f);
#pragma unroll
      for (int q = 0; q < 5; ++q)
        if (rank < ks[q] && ks[q] <= k) { hit[q] += 1.f; nd[q] += g; }
    }
  }
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    float h = block_sum(hit[q], red);
    float n = block_sum(nd[q], red);
    if (threadIdx.x == 0) { out10[2 * q] = h / B; out10[2 * q + 1] = n / B; }
  }
}


// ---- tensor-core filter + exact rescoring -----------------------------------------------------------------
constexpr int kBucketPad = 14;                       // buckets kept beyond k
constexpr int kMaxSel = KMAX + kBucketPad;           // 78
// bucket size: per pred row the maxima cost rows/bs*8 B of traffic and the rescoring (k+14)*bs*4D B, equal near 2 M rows
constexpr int kSmallBucketMaxRowsDefault = 1 << 21;
static int g_small_bucket_max_rows = kSmallBucketMaxRowsDefault;   // mtam_set_topk_bucket_crossover
constexpr int kMaxGroups = 16384 - 1;                    // groups of 32 buckets per pred row that fit in shared memory

__device__ __forceinline__ uint32_t order_key(float x) {   // monotone float -> uint32 (every real key > 0)
  const uint32_t b = __float_as_uint(x);
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float key_value(uint32_t u) {
  return __uint_as_float(u ^ ((u >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}

// exclusive scan of one int per thread over a 256-thread block, in thread order; total in *total
__device__ __forceinline__ int block_excl_scan256(int v, int* warp_tot, int* total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int t = warp_tot[i];
    if (i < w) base += t;
    tot += t;
  }
  *total = tot;
  return base + inc - v;
}

struct SelectScratch { int hist[256]; uint32_t mn[8], mx[8]; int warp_tot[8]; uint32_t prefix; int need; };

// keys[0..n) in shared memory, 256 threads.  Writes to out[0..take) the positions of the `take` largest keys
// (equal keys: lower position first) in ascending position order.  take <= n.
// MSB-first radix select, 8 bits per pass, that starts at the highest bit on which two keys differ: the scores
// of one row share their leading bits, which would otherwise pile every key onto one histogram bin.
__device__ void select_largest(const uint32_t* __restrict__ keys, int n, int take, int* __restrict__ out, SelectScratch* sc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t mn = 0xFFFFFFFFu, mx = 0;
  for (int i = tid; i < n; i += 256) { const uint32_t u = keys[i]; mn = min(mn, u); mx = max(mx, u); }
  mn = __reduce_min_sync(0xffffffffu, mn);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 0) { sc->mn[warp] = mn; sc->mx[warp] = mx; }
  sc->hist[tid] = 0;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w) { mn = min(mn, sc->mn[w]); mx = max(mx, sc->mx[w]); }
  uint32_t T = mx;
  int need = take;
  if (mn != mx) {
    int top = 32 - __clz(mn ^ mx);                         // bits [0, top) are the ones that vary
    T = top == 32 ? 0u : (mx >> top) << top;               // the shared leading bits
    bool first = true;
    while (top > 0) {
      const int nb = min(8, top), shift = top - nb;
      const uint32_t dmask = (1u << nb) - 1u;
      for (int i = tid; i < n; i += 256) {
        const uint32_t u = keys[i];
        if (first || (u >> top) == (T >> top)) atomicAdd(&sc->hist[(u >> shift) & dmask], 1);
      }
      __syncthreads();
      if (warp == 0) {
        // bins 255 .. 0 in descending order, 8 per lane: lane l owns bins 255-8l .. 248-8l
        int c[8], tot = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { c[q] = sc->hist[255 - (lane * 8 + q)]; tot += c[q]; }
        int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        int before = inc - tot;                   // keys in strictly higher bins than this lane's
        if (before < need && inc >= need) {       // the bin holding the need-th largest is one of this lane's
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (before < need && before + c[q] >= need) {
              sc->prefix = T | ((uint32_t)(255 - (lane * 8 + q)) << shift);
              sc->need = need - before;
              before = need;                      // stop
            } else if (before < need) {
              before += c[q];
            }
          }
        }
      }
      __syncthreads();
      T = sc->prefix;
      need = sc->need;
      sc->hist[tid] = 0;
      __syncthreads();
      top = shift;
      first = false;
    }
  }
  // T = the take-th largest key; every key above it is taken, and the first `need` equal to it
  const int seg = (n + 255) / 256;
  const int i0 = min(n, tid * seg), i1 = min(n, i0 + seg);
  int ngt = 0, neq = 0;
  for (int i = i0; i < i1; ++i) { const uint32_t u = keys[i]; ngt += u > T; neq += u == T; }
  int tot_gt, tot_eq;
  const int ogt = block_excl_scan256(ngt, sc->warp_tot, &tot_gt);
  int oeq = block_excl_scan256(neq, sc->warp_tot, &tot_eq);
  // ascending position order: an element's slot = number of taken elements before it
  int slot = ogt + min(oeq, need);
  for (int i = i0; i < i1; ++i) {
    const uint32_t u = keys[i];
    if (u > T) out[slot++] = i;
    else if (u == T) {
      if (oeq < need) out[slot++] = i;
      ++oeq;
    }
  }
  __syncthreads();
}

// one CTA per pred row: pick n_sel buckets by their maxima (two levels: groups of 32 buckets first), rescore their
// items in fp32, emit the sorted top k
template <int D>
__global__ void __launch_bounds__(256, 8) topk_refine_kernel(const float* __restrict__ pred, const float* __restrict__ table,
                                                          const float* __restrict__ bmax, int ld, int n_buckets, int log2bs,
                                                          int n_sel, int row_begin, int row_end, int k,
                                                          int32_t* __restrict__ idx_out, float* __restrict__ score_out) {
  extern __shared__ __align__(16) uint32_t dsm[];
  __shared__ SelectScratch sc;
  static_assert(D == 32 || D == 64, "8 lanes x 1 or 2 float4 per item");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int bs = 1 << log2bs;
  const int G = (n_buckets + 31) >> 5;             // groups of 32 buckets
  const int n_grp = min(G, n_sel);                 // groups kept
  const int n_cand = n_sel << log2bs;
  float* qs = reinterpret_cast<float*>(dsm);       // [D] pred row
  uint32_t* sup = dsm + D;                         // [G] group maxima (keys)
  uint32_t* bk = sup + G;                          // [n_grp*32] bucket maxima of the kept groups (keys)
  uint32_t* ck = bk + n_grp * 32;                  // [n_cand] candidate scores (keys)
  int* gsel = reinterpret_cast<int*>(ck + n_cand); // [n_grp] kept groups, ascending
  int* sel = gsel + n_grp;                         // [n_sel] kept buckets, ascending
  int* win = sel + n_sel;                          // [k] winning candidates, ascending item index
  const float* mrow = bmax + (int64_t)b * ld;

  if (tid < D / 4)
    reinterpret_cast<float4*>(qs)[tid] = __ldg(reinterpret_cast<const float4*>(pred + (int64_t)b * D) + tid);
  // ---- 1. group maxima: a lane reads 4 consecutive bucket maxima, 8 lanes make a group of 32 ----
  for (int g0 = warp * 4; g0 < G; g0 += 32) {
    const int g = g0 + (lane >> 3), i = g * 32 + (lane & 7) * 4;
    float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    if (i + 3 < n_buckets) {
      v = __ldg(reinterpret_cast<const float4*>(mrow + i));
    } else if (i < n_buckets) {                      // the row's last, partial quad
      v.x = __ldg(mrow + i);
      if (i + 1 < n_buckets) v.y = __ldg(mrow + i + 1);
      if (i + 2 < n_buckets) v.z = __ldg(mrow + i + 2);
    }
    float m = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
    if ((lane & 7) == 0 && g < G) sup[g] = order_key(m);
  }
  __syncthreads();
  // ---- 2. the n_grp best groups contain the n_sel best buckets (a group's key >= each of its buckets') ----
  select_largest(sup, G, n_grp, gsel, &sc);
  for (int j = warp; j < n_grp; j += 8) {
    const int i = gsel[j] * 32 + lane;
    bk[j * 32 + lane] = i < n_buckets ? order_key(__ldg(mrow + i)) : 0u;
  }
  __syncthreads();
  // ---- 3. the n_sel best buckets (ties -> lower bucket: positions in bk ascend with the bucket id) ----
  select_largest(bk, n_grp * 32, n_sel, sel, &sc);
  if (tid < n_sel) {
    const int p = sel[tid];
    sel[tid] = gsel[p >> 5] * 32 + (p & 31);
  }
  __syncthreads();
  // ---- 4. fp32 scores of those buckets' items: 8 lanes per item (each a coalesced float4 of either half of the row:
  //         two 4-term FMA chains, their sum, then a butterfly over the 8 lanes; D = 32: one chain) -- a fixed order,
  //         so an item's score does not depend on the batch or on the row range ----
  {
    constexpr int NQ = D / 32;                     // float4 per lane
    const int sub = lane & 7;
    const int* __restrict__ selr = sel;
    uint32_t* __restrict__ ckr = ck;
    float4 q[NQ];
#pragma unroll
    for (int h = 0; h < NQ; ++h) q[h] = reinterpret_cast<const float4*>(qs)[sub + 8 * h];
#pragma unroll 2
    for (int c0 = warp * 4; c0 < n_cand; c0 += 32) {
      const int c = c0 + (lane >> 3);
      const int item = row_begin + (selr[c >> log2bs] << log2bs) + (c & (bs - 1));
      const bool valid = item < row_end;
      const float4* x = reinterpret_cast<const float4*>(table + (int64_t)item * D) + sub;
      float acc = 0.f;
#pragma unroll
      for (int h = 0; h < NQ; ++h) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) v = __ldg(x + 8 * h);
        float a = q[h].x * v.x;
        a = fmaf(q[h].y, v.y, a);
        a = fmaf(q[h].z, v.z, a);
        a = fmaf(q[h].w, v.w, a);
        acc = h == 0 ? a : acc + a;
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      if (sub == 0) ckr[c] = valid ? order_key(acc) : 0u;    // past the row range: below every real score
    }
  }
  __syncthreads();
  // ---- 5. the k best candidates (ties -> lower item index), then their order by rank counting ----
  select_largest(ck, n_cand, k, win, &sc);
  if (tid < k) {
    const int p = win[tid];
    const uint32_t u = ck[p];
    int rank = 0;
    for (int j = 0; j < k; ++j) {
      const uint32_t o = ck[win[j]];
      rank += (o > u) || (o == u && j < tid);
    }
    idx_out[(int64_t)b * k + rank] = row_begin + (sel[p >> log2bs] << log2bs) + (p & (bs - 1));
    if (score_out) score_out[(int64_t)b * k + rank] = key_value(u);
  }
}

static size_t refine_smem_bytes(int D, int n_buckets, int bs, int n_sel, int k) {
  const int G = cdiv(n_buckets, 32), n_grp = std::min(G, n_sel);
  return (size_t)(D + G + n_grp * 32 + n_sel * bs + n_grp + n_sel + k) * sizeof(uint32_t);
}

// geometry of the tensor-core path for `rows` catalogue rows
struct TcTopkPlan { int bs, ld, n_buckets, n_sel, chunk_rows; };
static TcTopkPlan tc_topk_plan(int B, int rows, int k, size_t ws_bytes) {
  TcTopkPlan p;
  p.bs = rows > g_small_bucket_max_rows ? 64 : 16;
  p.ld = cdiv(cdiv(rows, 128) * (128 / p.bs), 8) * 8;
  p.n_buckets = cdiv(rows, p.bs);
  p.n_sel = std::min(p.n_buckets, k + kBucketPad);
  const int64_t fit = (int64_t)(ws_bytes / sizeof(float)) / p.ld;
  p.chunk_rows = (int)std::min<int64_t>(fit, B);
  if (p.chunk_rows < B) p.chunk_rows = p.chunk_rows / 128 * 128;   // whole 128-row tiles except for the last chunk
  return p;
}
static size_t tc_topk_workspace_bytes(int B, int rows) {
  const int bs = rows > g_small_bucket_max_rows ? 64 : 16;
  const size_t ld = (size_t)cdiv(cdiv(rows, 128) * (128 / bs), 8) * 8;
  // bucket maxima of up to B pred rows; beyond 1 GiB the pred rows are processed in chunks of >= 128
  const size_t cap_rows = std::max<size_t>(128, ((size_t)1 << 30) / (ld * sizeof(float)) / 128 * 128);
  return std::min<size_t>((size_t)B, cap_rows) * ld * sizeof(float) + 256;
}

template <int D>
static int score_topk_tc_launch(const float* pred, int B, const float* table, int row_begin, int row_end, int k,
                                int32_t* idx_out, float* score_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int rows = row_end - row_begin;
  const TcTopkPlan p = tc_topk_plan(B, rows, k, ws_bytes);
  if (p.chunk_rows < std::min(B, 128)) return set_error(MTAM_ERR_WORKSPACE, "score_topk (tensor-core path): workspace too small");
  if (cdiv(p.n_buckets, 32) > kMaxGroups)
    return set_error(MTAM_ERR_INVALID, "score_topk (tensor-core path): %d rows exceed the supported range", rows);
  const size_t smem = refine_smem_bytes(D, p.n_buckets, p.bs, p.n_sel, k);
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(topk_refine_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
  float* bmax = (float*)ws;
  for (int b0 = 0; b0 < B; b0 += p.chunk_rows) {
    const int nb = std::min(p.chunk_rows, B - b0);
    MTAM_TRY(ce_bucket_max_tc(D, pred + (int64_t)b0 * D, nb, table + (int64_t)row_begin * D, rows, p.bs, bmax, p.ld, st));
    topk_refine_kernel<D><<<nb, 256, smem, st>>>(pred + (int64_t)b0 * D, table, bmax, p.ld, p.n_buckets, p.bs == 64 ? 6 : 4, p.n_sel,
                                                 row_begin, row_end, k, idx_out + (int64_t)b0 * k,
                                                 score_out ? score_out + (int64_t)b0 * k : nullptr);
    MTAM_LAUNCH_CHECK();
  }
  return 0;
}

static int topk_chunks(int B, int rows) {
  int row_tiles = cdiv(B, KT);
  int want = std::max(1, cdiv(2 * kNumSMs, row_tiles));
  int maxc = std::max(1, cdiv(rows, KT));
  return std::min(want, maxc);
}

static size_t score_topk_fp32_workspace_bytes(int B, int rows, int k) {
  int nc = topk_chunks(B, rows);
  return align_up((size_t)nc * B * k * sizeof(float), 256) + align_up((size_t)nc * B * k * sizeof(int32_t), 256) + 512;
}
// enough for either path
size_t score_topk_workspace_bytes(int B, int rows, int k) {
  return std::max(score_topk_fp32_workspace_bytes(B, rows, k), tc_topk_workspace_bytes(B, rows));
}

int merge_topk(const int32_t* in_idx, const float* in_score, int n_lists, int B, int k, int32_t* out_idx,
               float* out_score, cudaStream_t st) {
  size_t smem = (size_t)n_lists * k * 8;
  if (smem > 200 * 1024) return set_error(MTAM_ERR_INVALID, "merge_topk: %d lists x k=%d exceed shared memory", n_lists, k);
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
  merge_topk_kernel<<<B, 128, smem, st>>>(in_idx, in_score, n_lists, B, k, out_idx, out_score);
  MTAM_LAUNCH_CHECK();
  return 0;
}

template <int D>
static int score_topk_launch(const float* pred, int B, const float* table, int row_begin, int row_end, int k,
                             int32_t* idx_out, float* score_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int rows = row_end - row_begin;
  const int nc = topk_chunks(B, rows);
  if (ws_bytes < score_topk_fp32_workspace_bytes(B, rows, k)) return set_error(MTAM_ERR_WORKSPACE, "score_topk: workspace too small");
  float* cs = (float*)ws;
  int32_t* ci = (int32_t*)((char*)ws + align_up((size_t)nc * B * k * sizeof(float), 256));
  int chunk_items = cdiv(cdiv(rows, nc), KT) * KT;
  size_t smem = (size_t)(2 * D * KPAD + 2 * KT * KMAX + 2 * KT * KT + 2 * KT) * sizeof(float);
  MTAM_CUDA_CHECK(cudaFuncSetAttribute(score_topk_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(nc, cdiv(B, KT));
  score_topk_kernel<D><<<grid, 256, smem, st>>>(pred, table, B, row_begin, row_end, k, chunk_items, cs, ci);
  MTAM_LAUNCH_CHECK();
  return merge_topk(ci, cs, nc, B, k, idx_out, score_out, st);
}

int score_topk(int mode, const float* pred, int B, int D, const float* table, int row_begin, int row_end, int k,
               int32_t* idx_out, float* score_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (k < 1 || k > KMAX) return set_error(MTAM_ERR_INVALID, "top-k: k=%d outside [1,%d]", k, KMAX);
  if (row_end - row_begin < k) return set_error(MTAM_ERR_INVALID, "top-k: fewer than k=%d rows to score", k);
  if (mode != MTAM_GEMM_FP32 && !gemm_mode_is_tc(mode)) return set_error(MTAM_ERR_INVALID, "top-k: unknown gemm_mode %d", mode);
  if (gemm_mode_is_tc(mode) && ce_tc_supported(D)) {
    if (D == 64) return score_topk_tc_launch<64>(pred, B, table, row_begin, row_end, k, idx_out, score_out, ws, ws_bytes, st);
    return score_topk_tc_launch<32>(pred, B, table, row_begin, row_end, k, idx_out, score_out, ws, ws_bytes, st);
  }
  switch (D) {
    case 32: return score_topk_launch<32>(pred, B, table, row_begin, row_end, k, idx_out, score_out, ws, ws_bytes, st);
    case 64: return score_topk_launch<64>(pred, B, table, row_begin, row_end, k, idx_out, score_out, ws, ws_bytes, st);
    case 128: return score_topk_launch<128>(pred, B, table, row_begin, row_end, k, idx_out, score_out, ws, ws_bytes, st);
  }
  return set_error(MTAM_ERR_INVALID, "top-k: num_units=%d not supported (32, 64, 128)", D);
}

}  // namespace mtam

extern "C" {

size_t mtam_score_topk_workspace(int32_t B, int32_t rows, int32_t k) { return mtam::score_topk_workspace_bytes(B, rows, k); }

int mtam_score_topk(int32_t gemm_mode, const float* pred, int32_t B, int32_t D, const float* item_table, int32_t row_begin,
                    int32_t row_end, int32_t k, int32_t* idx_out, float* score_out, void* workspace, size_t workspace_bytes,
                    void* stream) {
  if (!pred || !item_table || !idx_out || !workspace || B < 1)
    return mtam::set_error(MTAM_ERR_INVALID, "mtam_score_topk: null argument");
  return mtam::score_topk(gemm_mode, pred, B, D, item_table, row_begin, row_end, k, idx_out, score_out, workspace, workspace_bytes,
                          (cudaStream_t)stream);
}

int mtam_set_topk_bucket_crossover(int32_t rows) {
  if (rows < 0) return mtam::set_error(MTAM_ERR_INVALID, "mtam_set_topk_bucket_crossover: rows < 0");
  mtam::g_small_bucket_max_rows = rows == 0 ? mtam::kSmallBucketMaxRowsDefault : rows;
  return 0;
}

int mtam_score_bucket_max(const float* pred, int32_t B, int32_t D, const float* item_table, int32_t rows, int32_t bucket_size,
                          float* bmax, int32_t ld, void* stream) {
  if (!pred || !item_table || !bmax || B < 1 || rows < 1)
    return mtam::set_error(MTAM_ERR_INVALID, "mtam_score_bucket_max: bad argument");
  const int need = mtam::cdiv(mtam::cdiv(rows, 128) * (128 / std::max(bucket_size, 1)), 8) * 8;
  if (ld < need) return mtam::set_error(MTAM_ERR_INVALID, "mtam_score_bucket_max: ld=%d < %d", ld, need);
  return mtam::ce_bucket_max_tc(D, pred, B, item_table, rows, bucket_size, bmax, ld, (cudaStream_t)stream);
}

int mtam_merge_topk(const int32_t* in_idx, const float* in_score, int32_t n_lists, int32_t B, int32_t k, int32_t* out_idx,
                    float* out_score, void* stream) {
  if (!in_idx || !in_score || !out_idx || n_lists < 1 || B < 1 || k < 1)
    return mtam::set_error(MTAM_ERR_INVALID, "mtam_merge_topk: bad argument");
  return mtam::merge_topk(in_idx, in_score, n_lists, B, k, out_idx, out_score, (cudaStream_t)stream);
}

int mtam_hr_ndcg(const int32_t* topk_idx, int32_t B, int32_t k, const int32_t* target, float* out10, void* stream) {
  if (!topk_idx || !target || !out10 || B < 1 || k < 1) return mtam::set_error(MTAM_ERR_INVALID, "mtam_hr_ndcg: bad argument");
  mtam::hr_ndcg_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(topk_idx, B, k, target, out10);
  MTAM_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
