// K1 embedding-row gather and K10 deterministic scatter-add (stable LSD radix sort of
// (row id, token) + multi-level segmented reduction) for sm_100a.
//
// Reference call sites replaced:
//   tf.nn.embedding_lookup            Embedding/Behavior_embedding_time_aware_attention.py:68,75,82,90
//   its gradient (IndexedSlices ->    Model/base_model.py:292,296
//   unsorted_segment_sum in Adam)
//
// Both kernels are HBM-bound byte movers: 128-bit (gather) / full-line (reduce) coalesced access,
// grids sized in multiples of the SM count, no tensor-core reshaping.
#include "common.cuh"
#include "kernels.h"
#include "../../include/mtam.h"

namespace mtam {

// =============================================================================================
// gather: out[i,:] = table[idx[i],:]
// One float4 per thread; the D/4 threads of a row read the same idx (broadcast).  UNROLL rows
// in flight per thread to cover HBM latency.  Table reads use the default (L1/L2 allocating)
// path because hot rows repeat (Zipf ids, pad id 0); output is written streaming.
// =============================================================================================
template <int UNROLL>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float4* __restrict__ table,
                                                          const int32_t* __restrict__ idx, int64_t n,
                                                          int vpr /* float4 per row */,
                                                          float4* __restrict__ out) {
  const int64_t total = n * vpr;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; g + (UNROLL - 1) * stride < total; g += UNROLL * stride) {
    float4 v[UNROLL];
    int32_t r[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) r[u] = __ldg(idx + (g + u * stride) / vpr);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      v[u] = __ldg(table + (int64_t)r[u] * vpr + (int)((g + u * stride) % vpr));
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) __stcs(out + g + u * stride, v[u]);
  }
  for (; g < total; g += stride) {
    int32_t r = __ldg(idx + g / vpr);
    __stcs(out + g, __ldg(table + (int64_t)r * vpr + (int)(g % vpr)));
  }
}

int gather_rows(const float* table, int D, const int32_t* idx, int64_t n, float* out, cudaStream_t st) {
  if (n <= 0) return 0;
  if (D % 4 != 0) return set_error(MTAM_ERR_INVALID, "gather: D=%d must be a multiple of 4", D);
  int vpr = D / 4;
  int64_t total = n * vpr;
  int blocks = (int)std::min<int64_t>((total + 256 * 4 - 1) / (256 * 4), (int64_t)kNumSMs * 8);
  blocks = std::max(blocks, 1);
  gather_rows_kernel<4><<<blocks, 256, 0, st>>>((const float4*)table, idx, n, vpr, (float4*)out);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// exclusive scan of int32 (3-phase, deterministic)
// =============================================================================================
constexpr int SC_THREADS = 1024, SC_ITEMS = 4, SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ int block_excl_scan_1024(int v, int* sm /*>=33 ints*/, int* total) {
  // exclusive scan of one int per thread across a 1024-thread block
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) sm[w] = inc;
  __syncthreads();
  if (w == 0) {
    int s = sm[lane];
    int si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) si += t;
    }
    sm[lane] = si - s;
    if (lane == 31) sm[32] = si;
  }
  __syncthreads();
  *total = sm[32];
  return inc - v + sm[w];
}

__global__ void __launch_bounds__(SC_THREADS) scan_reduce_kernel(const int* in, int64_t n, int* block_sums) {
  __shared__ int sm[33];
  int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
  int s = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; ++i)
    if (base + i < n) s += in[base + i];
  int tot;
  block_excl_scan_1024(s, sm, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

// single block: in-place exclusive scan of `a[0..m)`, carry across chunks of 1024; writes total to *total_out
__global__ void __launch_bounds__(SC_THREADS) scan_small_kernel(int* a, int m, int* total_out) {
  __shared__ int sm[33];
  int carry = 0;
  for (int base = 0; base < m; base += SC_THREADS) {
    int i = base + threadIdx.x;
    int v = i < m ? a[i] : 0;
    int tot;
    int ex = block_excl_scan_1024(v, sm, &tot);
    if (i < m) a[i] = ex + carry;
    carry += tot;
    __syncthreads();
  }
  if (total_out && threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(SC_THREADS) scan_apply_kernel(const int* in, int* out, int64_t n,
                                                                 const int* block_offs) {
  __shared__ int sm[33];
  int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * SC_ITEMS;
  int v[SC_ITEMS];
  int s = 0;
#pragma unroll
  for (int i = 0; i < SC_ITEMS; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  int tot;
  int ex = block_excl_scan_1024(s, sm, &tot) + block_offs[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SC_ITEMS; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
}

size_t scan_tmp_ints(int64_t n) { return (size_t)cdiv(n, SC_TILE) + 1; }

// out may alias in.  tmp: scan_tmp_ints(n) ints.  total_out (device, optional) = sum of all.
int exclusive_scan_i32(const int* in, int* out, int64_t n, int* tmp, int* total_out, cudaStream_t st) {
  if (n <= 0) {
    if (total_out) MTAM_CUDA_CHECK(cudaMemsetAsync(total_out, 0, sizeof(int), st));
    return 0;
  }
  if (n <= 16 * SC_THREADS && in == out) {   // short arrays: one block, one launch
    scan_small_kernel<<<1, SC_THREADS, 0, st>>>(out, (int)n, total_out);
    MTAM_LAUNCH_CHECK();
    return 0;
  }
  int nb = cdiv(n, SC_TILE);
  scan_reduce_kernel<<<nb, SC_THREADS, 0, st>>>(in, n, tmp);
  scan_small_kernel<<<1, SC_THREADS, 0, st>>>(tmp, nb, total_out);
  scan_apply_kernel<<<nb, SC_THREADS, 0, st>>>(in, out, n, tmp);
  MTAM_LAUNCHES(2);
  MTAM_LAUNCH_CHECK();
  return 0;
}

// =============================================================================================
// stable LSD radix sort of (key=row id, val=token index), 8 bits per pass
// =============================================================================================
constexpr int RS_THREADS = 256, RS_ITEMS = 8, RS_TILE = RS_THREADS * RS_ITEMS, RS_WARPS = RS_THREADS / 32;

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const int32_t* __restrict__ keys, int64_t n,
                                                             int shift, int* __restrict__ hist, int nblk) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  int64_t base = (int64_t)blockIdx.x * RS_TILE;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < RS_ITEMS; ++i) {
    int64_t j = base + i * RS_THREADS + threadIdx.x;
    int d = (j < n) ? ((keys[j] >> shift) & 255) : 256;
    unsigned m = __match_any_sync(0xffffffffu, d);           // one shared-memory atomic per distinct digit per warp
    if (d < 256 && lane == (__ffs(m) - 1)) atomicAdd(&h[d], __popc(m));
  }
  __syncthreads();
  hist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];
}

// Stable: an item's destination = (scanned count of its digit before this block) + (same digit in
// earlier warps of the block) + (same digit earlier in this warp's contiguous segment).
__global__ void __launch_bounds__(RS_THREADS) rs_scatter_kernel(const int32_t* __restrict__ keys_in,
                                                                const int32_t* __restrict__ vals_in,
                                                                int32_t* __restrict__ keys_out,
                                                                int32_t* __restrict__ vals_out, int64_t n,
                                                                int shift, const int* __restrict__ offs,
                                                                int nblk) {
  __shared__ int wcnt[RS_WARPS][257];
  for (int i = threadIdx.x; i < RS_WARPS * 257; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
  __syncthreads();
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * (32 * RS_ITEMS);
  int32_t key[RS_ITEMS];
  int rank[RS_ITEMS], dig[RS_ITEMS];
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    int64_t j = wbase + r * 32 + lane;
    bool valid = j < n;
    int32_t k = valid ? keys_in[j] : 0;
    int d = valid ? ((k >> shift) & 255) : 256;
    unsigned m = __match_any_sync(0xffffffffu, d);
    int old = wcnt[w][d];
    __syncwarp();
    if (lane == (__ffs(m) - 1)) wcnt[w][d] = old + __popc(m);
    __syncwarp();
    rank[r] = old + __popc(m & ((1u << lane) - 1u));
    key[r] = k;
    dig[r] = d;
  }
  __syncthreads();
  {
    int d = threadIdx.x;  // 256 threads <-> 256 digits
    int run = offs[d * nblk + blockIdx.x];
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ++ww) {
      int c = wcnt[ww][d];
      wcnt[ww][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_ITEMS; ++r) {
    int64_t j = wbase + r * 32 + lane;
    if (j < n) {
      int dst = wcnt[w][dig[r]] + rank[r];
      keys_out[dst] = key[r];
      vals_out[dst] = vals_in ? vals_in[j] : (int32_t)j;
    }
  }
}

static int radix_passes(int table_rows) {
  int bits = 1;
  while ((1ll << bits) < (long long)table_rows) ++bits;
  return (bits + 7) / 8;
}

struct SortPlan {
  int64_t n;
  int nblk, passes;
  size_t hist_ints, scan_ints;
};
static SortPlan sort_plan(int64_t n, int table_rows) {
  SortPlan p;
  p.n = n;
  p.nblk = cdiv(n, RS_TILE);
  p.passes = radix_passes(table_rows);
  p.hist_ints = (size_t)256 * p.nblk;
  p.scan_ints = scan_tmp_ints((int64_t)p.hist_ints);
  return p;
}

size_t sort_workspace_bytes(int64_t n, int table_rows) {
  SortPlan p = sort_plan(n, table_rows);
  Bump b(nullptr, 0);
  b.take<int32_t>(n); b.take<int32_t>(n); b.take<int32_t>(n); b.take<int32_t>(n);
  b.take<int>(p.hist_ints); b.take<int>(p.scan_ints);
  return b.off + 256;
}

// Sorts idx[0..n) ascending, stable.  On return *keys_sorted / *perm point into the workspace.
int sort_by_row(const int32_t* idx, int64_t n, int table_rows, void* ws, size_t ws_bytes,
                const int32_t** keys_sorted, const int32_t** perm, cudaStream_t st) {
  SortPlan p = sort_plan(n, table_rows);
  Bump b(ws, ws_bytes);
  int32_t* ka = b.take<int32_t>(n);
  int32_t* kb = b.take<int32_t>(n);
  int32_t* va = b.take<int32_t>(n);
  int32_t* vb = b.take<int32_t>(n);
  int* hist = b.take<int>(p.hist_ints);
  int* stmp = b.take<int>(p.scan_ints);
  if (!b.ok()) return set_error(MTAM_ERR_WORKSPACE, "sort: workspace %zu < %zu", ws_bytes, b.off);
  const int32_t* kin = idx;
  const int32_t* vin = nullptr;
  int32_t* kout = ka;
  int32_t* vout = va;
  for (int pass = 0; pass < p.passes; ++pass) {
    int shift = pass * 8;
    rs_hist_kernel<<<p.nblk, RS_THREADS, 0, st>>>(kin, n, shift, hist, p.nblk);
    MTAM_TRY(exclusive_scan_i32(hist, hist, (int64_t)p.hist_ints, stmp, nullptr, st));
    rs_scatter_kernel<<<p.nblk, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, hist, p.nblk);
    MTAM_LAUNCHES(1);
    MTAM_LAUNCH_CHECK();
    kin = kout;
    vin = vout;
    kout = (kout == ka) ? kb : ka;
    vout = (vout == va) ? vb : va;
  }
  *keys_sorted = kin;
  *perm = vin;
  return 0;
}

// =============================================================================================
// multi-level segmented reduction over sorted keys
// A warp owns a tile of SR_CH consecutive sorted entries (lanes across D, full-line coalesced row
// reads).  Runs that end inside the tile are added to dst directly (at most one such add per key
// per level); the tile's last run is carried to the next level, whose input is the dense, still
// sorted list of carries.  Summation order is a pure function of the sorted order: deterministic.
// =============================================================================================
constexpr int SR_CH = 64;
constexpr int SR_WARPS = 4;
constexpr int SR_G = 8;   // entries whose row loads are in flight together

template <int VEC> struct VecT;
template <> struct VecT<1> { using T = float; };
template <> struct VecT<2> { using T = float2; };
template <> struct VecT<4> { using T = float4; };

__device__ __forceinline__ void vzero(float& a) { a = 0.f; }
__device__ __forceinline__ void vzero(float2& a) { a = make_float2(0.f, 0.f); }
__device__ __forceinline__ void vzero(float4& a) { a = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vadd(float& a, const float& b) { a += b; }
__device__ __forceinline__ void vadd(float2& a, const float2& b) { a.x += b.x; a.y += b.y; }
__device__ __forceinline__ void vadd(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
// fire-and-forget reduction into global memory (RED.ADD, no return value, no load latency on the
// critical path).  Deterministic here because a destination row receives at most one add per level.
__device__ __forceinline__ void vred(float* p, const float& a) { atomicAdd(p, a); }
__device__ __forceinline__ void vred(float* p, const float2& a) { atomicAdd(reinterpret_cast<float2*>(p), a); }
__device__ __forceinline__ void vred(float* p, const float4& a) { atomicAdd(reinterpret_cast<float4*>(p), a); }

// D == 32*VEC: each lane owns VEC contiguous floats of a row (one 128-bit / 64-bit / 32-bit access).
// A CTA of SRV_WARPS warps owns SRV_WARPS consecutive 64-entry tiles.  Every warp reduces its tile: runs that begin
// and end inside the tile are added to dst at once; the tile's first run ("head", it may continue the previous
// tile's last run) and last run ("tail") go to shared memory, where warp 0 stitches the tiles together in order,
// adds every run that ends inside the CTA to dst (one add per key per CTA) and carries the CTA's last run to the
// next level.  So a level shrinks the list 64*SRV_WARPS-fold.
// Entries per warp: 64 on the big first level (8 groups of 8 rows in flight), 8 on the short carry lists of the later
// levels, where one group per warp keeps the launch a single round of loads deep.
constexpr int SRV_WARPS = 8;

template <int VEC, int CH>
__global__ void __launch_bounds__(SRV_WARPS * 32) seg_reduce_vec_kernel(
    const int32_t* __restrict__ keys, const float* __restrict__ src, int ld_src,
    const int32_t* __restrict__ perm, int64_t n, int ld_dst, float* __restrict__ dst,
    int32_t* __restrict__ carry_keys, float* __restrict__ carry_rows) {
  using V = typename VecT<VEC>::T;
  constexpr int D = 32 * VEC;
  __shared__ V s_head[SRV_WARPS][32], s_tail[SRV_WARPS][32];
  __shared__ int32_t s_hkey[SRV_WARPS], s_tkey[SRV_WARPS];
  __shared__ int s_single[SRV_WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t tile = (int64_t)blockIdx.x * SRV_WARPS + w;
  const int64_t ntiles = (n + CH - 1) / CH;
  if (tile < ntiles) {
    const int64_t start = tile * CH;
    const int cnt = (int)min((int64_t)CH, n - start);
    // tile keys / source rows: two coalesced loads, broadcast later with shuffles
    const int l0 = min(lane, cnt - 1), l1 = min(lane + 32, cnt - 1);   // clamp: entries past the end alias the last one
    int32_t k0 = keys[start + l0], k1 = keys[start + l1];
    int32_t r0 = perm ? perm[start + l0] : (int32_t)(start + l0);
    int32_t r1 = perm ? perm[start + l1] : (int32_t)(start + l1);
    V acc, head;
    vzero(acc);
    vzero(head);
    int32_t cur = __shfl_sync(0xffffffffu, k0, 0);
    const int32_t hkey = cur;
    bool first_run = true;
    for (int i0 = 0; i0 < cnt; i0 += SR_G) {
      V x[SR_G];
      int32_t kk[SR_G];
#pragma unroll
      for (int u = 0; u < SR_G; ++u) {     // unconditional, independent loads: SR_G rows in flight
        int i = i0 + u;
        int32_t ka = __shfl_sync(0xffffffffu, k0, i & 31), kb = __shfl_sync(0xffffffffu, k1, i & 31);
        int32_t ra = __shfl_sync(0xffffffffu, r0, i & 31), rb = __shfl_sync(0xffffffffu, r1, i & 31);
        kk[u] = (i < 32) ? ka : kb;
        int32_t row = (i < 32) ? ra : rb;
        x[u] = __ldg(reinterpret_cast<const V*>(src + (int64_t)row * ld_src) + lane);
      }
#pragma unroll
      for (int u = 0; u < SR_G; ++u) {
        if (i0 + u < cnt) {
          if (kk[u] != cur) {              // warp-uniform: the run of `cur` ended
            if (first_run) { head = acc; first_run = false; }
            else vred(dst + (int64_t)cur * ld_dst + lane * VEC, acc);
            vzero(acc);
            cur = kk[u];
          }
          vadd(acc, x[u]);
        }
      }
    }
    s_head[w][lane] = head;
    s_tail[w][lane] = acc;
    if (lane == 0) { s_hkey[w] = hkey; s_tkey[w] = cur; s_single[w] = first_run ? 1 : 0; }
  }
  __syncthreads();
  if (w != 0) return;
  const int nw = (int)min((int64_t)SRV_WARPS, ntiles - (int64_t)blockIdx.x * SRV_WARPS);
  V cacc;
  vzero(cacc);
  int32_t ckey = 0;
  bool cvalid = false;
  for (int q = 0; q < nw; ++q) {
    const int32_t hk = s_hkey[q], tk = s_tkey[q];
    const V ta = s_tail[q][lane];
    if (s_single[q]) {                      // the whole tile is one run
      if (cvalid && tk == ckey) vadd(cacc, ta);
      else {
        if (cvalid) vred(dst + (int64_t)ckey * ld_dst + lane * VEC, cacc);
        ckey = tk; cacc = ta; cvalid = true;
      }
    } else {
      V ha = s_head[q][lane];
      if (cvalid && hk == ckey) {           // the head run continues the pending run and ends here
        vadd(cacc, ha);
        vred(dst + (int64_t)ckey * ld_dst + lane * VEC, cacc);
      } else {
        if (cvalid) vred(dst + (int64_t)ckey * ld_dst + lane * VEC, cacc);
        vred(dst + (int64_t)hk * ld_dst + lane * VEC, ha);
      }
      ckey = tk; cacc = ta; cvalid = true;
    }
  }
  const int64_t nctas = (ntiles + SRV_WARPS - 1) / SRV_WARPS;
  if (blockIdx.x == nctas - 1) {
    vred(dst + (int64_t)ckey * ld_dst + lane * VEC, cacc);
  } else {
    if (lane == 0) carry_keys[blockIdx.x] = ckey;
    reinterpret_cast<V*>(carry_rows + (int64_t)blockIdx.x * D)[lane] = cacc;
  }
}

// any D <= 256: lanes stride across the row
template <int MAXV>
__global__ void __launch_bounds__(SR_WARPS * 32) seg_reduce_level_kernel(
    const int32_t* __restrict__ keys, const float* __restrict__ src, int ld_src,
    const int32_t* __restrict__ perm, int64_t n, int D, int ld_dst, float* __restrict__ dst,
    int32_t* __restrict__ carry_keys, float* __restrict__ carry_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * SR_WARPS + (threadIdx.x >> 5);
  const int64_t ntiles = (n + SR_CH - 1) / SR_CH;
  if (tile >= ntiles) return;
  const int64_t start = tile * SR_CH;
  const int cnt = (int)min((int64_t)SR_CH, n - start);
  const int l0 = min(lane, cnt - 1), l1 = min(lane + 32, cnt - 1);
  int32_t k0 = keys[start + l0], k1 = keys[start + l1];
  int32_t r0 = perm ? perm[start + l0] : (int32_t)(start + l0);
  int32_t r1 = perm ? perm[start + l1] : (int32_t)(start + l1);
  float acc[MAXV];
#pragma unroll
  for (int v = 0; v < MAXV; ++v) acc[v] = 0.f;
  int32_t cur = __shfl_sync(0xffffffffu, k0, 0);
  for (int i0 = 0; i0 < cnt; i0 += 4) {
    float x[4][MAXV];
    int32_t kk[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int i = i0 + u;
      int32_t ka = __shfl_sync(0xffffffffu, k0, i & 31), kb = __shfl_sync(0xffffffffu, k1, i & 31);
      int32_t ra = __shfl_sync(0xffffffffu, r0, i & 31), rb = __shfl_sync(0xffffffffu, r1, i & 31);
      kk[u] = (i < 32) ? ka : kb;
      int32_t row = (i < 32) ? ra : rb;
      const float* p = src + (int64_t)row * ld_src;
#pragma unroll
      for (int v = 0; v < MAXV; ++v) x[u][v] = __ldg(p + min(lane + 32 * v, D - 1));   // clamped, masked at use
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + u < cnt) {
        if (kk[u] != cur) {
#pragma unroll
          for (int v = 0; v < MAXV; ++v) {
            int d = lane + 32 * v;
            if (d < D) atomicAdd(dst + (int64_t)cur * ld_dst + d, acc[v]);
            acc[v] = 0.f;
          }
          cur = kk[u];
        }
#pragma unroll
        for (int v = 0; v < MAXV; ++v) acc[v] += x[u][v];
      }
    }
  }
  if (tile == ntiles - 1) {
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
      int d = lane + 32 * v;
      if (d < D) atomicAdd(dst + (int64_t)cur * ld_dst + d, acc[v]);
    }
  } else {
    if (lane == 0) carry_keys[tile] = cur;
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
      int d = lane + 32 * v;
      if (d < D) carry_rows[tile * D + d] = acc[v];
    }
  }
}

size_t seg_reduce_workspace_bytes(int64_t n, int D) {
  Bump b(nullptr, 0);
  int64_t m = n;
  while (true) {
    int64_t nt = (m + SR_CH - 1) / SR_CH;
    if (nt <= 1) break;
    b.take<int32_t>(nt - 1);
    b.take<float>((nt - 1) * D);
    m = nt - 1;
  }
  return b.off + 256;
}

// dst[keys[i],:] += src[perm ? perm[i] : i, :]  over sorted keys
int seg_reduce_sorted(const int32_t* keys_sorted, const int32_t* perm, const float* src, int ld_src, int64_t n,
                      int D, float* dst, int ld_dst, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n <= 0) return 0;
  if (D > 256) return set_error(MTAM_ERR_INVALID, "seg_reduce: D=%d > 256", D);
  Bump b(ws, ws_bytes);
  const int32_t* k = keys_sorted;
  const int32_t* pm = perm;
  const float* s = src;
  int lds = ld_src;
  int64_t m = n;
  while (m > 0) {
    int maxv = (D + 31) / 32;
    const bool vec_ok = (D == 32 || D == 64 || D == 128) && (lds % (D / 32) == 0) && (ld_dst % (D / 32) == 0) &&
                        ((uintptr_t)s % 16 == 0) && ((uintptr_t)dst % 16 == 0);
    const int wch = (m > 65536) ? 64 : 8;            // entries per warp of the vector kernel at this level
    const int ch = vec_ok ? wch * SRV_WARPS : SR_CH; // entries folded into one carry at this level
    int64_t nt = (m + ch - 1) / ch;
    int32_t* ck = nullptr;
    float* cr = nullptr;
    if (nt > 1) {
      ck = b.take<int32_t>(nt - 1);
      cr = b.take<float>((nt - 1) * D);
      if (!b.ok()) return set_error(MTAM_ERR_WORKSPACE, "seg_reduce: workspace %zu < %zu", ws_bytes, b.off);
    }
    int blocks = vec_ok ? (int)nt : cdiv(nt, SR_WARPS);
#define SRV_LAUNCH(VEC_)                                                                                              \
  do {                                                                                                                \
    if (wch == 64) seg_reduce_vec_kernel<VEC_, 64><<<blocks, SRV_WARPS * 32, 0, st>>>(k, s, lds, pm, m, ld_dst, dst, ck, cr); \
    else seg_reduce_vec_kernel<VEC_, 8><<<blocks, SRV_WARPS * 32, 0, st>>>(k, s, lds, pm, m, ld_dst, dst, ck, cr);    \
  } while (0)
    if (vec_ok && D == 32) SRV_LAUNCH(1);
    else if (vec_ok && D == 64) SRV_LAUNCH(2);
    else if (vec_ok && D == 128) SRV_LAUNCH(4);
    else if (maxv <= 1)
      seg_reduce_level_kernel<1><<<blocks, SR_WARPS * 32, 0, st>>>(k, s, lds, pm, m, D, ld_dst, dst, ck, cr);
    else if (maxv <= 2)
      seg_reduce_level_kernel<2><<<blocks, SR_WARPS * 32, 0, st>>>(k, s, lds, pm, m, D, ld_dst, dst, ck, cr);
    else if (maxv <= 4)
      seg_reduce_level_kernel<4><<<blocks, SR_WARPS * 32, 0, st>>>(k, s, lds, pm, m, D, ld_dst, dst, ck, cr);
    else
      seg_reduce_level_kernel<8><<<blocks, SR_WARPS * 32, 0, st>>>(k, s, lds, pm, m, D, ld_dst, dst, ck, cr);
#undef SRV_LAUNCH
    MTAM_LAUNCH_CHECK();
    if (nt <= 1) break;
    k = ck;
    pm = nullptr;
    s = cr;
    lds = D;
    m = nt - 1;
  }
  return 0;
}

// distinct sorted keys -> unique_idx, n_unique
__global__ void head_flags_kernel(const int32_t* keys, int64_t n, int* flags) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
__global__ void compact_heads_kernel(const int32_t* keys, int64_t n, const int* pos, int32_t* unique_idx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (i == 0 || keys[i] != keys[i - 1])) unique_idx[pos[i]] = keys[i];
}

size_t unique_workspace_bytes(int64_t n) {
  Bump b(nullptr, 0);
  b.take<int>(n);
  b.take<int>(scan_tmp_ints(n));
  return b.off + 256;
}

int unique_sorted(const int32_t* keys_sorted, int64_t n, void* ws, size_t ws_bytes, int32_t* unique_idx,
                  int32_t* n_unique, cudaStream_t st) {
  Bump b(ws, ws_bytes);
  int* flags = b.take<int>(n);
  int* tmp = b.take<int>(scan_tmp_ints(n));
  if (!b.ok()) return set_error(MTAM_ERR_WORKSPACE, "unique: workspace too small");
  int blocks = cdiv(n, 256);
  head_flags_kernel<<<blocks, 256, 0, st>>>(keys_sorted, n, flags);
  MTAM_TRY(exclusive_scan_i32(flags, flags, n, tmp, n_unique, st));
  compact_heads_kernel<<<blocks, 256, 0, st>>>(keys_sorted, n, flags, unique_idx);
  MTAM_LAUNCHES(1);
  MTAM_LAUNCH_CHECK();
  return 0;
}

size_t scatter_add_workspace_bytes(int64_t n, int table_rows, int D) {
  return align_up(sort_workspace_bytes(n, table_rows), 256) + align_up(seg_reduce_workspace_bytes(n, D), 256) +
         align_up(unique_workspace_bytes(n), 256) + 1024;
}

int scatter_add_rows(float* dst, int table_rows, int D, int ld_dst, const int32_t* idx, const float* rows,
                     int ld_src, int64_t n, void* ws, size_t ws_bytes, int32_t* unique_idx, int32_t* n_unique,
                     cudaStream_t st) {
  if (n <= 0) {
    if (n_unique) MTAM_CUDA_CHECK(cudaMemsetAsync(n_unique, 0, sizeof(int32_t), st));
    return 0;
  }
  size_t s1 = align_up(sort_workspace_bytes(n, table_rows), 256);
  size_t s2 = align_up(seg_reduce_workspace_bytes(n, D), 256);
  size_t s3 = align_up(unique_workspace_bytes(n), 256);
  if (ws_bytes < s1 + s2 + s3)
    return set_error(MTAM_ERR_WORKSPACE, "scatter_add: workspace %zu < %zu", ws_bytes, s1 + s2 + s3);
  char* w = (char*)ws;
  const int32_t *ks, *pm;
  MTAM_TRY(sort_by_row(idx, n, table_rows, w, s1, &ks, &pm, st));
  MTAM_TRY(seg_reduce_sorted(ks, pm, rows, ld_src, n, D, dst, ld_dst, w + s1, s2, st));
  if (unique_idx && n_unique) MTAM_TRY(unique_sorted(ks, n, w + s1 + s2, s3, unique_idx, n_unique, st));
  return 0;
}

}  // namespace mtam

// ---- C-ABI ---------------------------------------------------------------------------------
extern "C" int mtam_gather(const float* table, int32_t table_rows, int32_t D, const int32_t* idx, int64_t n,
                           float* out, void* stream) {
  (void)table_rows;
  if (!table || !idx || !out || n < 0) return mtam::set_error(MTAM_ERR_INVALID, "mtam_gather: null/negative argument");
  return mtam::gather_rows(table, D, idx, n, out, (cudaStream_t)stream);
}

extern "C" size_t mtam_scatter_add_workspace(int64_t n, int32_t table_rows, int32_t D) {
  return mtam::scatter_add_workspace_bytes(n, table_rows, D);
}

extern "C" int mtam_scatter_add(float* dst, int32_t table_rows, int32_t D, const int32_t* idx, const float* rows,
                                int32_t ld_rows, int64_t n, void* workspace, size_t workspace_bytes, int32_t* unique_idx,
                                int32_t* n_unique, void* stream) {
  if (!dst || n < 0 || (n > 0 && (!idx || !rows || !workspace)))
    return mtam::set_error(MTAM_ERR_INVALID, "mtam_scatter_add: null/negative argument");
  if (ld_rows < D) return mtam::set_error(MTAM_ERR_INVALID, "mtam_scatter_add: ld_rows < D");
  return mtam::scatter_add_rows(dst, table_rows, D, D, idx, rows, ld_rows, n, workspace, workspace_bytes, unique_idx,
                                n_unique, (cudaStream_t)stream);
}

extern "C" size_t mtam_sort_workspace(int64_t n, int32_t key_bound) { return mtam::sort_workspace_bytes(n, key_bound); }

extern "C" int mtam_sort_indices(const int32_t* keys, int64_t n, int32_t key_bound, void* workspace, size_t workspace_bytes,
                                 const int32_t** keys_sorted, const int32_t** perm, void* stream) {
  if (n < 0 || key_bound <= 0 || !keys_sorted || !perm || (n > 0 && (!keys || !workspace)))
    return mtam::set_error(MTAM_ERR_INVALID, "mtam_sort_indices: bad argument");
  if (n == 0) {
    *keys_sorted = *perm = nullptr;
    return 0;
  }
  return mtam::sort_by_row(keys, n, key_bound, workspace, workspace_bytes, keys_sorted, perm, (cudaStream_t)stream);
}

extern "C" size_t mtam_scatter_add_sorted_workspace(int64_t n, int32_t D) { return mtam::seg_reduce_workspace_bytes(n, D); }

extern "C" int mtam_scatter_add_sorted(float* dst, int32_t D, const int32_t* keys_sorted, const int32_t* perm,
                                       const float* rows, int32_t ld_rows, int64_t n, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  if (!dst || n < 0 || (n > 0 && (!keys_sorted || !rows || !workspace)))
    return mtam::set_error(MTAM_ERR_INVALID, "mtam_scatter_add_sorted: null/negative argument");
  if (ld_rows < D) return mtam::set_error(MTAM_ERR_INVALID, "mtam_scatter_add_sorted: ld_rows < D");
  return mtam::seg_reduce_sorted(keys_sorted, perm, rows, ld_rows, n, D, dst, D, workspace, workspace_bytes,
                                 (cudaStream_t)stream);
}
