// Stable radix sort of (key = table row id, value = token index) -- the first half of the deterministic
// scatter-add (K10: the IndexedSlices aggregation of Model/base_model.py:292, tf's unsorted_segment_sum).
//
// One sweep per 8-bit digit: ONE histogram kernel reads the keys once and counts every digit position; each pass
// is then a single kernel whose CTAs (tiles of 8192 keys, handed out in launch order by a ticket counter) rank their
// keys, publish their per-digit counts, and pick up the counts of the tiles before them through flagged words in
// global memory (a two-level "decoupled look-back": the tiles of the same group of 16 in one batch of loads, then the
// totals of the earlier groups, stopping at the first group that has already published an inclusive prefix).  A tile only ever waits for tiles that took their ticket earlier, which are
// resident and publish before they wait, so the scheme cannot deadlock however the CTAs are scheduled.
// Per pass: one read of (keys, values), one scattered write.  No per-pass histogram or scan kernels.
// Stable: destination = digit base + keys of the same digit in earlier tiles + in earlier warps of the tile + earlier
// in the warp's own contiguous segment.
#include "common.cuh"
#include "kernels.h"
#include "../../include/mtam.h"

namespace mtam {

namespace {

constexpr int OS_THREADS = 512, OS_WARPS = OS_THREADS / 32;
constexpr int OS_RADIX = 256, OS_MAX_PASSES = 4;
constexpr uint32_t FLAG_AGG = 1u, FLAG_INCL = 2u;
constexpr int OS_GROUP = 16;            // tiles per look-back group = flag words in flight per thread

struct SortHeader {
  int ghist[OS_MAX_PASSES][OS_RADIX];   // global digit histograms
  int ticket[OS_MAX_PASSES];            // next tile id of each pass
  int pad[60];
};

__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }

// lanes of the warp holding the same 9-bit value (8-bit digit or the out-of-range marker): 9 ballots -- the match.any
// instruction is an order of magnitude slower
__device__ __forceinline__ unsigned match_digit(int d) {
  unsigned m = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 9; ++b) {
    const bool bit = (d >> b) & 1;
    const unsigned bal = __ballot_sync(0xffffffffu, bit);
    m &= bit ? bal : ~bal;
  }
  return m;
}

constexpr int OH_ITEMS = 8;             // keys per thread of the histogram kernel: all loads issued up front
__global__ void __launch_bounds__(OS_THREADS) os_hist_kernel(const int32_t* __restrict__ keys, int64_t n, int passes,
                                                             SortHeader* __restrict__ hdr) {
  __shared__ int h[OS_MAX_PASSES][OS_RADIX];
  for (int i = threadIdx.x; i < OS_MAX_PASSES * OS_RADIX; i += OS_THREADS) (&h[0][0])[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // few CTAs (every one ends with passes*256 global atomics), each streaming its share OH_ITEMS loads at a time
  for (int64_t base = (int64_t)blockIdx.x * (OS_THREADS * OH_ITEMS); base < n;
       base += (int64_t)gridDim.x * (OS_THREADS * OH_ITEMS)) {
    int32_t k[OH_ITEMS];
#pragma unroll
    for (int i = 0; i < OH_ITEMS; ++i) {
      const int64_t j = base + i * OS_THREADS + threadIdx.x;
      k[i] = (j < n) ? __ldg(keys + j) : -1;
    }
#pragma unroll
    for (int i = 0; i < OH_ITEMS; ++i) {
      const bool valid = k[i] >= 0;
      // warps filled with one key (the pad id) count it once; otherwise one shared-memory atomic per key and digit
      const int32_t k0 = __shfl_sync(0xffffffffu, k[i], 0);
      if (__all_sync(0xffffffffu, k[i] == k0)) {
        if (valid && lane < passes) atomicAdd(&h[lane][(k[i] >> (8 * lane)) & 255], 32);
      } else if (valid) {
        for (int p = 0; p < passes; ++p) atomicAdd(&h[p][(k[i] >> (8 * p)) & 255], 1);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < passes * OS_RADIX; i += OS_THREADS) {
    const int c = (&h[0][0])[i];
    if (c) atomicAdd(&hdr->ghist[0][0] + i, c);
  }
}

template <bool FIRST, int OS_ITEMS>
__global__ void __launch_bounds__(OS_THREADS, 2) os_pass_kernel(const int32_t* __restrict__ keys_in,
                                                             const int32_t* __restrict__ vals_in,
                                                             int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out,
                                                             int64_t n, int shift, const int* __restrict__ ghist,
                                                             int* __restrict__ ticket, uint32_t* __restrict__ state,
                                                             int ntiles) {
  constexpr int OS_TILE = OS_THREADS * OS_ITEMS;
  __shared__ int32_t sbuf[OS_TILE];                   // the tile in digit order: keys, then values
  __shared__ uint16_t wcnt[OS_WARPS][OS_RADIX + 2];   // per-warp digit counts -> offsets of the warp inside the digit
  __shared__ int delta_s[OS_RADIX];                   // global position - position in the tile, per digit
  __shared__ int lstart_s[OS_RADIX];                  // first position of the digit inside the tile
  __shared__ int wsum[2][OS_WARPS];
  __shared__ int tile_s;
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  if (tid == 0) tile_s = atomicAdd(ticket, 1);
  for (int i = tid; i < OS_WARPS * (OS_RADIX + 2); i += OS_THREADS) (&wcnt[0][0])[i] = 0;
  __syncthreads();
  const int tile = tile_s;
  const int64_t tbase = (int64_t)tile * OS_TILE;
  const int64_t wbase = tbase + (int64_t)w * (32 * OS_ITEMS);
  const int tile_n = (int)min((int64_t)OS_TILE, n - tbase);
  int32_t key[OS_ITEMS];
  int pos[OS_ITEMS];                                  // digit | rank-in-warp << 9, later the position in the tile
#pragma unroll
  for (int r = 0; r < OS_ITEMS; ++r) {
    const int64_t j = wbase + r * 32 + lane;
    key[r] = (j < n) ? __ldg(keys_in + j) : 0;
  }
  // the values travel with the keys from the start: loading them only when they are placed costs one more exposed
  // memory latency per pass
  int32_t val[OS_ITEMS];
#pragma unroll
  for (int r = 0; r < OS_ITEMS; ++r) {
    const int64_t j = wbase + r * 32 + lane;
    val[r] = FIRST ? (int32_t)j : ((j < n) ? __ldg(vals_in + j) : 0);
  }
#pragma unroll
  for (int r = 0; r < OS_ITEMS; ++r) {
    const int64_t j = wbase + r * 32 + lane;
    const int d = (j < n) ? ((key[r] >> shift) & 255) : OS_RADIX;
    const unsigned m = match_digit(d);
    const int old = wcnt[w][d];
    __syncwarp();
    if (lane == (__ffs(m) - 1)) wcnt[w][d] = (uint16_t)(old + __popc(m));
    __syncwarp();
    pos[r] = d | ((old + __popc(m & ((1u << lane) - 1u))) << 9);
  }
  __syncthreads();
  uint32_t excl = 0;
  int total = 0;
  if (tid < OS_RADIX) {
    const int d = tid;
    int run = 0;
#pragma unroll
    for (int ww = 0; ww < OS_WARPS; ++ww) {
      const int c = wcnt[ww][d];
      wcnt[ww][d] = (uint16_t)run;
      run += c;
    }
    total = run;
    // prefix over the earlier tiles in two levels, so that the chain of dependent round trips stays 2-3 long however
    // many tiles run at once: (1) the counts of the earlier tiles of this tile's group of OS_GROUP, one batch of loads;
    // (2) the totals of the earlier groups (published by each group's last tile), batches of OS_GROUP with an
    // inclusive-prefix shortcut published by each group's last tile once it knows its own
    uint32_t* tstate = state;
    uint32_t* gstate = state + (size_t)ntiles * OS_RADIX;
    st_flag(tstate + (size_t)tile * OS_RADIX + d, ((uint32_t)total << 2) | FLAG_AGG);
    const int grp = tile / OS_GROUP, first = grp * OS_GROUP;
    uint32_t v[OS_GROUP];
#pragma unroll
    for (int q = 0; q < OS_GROUP; ++q)
      v[q] = (first + q < tile) ? ld_flag(tstate + (size_t)(first + q) * OS_RADIX + d) : FLAG_AGG;
#pragma unroll
    for (int q = 0; q < OS_GROUP; ++q) {
      while ((v[q] & 3u) == 0u) v[q] = ld_flag(tstate + (size_t)(first + q) * OS_RADIX + d);
      excl += v[q] >> 2;
    }
    const bool closes = tile == first + OS_GROUP - 1;        // the group's last tile (a partial last group has none)
    const uint32_t gtotal = excl + (uint32_t)total;
    if (closes) st_flag(gstate + (size_t)grp * OS_RADIX + d, (gtotal << 2) | (grp == 0 ? FLAG_INCL : FLAG_AGG));
    uint32_t gexcl = 0;
    int t = grp - 1;
    bool done = t < 0;
    while (!done) {
#pragma unroll
      for (int q = 0; q < OS_GROUP; ++q) v[q] = (t - q >= 0) ? ld_flag(gstate + (size_t)(t - q) * OS_RADIX + d) : FLAG_INCL;
#pragma unroll
      for (int q = 0; q < OS_GROUP; ++q) {
        if (!done) {
          while ((v[q] & 3u) == 0u) v[q] = ld_flag(gstate + (size_t)(t - q) * OS_RADIX + d);
          gexcl += v[q] >> 2;
          done = (v[q] & 3u) == FLAG_INCL;
        }
      }
      t -= OS_GROUP;
    }
    if (closes && grp > 0) st_flag(gstate + (size_t)grp * OS_RADIX + d, ((gexcl + gtotal) << 2) | FLAG_INCL);
    excl += gexcl;
  }
  // two exclusive scans over the 256 digits: the global histogram (digit bases) and this tile's counts
  {
    const int g = (tid < OS_RADIX) ? ghist[tid] : 0;
    int ig = g, it = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, ig, o), b = __shfl_up_sync(0xffffffffu, it, o);
      if (lane >= o) { ig += a; it += b; }
    }
    if (lane == 31) { wsum[0][w] = ig; wsum[1][w] = it; }
    __syncthreads();
    int bg = 0, bt = 0;
#pragma unroll
    for (int ww = 0; ww < OS_RADIX / 32; ++ww) {
      bg += (ww < w) ? wsum[0][ww] : 0;
      bt += (ww < w) ? wsum[1][ww] : 0;
    }
    if (tid < OS_RADIX) {
      const int lstart = bt + it - total;
      lstart_s[tid] = lstart;
      delta_s[tid] = bg + ig - g + (int)excl - lstart;
    }
  }
  __syncthreads();
  // keys into digit order inside the tile, then out in runs of consecutive addresses
#pragma unroll
  for (int r = 0; r < OS_ITEMS; ++r) {
    const int d = pos[r] & 511;
    if (d < OS_RADIX) {
      pos[r] = lstart_s[d] + wcnt[w][d] + (pos[r] >> 9);
      sbuf[pos[r]] = key[r];
    } else {
      pos[r] = -1;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < OS_ITEMS; ++r) {
    const int i = r * OS_THREADS + tid;
    if (i < tile_n) {
      const int32_t k = sbuf[i];
      key[r] = delta_s[(k >> shift) & 255] + i;     // destination, reused for the value
      keys_out[key[r]] = k;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < OS_ITEMS; ++r)
    if (pos[r] >= 0) sbuf[pos[r]] = val[r];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < OS_ITEMS; ++r) {
    const int i = r * OS_THREADS + tid;
    if (i < tile_n) vals_out[key[r]] = sbuf[i];
  }
}

// flag words of one pass: one row of 256 per tile and per group of tiles
size_t state_words(int64_t ntiles) { return (size_t)(ntiles + (ntiles + OS_GROUP - 1) / OS_GROUP) * OS_RADIX; }

int radix_passes(int key_bound) {
  int bits = 1;
  while ((1ll << bits) < (long long)key_bound) ++bits;
  return (bits + 7) / 8;
}


// keys per thread of a pass: the smallest tile that still lets every tile be resident at once (2 CTAs per SM), so that
// one balanced wave covers the array; large arrays use the largest tile
int pass_items(int64_t n) {
  const int opts[4] = {4, 8, 11, 16};
  for (int i = 0; i < 4; ++i)
    if ((n + (int64_t)OS_THREADS * opts[i] - 1) / ((int64_t)OS_THREADS * opts[i]) <= 2 * kNumSMs) return opts[i];
  return 16;
}

}  // namespace

size_t sort_workspace_bytes(int64_t n, int key_bound) {
  const int passes = radix_passes(key_bound);
  const int64_t ntiles = (n + OS_THREADS * 4 - 1) / (OS_THREADS * 4);     // the smallest tile: the most tiles
  Bump b(nullptr, 0);
  b.take<int32_t>(n); b.take<int32_t>(n); b.take<int32_t>(n); b.take<int32_t>(n);
  b.take<char>(sizeof(SortHeader) + (size_t)passes * state_words(ntiles) * sizeof(uint32_t));
  return b.off + 256;
}

// Sorts idx[0..n) ascending, stable; values = original positions.  On return *keys_sorted / *perm point into the
// workspace.  Keys must lie in [0, key_bound).
int sort_by_row(const int32_t* idx, int64_t n, int key_bound, void* ws, size_t ws_bytes, const int32_t** keys_sorted,
                const int32_t** perm, cudaStream_t st) {
  if (n >= (1ll << 30)) return set_error(MTAM_ERR_INVALID, "sort: %lld keys exceed 2^30", (long long)n);
  const int passes = radix_passes(key_bound);
  const int items = pass_items(n);
  const int64_t tile = (int64_t)OS_THREADS * items;
  const int64_t ntiles = (n + tile - 1) / tile;
  Bump b(ws, ws_bytes);
  int32_t* ka = b.take<int32_t>(n);
  int32_t* kb = b.take<int32_t>(n);
  int32_t* va = b.take<int32_t>(n);
  int32_t* vb = b.take<int32_t>(n);
  const size_t ctl_bytes = sizeof(SortHeader) + (size_t)passes * state_words(ntiles) * sizeof(uint32_t);
  char* ctl = b.take<char>(ctl_bytes);
  if (!b.ok()) return set_error(MTAM_ERR_WORKSPACE, "sort: workspace %zu < %zu", ws_bytes, b.off);
  SortHeader* hdr = reinterpret_cast<SortHeader*>(ctl);
  uint32_t* state = reinterpret_cast<uint32_t*>(ctl + sizeof(SortHeader));
  MTAM_CUDA_CHECK(cudaMemsetAsync(ctl, 0, ctl_bytes, st));
  os_hist_kernel<<<std::min(cdiv(n, OS_THREADS * OH_ITEMS), kNumSMs), OS_THREADS, 0, st>>>(idx, n, passes, hdr);
  MTAM_LAUNCH_CHECK();
  const int32_t* kin = idx;
  const int32_t* vin = nullptr;
  int32_t* kout = ka;
  int32_t* vout = va;
  for (int p = 0; p < passes; ++p) {
    uint32_t* stp = state + (size_t)p * state_words(ntiles);
#define OS_PASS(FIRST_, ITEMS_)                                                                                     \
  os_pass_kernel<FIRST_, ITEMS_><<<(int)ntiles, OS_THREADS, 0, st>>>(kin, vin, kout, vout, n, 8 * p, hdr->ghist[p], \
                                                                      &hdr->ticket[p], stp, (int)ntiles)
#define OS_PASS_ITEMS(FIRST_)                        \
  do {                                               \
    if (items == 4) OS_PASS(FIRST_, 4);              \
    else if (items == 8) OS_PASS(FIRST_, 8);         \
    else if (items == 11) OS_PASS(FIRST_, 11);       \
    else OS_PASS(FIRST_, 16);                        \
  } while (0)
    if (p == 0) OS_PASS_ITEMS(true);
    else OS_PASS_ITEMS(false);
#undef OS_PASS_ITEMS
#undef OS_PASS
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

}  // namespace mtam
