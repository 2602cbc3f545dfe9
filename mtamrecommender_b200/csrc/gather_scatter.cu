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
// multi-level segmented reduction over sorted keys  (the sort itself: radix_sort.cu)
//
// Every destination row is written EXACTLY ONCE.  A CTA owns a contiguous piece of the sorted list; every run of equal
// keys that lies strictly inside the piece is reduced and written at once; the piece's first and last runs (which may
// continue in the neighbouring pieces) are carried -- as two (key, partial row) entries per CTA -- to the next level,
// whose input is that short, still sorted list.  The last level is a single CTA and writes everything.  Hence
//   * the result is a pure function of the sorted order (bit-reproducible), and
//   * the write can be a plain store (`accumulate == 0`: dst[key] = sum of its rows; rows of dst that no key names
//     are left alone) -- no read of dst at all -- or an add into the existing row (`accumulate != 0`).
// =============================================================================================
constexpr int SR_CH = 64;
constexpr int SR_WARPS = 4;

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
// the one write of a destination row: plain streaming store, or a fire-and-forget reduction (RED.ADD, no return value)
template <bool ACC> __device__ __forceinline__ void vput(float* p, const float& a) { if (ACC) atomicAdd(p, a); else *p = a; }
template <bool ACC> __device__ __forceinline__ void vput(float* p, const float2& a) {
  if (ACC) atomicAdd(reinterpret_cast<float2*>(p), a); else *reinterpret_cast<float2*>(p) = a;
}
template <bool ACC> __device__ __forceinline__ void vput(float* p, const float4& a) {
  if (ACC) atomicAdd(reinterpret_cast<float4*>(p), a); else *reinterpret_cast<float4*>(p) = a;
}

// D == 32*VEC (rows of 128 / 256 / 512 bytes): persistent kernel, one CTA per SM, 8 warps.  A warp owns a CONTIGUOUS range
// of the sorted list and walks it in chunks of R rows.  The rows of a chunk are fetched through the permutation by
// bulk asynchronous copies (cp.async.bulk, the TMA engine: one copy per row, issued by the lane that holds the row's
// index) into a per-warp ring of NBUF shared-memory buffers, so NBUF-1 chunks -- 8 warps x 2 x 8 KB = 128 KB per SM at
// D = 64 -- are in flight while the warp reduces the chunk that has landed, and no register holds data in flight.
// Runs that begin and end inside the warp's range are written at once; the range's first run ("head") and last run
// ("tail") go to shared memory, where warp 0 stitches the CTA's 8 ranges together in order, writes every run that lies
// strictly inside the CTA and carries the CTA's first and last runs to the next level (2 entries per CTA, <= 296 in
// all: the second level is a single CTA).
constexpr int SRV_WARPS = 8;
constexpr int SRV_NBUF = 3;

__device__ __forceinline__ uint32_t sr_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sr_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sr_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sr_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sr_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sr_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(sr_smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(sr_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void sr_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(sr_smem_u32(bar)), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();      // a protocol error traps instead of hanging the GPU
  }
}
__device__ __forceinline__ void sr_lds(float& v, uint32_t a) { asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); }
__device__ __forceinline__ void sr_lds(float2& v, uint32_t a) {
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
}
__device__ __forceinline__ void sr_lds(float4& v, uint32_t a) {
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
}

template <int VEC> struct SrvGeom {
  static constexpr int D = 32 * VEC, ROWB = D * 4;
  static constexpr int R = VEC == 4 ? 16 : 32;                  // rows per chunk
  static constexpr int BUFB = R * ROWB;                         // 4 / 8 / 8 KB
  static constexpr size_t smem = (size_t)SRV_WARPS * SRV_NBUF * BUFB + 128;
};

// one CTA's share of a level: CTA `cta` of `nctas`
template <int VEC, bool ACC>
__device__ __forceinline__ void srv_level_cta(const int32_t* __restrict__ keys, const float* __restrict__ src, int ld_src,
                                              const int32_t* __restrict__ perm, int64_t n, int64_t per_warp, int ld_dst,
                                              float* __restrict__ dst, int32_t* __restrict__ carry_keys,
                                              float* __restrict__ carry_rows, int cta, int nctas) {
  using V = typename VecT<VEC>::T;
  using G = SrvGeom<VEC>;
  constexpr int D = G::D, R = G::R;
  constexpr int CPR = G::ROWB / 16;          // 16-byte pieces per row: 8 / 16 / 32
  constexpr int RPI = 32 / CPR;              // rows one warp-wide cp.async covers: 4 / 2 / 1
  extern __shared__ uint8_t sr_dyn[];
  __shared__ V s_head[SRV_WARPS][32], s_tail[SRV_WARPS][32];
  __shared__ int32_t s_hkey[SRV_WARPS], s_tkey[SRV_WARPS];
  __shared__ int s_single[SRV_WARPS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sr_dyn) + 127) & ~(uintptr_t)127) +
                  (size_t)w * SRV_NBUF * G::BUFB;
  const uint32_t ring_s = sr_smem_u32(ring);
  const int64_t start = ((int64_t)cta * SRV_WARPS + w) * per_warp;
  const int64_t end = min(n, start + per_warp);
  if (start < end) {
    const int nchunks = (int)((end - start + R - 1) / R);
    // chunk c -> buffer c % NBUF: lane l keeps the key of entry l; the rows travel as 16-byte cp.async pieces (RPI rows
    // per warp-wide instruction), one commit group per chunk
    auto issue = [&](int c) -> int32_t {
      const int b = c % SRV_NBUF;
      const int64_t s0 = start + (int64_t)c * R;
      const int cnt = (int)min((int64_t)R, end - s0);
      const int l = min(lane, cnt - 1);      // entries past the end alias the last one (their copies are skipped)
      const int32_t k = __ldcg(keys + s0 + l);   // (L2: the second level reads what other CTAs of this launch wrote)
      const int32_t row = perm ? __ldcg(perm + s0 + l) : (int32_t)(s0 + l);
      const int piece = lane % CPR, sub = lane / CPR;
      const uint32_t dst_s = ring_s + (uint32_t)b * G::BUFB + (uint32_t)piece * 16u;
#pragma unroll
      for (int i = 0; i < R; i += RPI) {
        const int e = i + sub;
        const int32_t r = __shfl_sync(0xffffffffu, row, e & 31);
        if (e < cnt) {
          const float* g = src + (int64_t)r * ld_src + piece * 4;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_s + (uint32_t)e * G::ROWB), "l"(g) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      return k;
    };
    int32_t kq[SRV_NBUF];                    // keys of the chunks in flight (lane l: entry l)
#pragma unroll
    for (int c = 0; c < SRV_NBUF - 1; ++c) {
      if (c < nchunks) kq[c] = issue(c);
      else { kq[c] = 0; asm volatile("cp.async.commit_group;" ::: "memory"); }
    }
    V acc;
    vzero(acc);
    int32_t cur = __shfl_sync(0xffffffffu, kq[0], 0);
    const int32_t hkey = cur;
    bool first_run = true;
    float* outp = dst + lane * VEC;          // where the open run's row goes
    float* const dst_lane = dst + lane * VEC;
    const int64_t ldd = ld_dst;
    {
      V z;
      vzero(z);
      s_head[w][lane] = z;
    }
    for (int c0 = 0; c0 < nchunks; c0 += SRV_NBUF) {
#pragma unroll
      for (int u = 0; u < SRV_NBUF; ++u) {
        const int c = c0 + u;
        if (c < nchunks) {                   // warp-uniform
          // refill the buffer consumed at the previous step
          const int un = (u + SRV_NBUF - 1) % SRV_NBUF;
          if (c + SRV_NBUF - 1 < nchunks) kq[un] = issue(c + SRV_NBUF - 1);
          else asm volatile("cp.async.commit_group;" ::: "memory");
          asm volatile("cp.async.wait_group %0;" ::"n"(SRV_NBUF - 1) : "memory");
          __syncwarp();
          const int cnt = (int)min((int64_t)R, end - (start + (int64_t)c * R));
          const int32_t kmine = kq[u];
          const int32_t kprev = __shfl_up_sync(0xffffffffu, kmine, 1);
          // bit i: entry i opens a new run
          unsigned starts = __ballot_sync(0xffffffffu, lane < cnt && kmine != (lane == 0 ? cur : kprev));
          const uint32_t rows_s = ring_s + (uint32_t)u * G::BUFB + (uint32_t)lane * (VEC * 4);
#pragma unroll 8
          for (int i = 0; i < cnt; ++i) {
            V x;
            sr_lds(x, rows_s + (uint32_t)i * G::ROWB);
            if ((starts >> i) & 1u) {        // warp-uniform: the open run ended
              // the range's first run may continue the previous range's last one: it goes to shared memory
              if (first_run) { s_head[w][lane] = acc; first_run = false; }
              else vput<ACC>(outp, acc);
              vzero(acc);
              cur = __shfl_sync(0xffffffffu, kmine, i);
              outp = dst_lane + (int64_t)cur * ldd;
            }
            vadd(acc, x);
          }
          __syncwarp();                      // every lane is done with the buffer before it is refilled
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    s_tail[w][lane] = acc;                   // the open (last) run
    if (lane == 0) { s_hkey[w] = hkey; s_tkey[w] = cur; s_single[w] = first_run ? 1 : 0; }
  }
  __syncthreads();
  if (w != 0) return;
  const bool only = nctas == 1;             // the last level: nothing is carried
  const int64_t cta_first = (int64_t)cta * SRV_WARPS * per_warp;
  const int nw = (int)min((int64_t)SRV_WARPS, (n - cta_first + per_warp - 1) / per_warp);
  V cacc;
  vzero(cacc);
  int32_t ckey = 0;
  bool cvalid = false, first_emit = !only;  // the CTA's first closed run is carried, not written
  // a run that closed inside the CTA: write it, unless it is the CTA's first one
  auto emit = [&](int32_t key, const V& v) {
    if (first_emit) {
      if (lane == 0) carry_keys[2 * cta] = key;
      reinterpret_cast<V*>(carry_rows + (int64_t)(2 * cta) * D)[lane] = v;
      first_emit = false;
    } else {
      vput<ACC>(dst + (int64_t)key * ld_dst + lane * VEC, v);
    }
  };
  for (int q = 0; q < nw; ++q) {
    const int32_t hk = s_hkey[q], tk = s_tkey[q];
    const V ta = s_tail[q][lane];
    if (s_single[q]) {                      // the whole range is one run
      if (cvalid && tk == ckey) vadd(cacc, ta);
      else {
        if (cvalid) emit(ckey, cacc);
        ckey = tk; cacc = ta; cvalid = true;
      }
    } else {
      V ha = s_head[q][lane];
      if (cvalid && hk == ckey) {           // the head run continues the pending run and ends here
        vadd(cacc, ha);
        emit(ckey, cacc);
      } else {
        if (cvalid) emit(ckey, cacc);
        emit(hk, ha);
      }
      ckey = tk; cacc = ta; cvalid = true;
    }
  }
  if (only) {
    vput<ACC>(dst + (int64_t)ckey * ld_dst + lane * VEC, cacc);
  } else {
    if (first_emit) {                       // the whole CTA is one run: it travels as the head, the tail is an empty piece
      emit(ckey, cacc);
      vzero(cacc);
    }
    if (lane == 0) carry_keys[2 * cta + 1] = ckey;
    reinterpret_cast<V*>(carry_rows + (int64_t)(2 * cta + 1) * D)[lane] = cacc;
  }
}

// Level 1 on every CTA; the CTA that finishes last (a counter in global memory, zeroed by the host) then runs level 2
// -- the 2 * gridDim.x <= 296 carried pieces -- alone, in the same launch.
template <int VEC, bool ACC>
__global__ void __launch_bounds__(SRV_WARPS * 32, 1) seg_reduce_vec_kernel(
    const int32_t* __restrict__ keys, const float* __restrict__ src, int ld_src,
    const int32_t* __restrict__ perm, int64_t n, int64_t per_warp, int ld_dst, float* __restrict__ dst,
    int32_t* __restrict__ carry_keys, float* __restrict__ carry_rows, unsigned* __restrict__ done_ctas) {
  srv_level_cta<VEC, ACC>(keys, src, ld_src, perm, n, per_warp, ld_dst, dst, carry_keys, carry_rows, (int)blockIdx.x,
                          (int)gridDim.x);
  if (gridDim.x == 1) return;
  __shared__ unsigned s_ticket;
  __threadfence();                          // this CTA's carries (written by warp 0) before its ticket
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(done_ctas, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  const int64_t m = 2 * (int64_t)gridDim.x;
  const int64_t pw = ((m + SRV_WARPS - 1) / SRV_WARPS + 31) / 32 * 32;
  srv_level_cta<VEC, ACC>(carry_keys, carry_rows, 32 * VEC, nullptr, m, pw, ld_dst, dst, nullptr, nullptr, 0, 1);
}

// geometry of a vector-kernel level of m entries: entries per warp (a multiple of 32) and CTAs (<= one per SM)
static void srv_level(int64_t m, int64_t* per_warp, int* ctas) {
  int64_t pw = std::max<int64_t>(64, (m + (int64_t)kNumSMs * SRV_WARPS - 1) / ((int64_t)kNumSMs * SRV_WARPS));
  pw = (pw + 31) / 32 * 32;
  *per_warp = pw;
  *ctas = (int)((m + pw * SRV_WARPS - 1) / (pw * SRV_WARPS));
}

// any D <= 256: lanes stride across the row; one warp = one piece of SR_CH entries (first and last run carried)
template <int MAXV, bool ACC>
__global__ void __launch_bounds__(SR_WARPS * 32) seg_reduce_level_kernel(
    const int32_t* __restrict__ keys, const float* __restrict__ src, int ld_src,
    const int32_t* __restrict__ perm, int64_t n, int D, int ld_dst, float* __restrict__ dst,
    int32_t* __restrict__ carry_keys, float* __restrict__ carry_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * SR_WARPS + (threadIdx.x >> 5);
  const int64_t ntiles = (n + SR_CH - 1) / SR_CH;
  if (tile >= ntiles) return;
  const bool only = ntiles == 1;
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
  bool first_emit = !only;
  auto emit = [&](int32_t key, int slot) {    // slot >= 0: carry entry, else the row's one write
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
      const int d = lane + 32 * v;
      if (d < D) {
        if (slot >= 0) carry_rows[(tile * 2 + slot) * D + d] = acc[v];
        else if (ACC) atomicAdd(dst + (int64_t)key * ld_dst + d, acc[v]);
        else dst[(int64_t)key * ld_dst + d] = acc[v];
      }
      acc[v] = 0.f;
    }
    if (slot >= 0 && lane == 0) carry_keys[tile * 2 + slot] = key;
  };
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
          if (first_emit) { emit(cur, 0); first_emit = false; }
          else emit(cur, -1);
          cur = kk[u];
        }
#pragma unroll
        for (int v = 0; v < MAXV; ++v) acc[v] += x[u][v];
      }
    }
  }
  if (only) {
    emit(cur, -1);
  } else {
    if (first_emit) emit(cur, 0);           // one run: it travels as the head; the tail is an empty piece of the same key
    emit(cur, 1);
  }
}

size_t seg_reduce_workspace_bytes(int64_t n, int D) {
  // carries of one level: the vector kernels leave 2 per CTA (<= 2*148), the generic kernel 2 per 64 entries
  Bump b(nullptr, 0);
  b.take<unsigned>(4);
  b.take<int32_t>(2 * kNumSMs);
  b.take<float>((size_t)2 * kNumSMs * D);
  int64_t m = n;
  while (true) {
    int64_t nt = (m + SR_CH - 1) / SR_CH;
    if (nt <= 1) break;
    b.take<int32_t>(2 * nt);
    b.take<float>(2 * nt * D);
    m = 2 * nt;
  }
  return b.off + 256;
}

// accumulate != 0: dst[keys[i],:] += src[perm ? perm[i] : i, :];  accumulate == 0: dst[key,:] = the sum of its rows
// (rows of dst no key names are not touched).  Keys sorted ascending.
int seg_reduce_sorted(const int32_t* keys_sorted, const int32_t* perm, const float* src, int ld_src, int64_t n,
                      int D, float* dst, int ld_dst, void* ws, size_t ws_bytes, cudaStream_t st, int accumulate) {
  if (n <= 0) return 0;
  if (D > 256) return set_error(MTAM_ERR_INVALID, "seg_reduce: D=%d > 256", D);
  Bump b(ws, ws_bytes);
  // bulk staging needs 16-byte aligned rows; vector stores need aligned destination rows
  const bool vec_ok = (D == 32 || D == 64 || D == 128) && (ld_src % 4 == 0) && (ld_dst % (D / 32) == 0) &&
                      ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
  if (vec_ok) {   // one launch: level 1 everywhere, level 2 on the CTA that finishes last
    int64_t per_warp;
    int ctas;
    srv_level(n, &per_warp, &ctas);
    unsigned* done = b.take<unsigned>(4);
    int32_t* ck = b.take<int32_t>(2 * (size_t)ctas);
    float* cr = b.take<float>(2 * (size_t)ctas * D);
    if (!b.ok()) return set_error(MTAM_ERR_WORKSPACE, "seg_reduce: workspace %zu < %zu", ws_bytes, b.off);
    if (ctas > 1) MTAM_CUDA_CHECK(cudaMemsetAsync(done, 0, sizeof(unsigned), st));
#define SRV_LAUNCH(VEC_)                                                                                              \
  do {                                                                                                                \
    const size_t sm_ = SrvGeom<VEC_>::smem;                                                                           \
    if (accumulate) {                                                                                                 \
      MTAM_CUDA_CHECK(cudaFuncSetAttribute(seg_reduce_vec_kernel<VEC_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_)); \
      seg_reduce_vec_kernel<VEC_, true><<<ctas, SRV_WARPS * 32, sm_, st>>>(keys_sorted, src, ld_src, perm, n, per_warp, ld_dst, dst, ck, cr, done); \
    } else {                                                                                                          \
      MTAM_CUDA_CHECK(cudaFuncSetAttribute(seg_reduce_vec_kernel<VEC_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_)); \
      seg_reduce_vec_kernel<VEC_, false><<<ctas, SRV_WARPS * 32, sm_, st>>>(keys_sorted, src, ld_src, perm, n, per_warp, ld_dst, dst, ck, cr, done); \
    }                                                                                                                 \
  } while (0)
    if (D == 32) SRV_LAUNCH(1);
    else if (D == 64) SRV_LAUNCH(2);
    else SRV_LAUNCH(4);
#undef SRV_LAUNCH
    MTAM_LAUNCH_CHECK();
    return 0;
  }
  const int32_t* k = keys_sorted;
  const int32_t* pm = perm;
  const float* s = src;
  int lds = ld_src;
  int64_t m = n;
  const int maxv = (D + 31) / 32;
  while (m > 0) {
    const int64_t nt = (m + SR_CH - 1) / SR_CH;    // pieces at this level (each leaves 2 carries)
    int32_t* ck = nullptr;
    float* cr = nullptr;
    if (nt > 1) {
      ck = b.take<int32_t>(2 * nt);
      cr = b.take<float>(2 * nt * D);
      if (!b.ok()) return set_error(MTAM_ERR_WORKSPACE, "seg_reduce: workspace %zu < %zu", ws_bytes, b.off);
    }
#define SRL_LAUNCH(MAXV_)                                                                                             \
  do {                                                                                                                \
    const int blocks = cdiv(nt, SR_WARPS);                                                                            \
    if (accumulate) seg_reduce_level_kernel<MAXV_, true><<<blocks, SR_WARPS * 32, 0, st>>>(k, s, lds, pm, m, D, ld_dst, dst, ck, cr); \
    else seg_reduce_level_kernel<MAXV_, false><<<blocks, SR_WARPS * 32, 0, st>>>(k, s, lds, pm, m, D, ld_dst, dst, ck, cr);           \
  } while (0)
    if (maxv <= 1) SRL_LAUNCH(1);
    else if (maxv <= 2) SRL_LAUNCH(2);
    else if (maxv <= 4) SRL_LAUNCH(4);
    else SRL_LAUNCH(8);
#undef SRL_LAUNCH
    MTAM_LAUNCH_CHECK();
    if (nt <= 1) break;
    k = ck;
    pm = nullptr;
    s = cr;
    lds = D;
    m = 2 * nt;
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
                     cudaStream_t st, int accumulate) {
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
  MTAM_TRY(seg_reduce_sorted(ks, pm, rows, ld_src, n, D, dst, ld_dst, w + s1, s2, st, accumulate));
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
                                int32_t ld_rows, int64_t n, int32_t accumulate, void* workspace, size_t workspace_bytes,
                                int32_t* unique_idx, int32_t* n_unique, void* stream) {
  if (!dst || n < 0 || (n > 0 && (!idx || !rows || !workspace)))
    return mtam::set_error(MTAM_ERR_INVALID, "mtam_scatter_add: null/negative argument");
  if (ld_rows < D) return mtam::set_error(MTAM_ERR_INVALID, "mtam_scatter_add: ld_rows < D");
  return mtam::scatter_add_rows(dst, table_rows, D, D, idx, rows, ld_rows, n, workspace, workspace_bytes, unique_idx,
                                n_unique, (cudaStream_t)stream, accumulate);
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
                                       const float* rows, int32_t ld_rows, int64_t n, int32_t accumulate, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  if (!dst || n < 0 || (n > 0 && (!keys_sorted || !rows || !workspace)))
    return mtam::set_error(MTAM_ERR_INVALID, "mtam_scatter_add_sorted: null/negative argument");
  if (ld_rows < D) return mtam::set_error(MTAM_ERR_INVALID, "mtam_scatter_add_sorted: ld_rows < D");
  return mtam::seg_reduce_sorted(keys_sorted, perm, rows, ld_rows, n, D, dst, D, workspace, workspace_bytes,
                                 (cudaStream_t)stream, accumulate);
}
