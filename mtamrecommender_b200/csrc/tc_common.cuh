// tcgen05 / TMEM / mbarrier primitives for sm_100a (inline PTX) and the shared-memory tile format
// used by every tensor-core kernel of this library.
//
// Tile format ("row128"): a tile is R rows of 128 bytes (32 fp32), 1024-byte aligned; the 16-byte
// chunk c of row r is stored at chunk position c ^ (r & 7) (the SWIZZLE_128B pattern).  The same bytes
// serve two descriptor views:
//   K-major  operand: row = M/N index, the 32 floats of a row are 32 consecutive K values;
//   MN-major operand: row = K index,   the 32 floats of a row are 32 consecutive M/N values.
// Descriptor bit layout follows cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor).
//
// Precision: kind::tf32 reads fp32 containers and uses 10 mantissa bits.  Every operand is therefore
// stored twice, hi = x with the low 13 mantissa bits cleared (exactly representable in tf32, so the
// hardware's truncation/rounding is a no-op) and lo = x - hi (exact in fp32); a product is issued as
// lo*hi + hi*lo + hi*hi with fp32 accumulation in TMEM ("3xTF32": ~2^-21 relative error per product).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mtam {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// explicit shared-memory 128-bit accesses (a pointer carved out of dynamic smem can decay to a generic LD/ST)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Bounded spin: a protocol error traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();
  }
}

// ---- TMEM ---------------------------------------------------------------------------------------
// one full warp allocates `ncols` (power of two >= 32) columns; the base address lands in *slot (smem)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp gets lane (base_lane + i), columns [c, c+16)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors --------------------------------------------------------------------------------
// shared-memory matrix descriptor, version 1 (Blackwell); layout_type: 2 = SWIZZLE_128B,
// 1 = SWIZZLE_128B_BASE32B (the only layout the hardware accepts for MN-major 32-bit operands)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version
  d |= (uint64_t)layout_type << 61;
  return d;
}
// K-major view of a row128 tile (rows = M/N index, 16-byte chunk c of row r at position c ^ (r & 7)):
// 8-row groups are 1024 B apart; LBO unused for swizzled K-major (set 16 B)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile_saddr, int kstep /*8 floats = 32 B*/) {
  return smem_desc(tile_saddr + kstep * 32, 16, 1024, 2);
}
// MN-major view of a row128 tile (rows = K index, 32-byte unit u of row r at position u ^ (r & 3), i.e.
// Swizzle<2,5,2>, atoms of 4 K-rows): MN blocks of 32 floats are `block_stride` bytes apart, the two
// 4-row atoms of one MMA (K = 8) are 512 B apart, consecutive MMAs advance by 1024 B.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_saddr, int kstep, uint32_t block_stride) {
  return smem_desc(tile_saddr + kstep * 1024, block_stride, 512, 1);
}
// The two 32-bit halves of a descriptor, so that an issue loop can advance the address field with one 32-bit add
// (the address field is the low 14 bits of `lo` in 16-byte units; offsets inside one allocation never carry out of it).
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (layout_type << 29);
}
__device__ __forceinline__ uint64_t desc_join(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
constexpr uint32_t kDescHiK = desc_hi(1024, 2);    // K-major SWIZZLE_128B
constexpr uint32_t kDescHiMN = desc_hi(512, 1);    // MN-major SWIZZLE_128B_BASE32B
// instruction descriptor, kind::tf32, fp32 accumulate (cute UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// One lane of a CONVERGED warp.  tcgen05.mma / commit take their operands from uniform registers; inside a plain
// `if (lane == 0)` the compiler cannot know that a single thread is active and wraps every such instruction in an
// elect / execute / branch-if-any-left loop with fresh R2UR moves (~50 issue cycles per MMA, more than a 128x64x8 MMA
// takes to execute).  With the issuing warp kept converged and the issuing thread chosen by elect.sync, the
// instructions are emitted straight.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n @p mov.u32 %0, 1;\n}" : "+r"(pred));
  return pred != 0;
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M lanes x K columns, one tf32 per 32-bit column, K-major only)
// is read from tensor memory -- used to feed an epilogue's result straight back into the next product.
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// registers -> TMEM: thread i of the warp writes lane (base_lane + i), columns [c, c+16).  Completion is
// awaited with tmem_st_wait() before the data is handed to another thread / the tensor core.
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 32 columns per call, no wait: issue several, then tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ---- bulk asynchronous copy (TMA engine, no tensor map): global -> shared, completion on an mbarrier ----
// the issuing thread first announces the byte count, then starts the copy; 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// plain arrive (count 1) by a thread of this CTA
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- row128 tile stores ---------------------------------------------------------------------------
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ int chunk_pos_k(int row, int chunk) { return chunk ^ (row & 7); }
__device__ __forceinline__ int chunk_pos_mn(int row, int chunk) { return ((((chunk >> 1) ^ (row & 3)) << 1) | (chunk & 1)); }
// round to nearest tf32, ties away from zero (the tensor core itself would truncate the low 13 bits): the single-term
// mode's operand.  Two integer instructions (cvt.rna.tf32.f32 runs at a fraction of the ALU rate).
__device__ __forceinline__ uint32_t tf32_rn(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
// single-term mode: only the (rounded) hi tile is written
template <bool MN>
__device__ __forceinline__ void store_chunk_rn(float* hi_tile, int row, int chunk, float4 v) {
  const uint32_t off = (uint32_t)(row * 32 + ((MN ? chunk_pos_mn(row, chunk) : chunk_pos_k(row, chunk)) << 2)) * 4u;
  sts128(smem_u32(hi_tile) + off, tf32_rn(v.x), tf32_rn(v.y), tf32_rn(v.z), tf32_rn(v.w));
}
// writes the 16-byte chunk `chunk` of row `row` of the hi and lo tiles (MN: the MN-major swizzle)
template <bool MN>
__device__ __forceinline__ void store_chunk_split(float* hi_tile, float* lo_tile, int row, int chunk, float4 v) {
  const uint32_t off = (uint32_t)(row * 32 + ((MN ? chunk_pos_mn(row, chunk) : chunk_pos_k(row, chunk)) << 2)) * 4u;
  const float hx = tf32_hi(v.x), hy = tf32_hi(v.y), hz = tf32_hi(v.z), hw = tf32_hi(v.w);
  // explicit st.shared: tiles carved out of dynamic shared memory otherwise compile to generic stores
  sts128(smem_u32(hi_tile) + off, __float_as_uint(hx), __float_as_uint(hy), __float_as_uint(hz), __float_as_uint(hw));
  sts128(smem_u32(lo_tile) + off, __float_as_uint(v.x - hx), __float_as_uint(v.y - hy), __float_as_uint(v.z - hz),
         __float_as_uint(v.w - hw));
}

}  // namespace tc
}  // namespace mtam
