// Hand-written sm_100a primitives: tcgen05 (5th-gen tensor core) MMA with TMEM accumulators,
// TMEM allocation / loads, shared-memory matrix descriptors, mbarriers, proxy fences.
// Field layouts follow the PTX ISA "tcgen05" chapter (cross-checked against the bitfields in
// cute/arch/mma_sm100_desc.hpp); nothing here depends on CUTLASS.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ttb {
namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- canonical 128-byte-swizzled tile ------------------------------------------------------
// A matrix X[rows][cols] of 32-bit elements is stored as column blocks of 32 elements
// (128 bytes): block b, row r, 16-byte chunk c  ->  b*rows*128 + r*128 + ((c ^ (r & 7)) << 4).
// `rows` must be a multiple of 8 and the tile base 1024-byte aligned.  This one physical layout
// is simultaneously
//   * the K-major  SWIZZLE_128B operand  [MN = rows][K = cols]   (SBO = 1024 B, K-step +32 B inside
//     a block, next block every rows*128 B), and
//   * the MN-major SWIZZLE_128B operand  [K = rows][MN = cols]   (LBO = rows*128 B between 32-wide
//     MN blocks, SBO = 1024 B between 8-row K groups, K-step +1024 B),
// so the same staged bytes feed a GEMM and its transpose (used by the backward kernel).
__device__ __forceinline__ uint32_t sw128_offset(int rows, int r, int col /* element index */) {
  const int b = col >> 5;
  const int c = (col >> 2) & 7;
  return (uint32_t)(b * rows * 128 + r * 128 + ((c ^ (r & 7)) << 4) + ((col & 3) << 2));
}

// 64-bit shared-memory matrix descriptor (PTX "matrix-descriptor"): start addr >> 4 in [0,14),
// LBO >> 4 in [16,30), SBO >> 4 in [32,46), version = 1 in [46,48), layout type in [61,64)
// (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// 32-bit instruction descriptor for kind::tf32, fp32 accumulate.
//   c_format [4,6) = 1 (F32); a_format [7,10) = 2 (TF32); b_format [10,13) = 2;
//   a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major); N>>3 in [17,23); M>>4 in [24,29)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem], one CTA, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(mbar))
               : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* mbar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(mbar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a tensor-core op that never completes (bad descriptor) traps instead of
// hanging the GPU (try_wait already sleeps in hardware between polls).
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(mbar, parity); ++spins) {
    if (spins > (1u << 20)) __trap();
  }
}

// ---- TMA (bulk async copy engine), non-tensor form: global -> shared, completion on an mbarrier
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes)
               : "memory");
}
// bytes must be a multiple of 16, both addresses 16-byte aligned (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* mbar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar))
      : "memory");
}

// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// TMEM allocation: one full warp executes; the base address lands in shared memory
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
  static_assert(COLS == 32 || COLS == 64 || COLS == 128 || COLS == 256 || COLS == 512, "pow2 >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS)
               : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns starting at taddr
// (lane field of taddr must be 32 * (warp_id % 4)).  Caller must tmem_ld_wait() before use.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// same, 8 columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// round-to-nearest fp32 -> tf32 (kept in a 32-bit container); the tensor core would otherwise
// truncate the low 13 mantissa bits, which biases every product towards zero
// (same result as cvt.rna.tf32.f32 -- nearest, ties away -- for finite inputs, in 2 integer ops)
__device__ __forceinline__ float to_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ float4 to_tf32(float4 v) {
  return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
}

}  // namespace sm100
}  // namespace ttb
