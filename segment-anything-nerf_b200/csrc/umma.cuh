// Minimal tcgen05 / TMEM / mbarrier toolkit (inline PTX, sm_100a) for the fused MLP kernels.
//
// Operand tiles live in shared memory in the UMMA "interleaved" (no-swizzle) canonical layout with 32-bit
// elements: a core matrix is 8 rows x 16 bytes stored contiguously (128 B).  We lay a [R rows x K] tile out
// chunk-major:   byte(r, k) = (k/4) * (R*16) + r*16 + (k%4)*4
// i.e. every 4-element K-chunk is a dense column of R*16 bytes.  The SAME bytes can be described
//   * K-major  (row = M/N index, contraction over k):          LBO = R*16   (next K-chunk), SBO = 128 (next 8 rows)
//   * MN-major (the 4-element chunk index = M/N, contraction over rows):  LBO = 128 (next 8 rows), SBO = R*16
// which is what lets one activation tile feed both the next layer (K-major A) and the weight-gradient GEMM
// (MN-major A/B) without a transposed copy.  Field layout of the descriptors follows the PTX ISA (and CUTLASS'
// cute/arch/mma_sm100_desc.hpp, used as a cross-check only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sanerf {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 64-bit shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}

// 64-bit descriptor of an MN-major tf32 operand.  The tensor core accepts 32-bit MN-major operands ONLY in the
// "128-byte swizzle with 32-byte atomicity" layout (layout type 1; CUTLASS: sm100_common.inl "for mn-major tf32
// operands, SW128_32B is the only available smem layout" — the no-swizzle MN-major form silently yields zeros).
//   lbo = bytes between consecutive 32-element blocks along M/N, sbo = bytes between consecutive groups of 4 K rows
__device__ __forceinline__ uint64_t smem_desc_mn32(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return smem_desc(saddr, lbo_bytes, sbo_bytes) | (static_cast<uint64_t>(1) << 61);
}

// Byte offset of element (k = contraction row, f = M/N index) in an MN-major tf32 tile of KROWS contraction rows:
// a 32-element block of f is a 128-byte line, 4 consecutive k form a 512-byte atom whose 32-byte chunks are XORed
// with (k & 3) (Swizzle<2,5,2> on the byte address: the tile base must be 512-byte aligned), atoms stack along k
// (sbo = 512) and 32-wide blocks of f follow each other every KROWS*128 bytes (lbo).
__host__ __device__ __forceinline__ uint32_t mn32_off(uint32_t krows, uint32_t k, uint32_t f) {
    return (f >> 5) * (krows * 128u) + (k >> 2) * 512u + (k & 3u) * 128u + ((((f & 31u) >> 3) ^ (k & 3u)) << 5) +
           (f & 7u) * 4u;
}

// The same descriptors as two 32-bit halves, so that the issuing thread advances an operand with ONE 32-bit add
// (low half += bytes >> 4) instead of rebuilding 64 bits per instruction:
//   low  = start address >> 4 | (LBO >> 4) << 16        high = SBO >> 4 | version 1 << 14 | layout type << 29
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
    return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout_type) {
    return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout_type << 29);
}
constexpr uint32_t kLayoutNone = 0, kLayoutMn32 = 1;

// exactly one lane of a fully converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0u;
}

// 32-bit instruction descriptor for kind::tf32, fp32 accumulate
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
    return (1u << 4) /* D = f32 */ | (2u << 7) /* A = tf32 */ | (2u << 10) /* B = tf32 */ | (a_mn_major << 15) |
           (b_mn_major << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T: A is [128 lanes x K columns] of 32-bit elements (K-major only), one thread issues
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// split-descriptor forms (see desc_lo / desc_hi)
__device__ __forceinline__ void mma_tf32_ss2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
        :: "r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts2(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}"
        :: "r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}

// all previously issued MMAs of this thread arrive on the mbarrier when complete
__device__ __forceinline__ void commit(uint32_t mbar_saddr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar_saddr) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t saddr, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(saddr), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t saddr, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(saddr), "r"(parity) : "memory");
        if (!done && spin > (1u << 20)) __trap();
    }
}

// TMEM allocation (one full warp executes these)
template <uint32_t COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_saddr) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_saddr), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(COLS) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 consecutive 32-bit columns (thread t gets lane base+t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// TMEM -> registers: 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// TMEM -> registers: 2 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float& a, float& b) {
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    a = __uint_as_float(r0);
    b = __uint_as_float(r1);
}

// registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns (no wait: call tmem_st_wait() before the
// data is consumed by an MMA or another warp)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
           "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
           "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
           "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
           "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}
// registers -> TMEM: 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
        :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
           "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
           "r"(__float_as_uint(v[7])) : "memory");
}
// registers -> TMEM: 4 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 :: "r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d))
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// TMEM address of (lane, column) relative to an allocation base: lane in bits 31..16, column in bits 15..0
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) { return base + (lane << 16) + col; }

// fp32 -> (hi, lo) with hi exactly representable in tf32 (13 low mantissa bits clear) and lo = x - hi (exact)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// byte offset of element (r, k) in a chunk-major tile with R rows
__device__ __forceinline__ uint32_t tile_off(uint32_t R, uint32_t r, uint32_t k) {
    return (k >> 2) * (R * 16u) + r * 16u + (k & 3u) * 4u;
}

}  // namespace umma
}  // namespace sanerf
