// Diagnostics entry point: one tcgen05 (UMMA) tile product with operands staged exactly the way the fused MLP
// kernels stage them (csrc/umma.cuh: chunk-major interleaved shared-memory tiles, TMEM accumulators, and the
// TMEM-resident A operand).  tests/test_gpu_mlp.py runs every mode against a plain matmul on small integers, so a
// wrong descriptor field, a wrong TMEM lane mapping or a wrong major-ness shows up as an exact mismatch rather than
// as a tolerance question.  No reference equivalent (the reference has no tensor-core code).
#include "common.cuh"
#include "umma.cuh"

namespace sanerf {

// modes
//  0: D[M,N] = A[M,K] . B[N,K]^T     A, B row-major in global; both staged K-major                   (M = 128)
//  1: D[M,N] = At[K,M]^T . Bt[K,N]   At, Bt row-major in global; both staged MN-major (128-byte swizzle with
//     32-byte atomicity, umma::mn32_off)                                                            (M = 64 or 128)
//  2: as 0, A operand copied to TMEM with tcgen05.st and consumed from there                        (M = 128)
//  3: as 0 with the 3xTF32 split (hi*hi + hi*lo + lo*hi): fp32-accurate                             (M = 128)
__global__ void __launch_bounds__(128) umma_selftest_kernel(int mode, int variant, uint32_t M, uint32_t N, uint32_t K,
                                                            const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_slot;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // tiles: A (hi, lo) then B (hi, lo); rows = M / N for K-major staging, = K for MN-major staging
    const uint32_t a_rows = (mode == 1) ? K : M, a_cols = (mode == 1) ? M : K;
    const uint32_t b_rows = (mode == 1) ? K : N, b_cols = (mode == 1) ? N : K;
    // MN-major tiles are padded to whole 32-element blocks along M/N (one 128-byte line per contraction row)
    const uint32_t a_bytes = (mode == 1) ? K * 128u * div_up(M, 32u) : a_rows * a_cols * 4;
    const uint32_t b_bytes = (mode == 1) ? K * 128u * div_up(N, 32u) : b_rows * b_cols * 4;
    uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle is address-based
    uint8_t* sA = smem;
    uint8_t* sAlo = sA + a_bytes;
    uint8_t* sB = sAlo + a_bytes;
    uint8_t* sBlo = sB + b_bytes;

    if (warp == 0) umma::tmem_alloc<256>(umma::smem_u32(&tmem_base_slot));
    if (tid == 0) {
        umma::mbar_init(umma::smem_u32(&mbar), 1);
        umma::fence_mbar_init();
    }
    if (mode == 1)
        for (uint32_t i = tid; i < (2 * a_bytes + 2 * b_bytes) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.0f;
    __syncthreads();
    for (uint32_t i = tid; i < a_rows * a_cols; i += blockDim.x) {
        const uint32_t r = i / a_cols, c = i - r * a_cols;
        const uint32_t off = (mode == 1) ? umma::mn32_off(K, r, c) : umma::tile_off(a_rows, r, c);
        float hi, lo;
        umma::split_tf32(A[i], hi, lo);
        *reinterpret_cast<float*>(sA + off) = (mode == 3) ? hi : A[i];
        *reinterpret_cast<float*>(sAlo + off) = lo;
    }
    for (uint32_t i = tid; i < b_rows * b_cols; i += blockDim.x) {
        const uint32_t r = i / b_cols, c = i - r * b_cols;
        const uint32_t off = (mode == 1) ? umma::mn32_off(K, r, c) : umma::tile_off(b_rows, r, c);
        float hi, lo;
        umma::split_tf32(B[i], hi, lo);
        *reinterpret_cast<float*>(sB + off) = (mode == 3) ? hi : B[i];
        *reinterpret_cast<float*>(sBlo + off) = lo;
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_base_slot;
    const uint32_t d_col = 0, a_col = 128;

    if (mode == 2) {   // thread t owns TMEM lane t = row t of A
        for (uint32_t c0 = 0; c0 < K; c0 += 16) {
            float v[16];
#pragma unroll
            for (uint32_t j = 0; j < 16; ++j) v[j] = A[(size_t)tid * K + c0 + j];
            umma::tmem_st16(umma::tmem_addr(tmem, warp * 32, a_col + c0), v);
        }
        umma::tmem_st_wait();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
    }

    if (tid == 0) {
        const uint32_t mn = (mode == 1) ? 1u : 0u;
        const uint32_t idesc = umma::idesc_tf32(M, N, mn, mn);
        const uint32_t passes = (mode == 3) ? 3u : 1u;
        uint32_t acc = 0;
        for (uint32_t ks = 0; ks < K / 8; ++ks) {
            for (uint32_t pass = 0; pass < passes; ++pass) {
                const uint8_t* ta = (pass == 2) ? sAlo : sA;
                const uint8_t* tb = (pass == 1) ? sBlo : sB;
                uint64_t ad, bd;
                if (mode == 1) {   // MN-major: 8 contraction rows = two 512-byte swizzle atoms
                    const uint32_t lbo = K * 128u, sbo = 512u;
                    ad = umma::smem_desc_mn32(umma::smem_u32(ta) + ks * 1024u, (variant & 1) ? sbo : lbo, (variant & 1) ? lbo : sbo);
                    bd = umma::smem_desc_mn32(umma::smem_u32(tb) + ks * 1024u, (variant & 1) ? sbo : lbo, (variant & 1) ? lbo : sbo);
                } else {           // K-major: 8 contraction elements = two 16-byte chunks
                    ad = umma::smem_desc(umma::smem_u32(ta) + ks * 2u * a_rows * 16u, a_rows * 16u, 128u);
                    bd = umma::smem_desc(umma::smem_u32(tb) + ks * 2u * b_rows * 16u, b_rows * 16u, 128u);
                }
                if (mode == 2) umma::mma_tf32_ts(tmem + d_col, tmem + a_col + ks * 8u, bd, idesc, acc);
                else umma::mma_tf32(tmem + d_col, ad, bd, idesc, acc);
                acc = 1;
            }
        }
        umma::commit(umma::smem_u32(&mbar));
    }
    umma::mbar_wait(umma::smem_u32(&mbar), 0);
    umma::fence_after_sync();

    // accumulator rows: M = 128 -> row = TMEM lane; M = 64 -> row (16w + l) lives in lane 32w + l, l < 16
    const uint32_t row = (M == 128) ? tid : (warp * 16 + lane);
    const bool owns = (M == 128) || (lane < 16);
    for (uint32_t c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        umma::tmem_ld16(umma::tmem_addr(tmem, warp * 32, d_col + c0), v);
        if (owns) {
#pragma unroll
            for (uint32_t j = 0; j < 16; ++j)
                if (c0 + j < N) D[(size_t)row * N + c0 + j] = v[j];
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<256>(tmem);
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_umma_selftest(int mode, uint32_t M, uint32_t N, uint32_t K, const float* A, const float* B,
                                    float* D, void* stream) {
    SANERF_REQUIRE_PTR(A); SANERF_REQUIRE_PTR(B); SANERF_REQUIRE_PTR(D);
    const int variant = mode >> 8;     // debugging knob (descriptor field permutations); 0 is the shipped encoding
    mode &= 0xff;
    if (mode < 0 || mode > 3) return fail(SANERF_ERR_INVALID_ARG, "umma_selftest: mode 0..3");
    if (!(M == 128 || (M == 64 && mode == 1))) return fail(SANERF_ERR_INVALID_ARG, "umma_selftest: M = 128 (or 64 in mode 1)");
    if (N < 16 || N > 128 || (N % 16) != 0 || K < 8 || K > 128 || (K % 8) != 0 || (mode == 2 && (K % 16) != 0))
        return fail(SANERF_ERR_INVALID_ARG, "umma_selftest: N in 16..128 step 16, K in 8..128 step 8 (16 in mode 2)");
    const size_t smem = 1024 + ((mode == 1) ? 2 * (size_t)K * 128 * (div_up(M, 32u) + div_up(N, 32u))
                                            : 2 * (size_t)(M * K + N * K) * 4);
    if (smem > 220 * 1024) return fail(SANERF_ERR_INVALID_ARG, "umma_selftest: tiles exceed shared memory");
    cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    umma_selftest_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(mode, variant, M, N, K, A, B, D);
    return check_launch("umma_selftest_kernel");
}
