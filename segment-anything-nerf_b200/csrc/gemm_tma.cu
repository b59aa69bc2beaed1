// TMA-fed variant of the fp32-parity tensor-core GEMM (csrc/gemm_tc.cu) for the dependent chain of the samvit head
// (nerf/network.py:36-75, 120-123):
//   forward layer    C[M,N] = act(A[M,K] . B[N,K]^T + bias)            A, B row-major (both K-major)
//   data gradient    C[M,N] = (A[M,K] . Bt[K,N]) * act'(mask[M,N])      Bt = the nn.Linear weight as stored ([out, in] = [K, N]):
//                                                                      an MN-major B operand, loaded as [32 k x 32 n] boxes in
//                                                                      the 128-byte-swizzle / 32-byte-atom layout (the only
//                                                                      MN-major form 32-bit operands have, umma.cuh)
//
// gemm_tc stages its operands LDG -> registers -> hi/lo split -> STS, two chunks ahead, and is a chain of load latencies
// (~1 us per 32-wide K chunk, tensor pipe ~10 %).  Here the global -> shared movement is the TMA's: one elected thread
// issues cp.async.bulk.tensor.2d for the [128 x 32] A box and the [64 x 32] B box of a K chunk into a 128-byte-swizzled
// stage (the canonical K-major SWIZZLE_128B UMMA layout: a row of the box = one 128-byte swizzle span), completion is a
// transaction count on the stage's mbarrier, and as many chunks as the ring has stages are in flight without occupying a
// register.  The eight worker warps only derive the 3xTF32 planes IN shared memory (hi = tf32 truncation written back in
// place, lo = x - hi into the lo plane at the same swizzled offset: no layout arithmetic), fence to the async proxy and
// arrive; a ninth warp issues the MMAs and commits the stage back; a tenth is the TMA producer.
#include <cuda.h>

#include "common.cuh"
#include "umma.cuh"

namespace sanerf {

#ifndef SANERF_GTMA_STAGES
#define SANERF_GTMA_STAGES 4
#endif
namespace gtma {
constexpr uint32_t kBM = 128, kBN = 64, kKC = 32, kWorkers = 256, kStages = SANERF_GTMA_STAGES;
constexpr uint32_t kThreads = kWorkers + 64;                        // + MMA-issuing warp + TMA-producer warp
constexpr uint32_t kATile = kBM * kKC * 4, kBTile = kBN * kKC * 4;  // 16 KB, 8 KB (one plane)
constexpr uint32_t kStageBytes = 2 * kATile + 2 * kBTile;           // A hi | A lo | B hi | B lo = 48 KB
constexpr uint32_t kSmem = kStages * kStageBytes + 1024;            // + slack to align the ring to 1024 bytes
constexpr uint32_t kTmemCols = 64;
constexpr uint32_t kLayoutSw128 = 2;                                // UMMA shared-memory descriptor: 128-byte swizzle
}  // namespace gtma

struct GemmTmaParams {
    float* C;
    const float* bias;
    const float* mask;
    uint32_t ldc, ldm, M, N, K, mask_cols;
    int act, precision, epilogue;
    float slope;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst_saddr, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t bar_saddr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        :: "r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar_saddr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_saddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void gtma_arrive(uint32_t saddr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(saddr) : "memory");
}

template <bool B_MN>
__global__ void __launch_bounds__(gtma::kThreads) gemm_tma_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                  const __grid_constant__ CUtensorMap mapB,
                                                                  const GemmTmaParams p) {
    pdl_begin();
    using namespace gtma;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_tma[kStages], s_full[kStages], s_empty[kStages], s_done;
    __shared__ uint32_t s_tmem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023u) & ~uintptr_t(1023));

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;
    const uint32_t nc = (p.K + kKC - 1u) / kKC;
    const bool split = (p.precision == 0);

    if (warp == 0) umma::tmem_alloc<kTmemCols>(umma::smem_u32(&s_tmem));
    if (tid == 32) {
        for (uint32_t s = 0; s < kStages; ++s) {
            umma::mbar_init(umma::smem_u32(&s_tma[s]), 1);                  // the producer's expect_tx arrival + the bytes
            umma::mbar_init(umma::smem_u32(&s_full[s]), kWorkers / 32);     // one arrival per worker warp
            umma::mbar_init(umma::smem_u32(&s_empty[s]), 1);                // tcgen05.commit of the chunk that used the stage
        }
        umma::mbar_init(umma::smem_u32(&s_done), 1);
        umma::fence_mbar_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s_tmem;

    if (warp == kWorkers / 32 + 1) {
        // ===================== TMA producer: one elected lane keeps the ring full ===============================
        if (umma::elect_one()) {
            for (uint32_t c = 0; c < nc; ++c) {
                const uint32_t s = c % kStages;
                if (c >= kStages) umma::mbar_wait(umma::smem_u32(&s_empty[s]), ((c / kStages) - 1u) & 1u);
                const uint32_t bar = umma::smem_u32(&s_tma[s]);
                const uint32_t base = umma::smem_u32(smem + s * kStageBytes);
                mbar_expect_tx(bar, kATile + kBTile);
                tma_load_2d(base, &mapA, c * kKC, m0, bar);                 // rows / columns past the tensor are zero-filled
                if constexpr (!B_MN) {
                    tma_load_2d(base + 2 * kATile, &mapB, c * kKC, n0, bar);
                } else {                                                    // two 32-wide blocks of n, 32 contraction rows each
                    tma_load_2d(base + 2 * kATile, &mapB, n0, c * kKC, bar);
                    tma_load_2d(base + 2 * kATile + kKC * 128u, &mapB, n0 + 32u, c * kKC, bar);
                }
            }
        }
        __syncwarp();
    } else if (warp == kWorkers / 32) {
        // ===================== MMA issuer ============================================================================
        constexpr uint32_t idesc = umma::idesc_tf32(kBM, kBN, 0, B_MN ? 1u : 0u);
        constexpr uint32_t dhi = umma::desc_hi(1024u, kLayoutSw128);         // K-major: SBO = 8 rows x 128 bytes
        // MN-major B: LBO = next 32-wide block of n (32 k rows x 128 bytes), SBO = next atom of 4 k rows (512 bytes);
        // 8 contraction rows per MMA = two atoms = 1024 bytes
        constexpr uint32_t dhi_b = B_MN ? umma::desc_hi(512u, umma::kLayoutMn32) : dhi;
        constexpr uint32_t b_lbo = B_MN ? kKC * 128u : 16u, b_step = B_MN ? (1024u >> 4) : 2u;
        for (uint32_t c = 0; c < nc; ++c) {
            const uint32_t s = c % kStages;
            umma::mbar_wait(umma::smem_u32(&s_full[s]), (c / kStages) & 1u);
            if (umma::elect_one()) {
                umma::fence_proxy_async();
                umma::fence_after_sync();
                const uint32_t sa = umma::smem_u32(smem + s * kStageBytes);
                const uint32_t dAh = umma::desc_lo(sa, 16u), dAl = umma::desc_lo(sa + kATile, 16u);
                const uint32_t dBh = umma::desc_lo(sa + 2 * kATile, b_lbo), dBl = umma::desc_lo(sa + 2 * kATile + kBTile, b_lbo);
                uint32_t acc = (c > 0u) ? 1u : 0u;
#pragma unroll
                for (uint32_t ks = 0; ks < kKC / 8u; ++ks) {
                    const uint32_t o = ks * 2u, ob = ks * b_step;             // A: 8 tf32 = 32 bytes inside the swizzle span
                    if (split) {
                        umma::mma_tf32_ss2(tmem, dAl + o, dhi, dBh + ob, dhi_b, idesc, acc);
                        umma::mma_tf32_ss2(tmem, dAh + o, dhi, dBl + ob, dhi_b, idesc, 1u);
                        acc = 1u;
                    }
                    umma::mma_tf32_ss2(tmem, dAh + o, dhi, dBh + ob, dhi_b, idesc, acc);
                    acc = 1u;
                }
                umma::commit(umma::smem_u32(&s_empty[s]));
                if (c + 1u == nc) umma::commit(umma::smem_u32(&s_done));
            }
            __syncwarp();
        }
    } else {
        // ===================== workers: hi / lo planes derived in place ==============================================
        for (uint32_t c = 0; c < nc; ++c) {
            const uint32_t s = c % kStages;
            umma::mbar_wait(umma::smem_u32(&s_tma[s]), (c / kStages) & 1u);
            uint8_t* stage = smem + s * kStageBytes;
            // A: 1024 float4, B: 512 float4; the lo plane mirrors the hi plane's (swizzled) offsets
#pragma unroll
            for (uint32_t j = 0; j < (kATile + kBTile) / 16u / kWorkers; ++j) {
                const uint32_t i = tid + kWorkers * j;
                const bool in_a = i < kATile / 16u;
                float4* hp = reinterpret_cast<float4*>(in_a ? stage : stage + 2 * kATile) + (in_a ? i : i - kATile / 16u);
                float4* lp = reinterpret_cast<float4*>(in_a ? stage + kATile : stage + 2 * kATile + kBTile) + (in_a ? i : i - kATile / 16u);
                const float4 x = *hp;
                if (split) {
                    float4 h, l;
                    umma::split_tf32(x.x, h.x, l.x); umma::split_tf32(x.y, h.y, l.y);
                    umma::split_tf32(x.z, h.z, l.z); umma::split_tf32(x.w, h.w, l.w);
                    // the hi plane needs no rewrite: the tensor core ignores the 13 low mantissa bits of a tf32 operand
                    // itself (results bit-identical to the explicit truncation, tests/test_gpu_mlp.py; 10.8 -> 9.8 us)
                    *lp = l;
                } else {
                    *hp = make_float4(rna_tf32(x.x), rna_tf32(x.y), rna_tf32(x.z), rna_tf32(x.w));
                }
            }
            // the generic -> async proxy fence for the lo planes is executed by the issuing thread after it has observed this
            // barrier (one MEMBAR per chunk instead of one per worker warp; profiles/r2b_head_backward.md)
            __syncwarp();
            if (lane == 0) gtma_arrive(umma::smem_u32(&s_full[s]));
        }
    }
    umma::mbar_wait(umma::smem_u32(&s_done), 0u);        // every MMA of this CTA has completed
    umma::fence_after_sync();

    // ---- epilogue (as gemm_tc, epilogue 0): worker warp w reads TMEM lanes 32 (w & 3) .., columns 32 (w >> 2) ..
    if (warp < kWorkers / 32) {
        const uint32_t q = warp & 3u, half = warp >> 2;
        const uint32_t m = m0 + q * 32u + lane;
        const uint32_t taddr = umma::tmem_addr(tmem, q * 32u, half * 32u);
        const bool c_vec = (p.ldc % 4u == 0u) && ((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0u);
#pragma unroll
        for (uint32_t g = 0; g < 2; ++g) {
            float v[16];
            umma::tmem_ld16(taddr + g * 16u, v);
            const uint32_t nb = n0 + half * 32u + g * 16u;
            if (m < p.M && nb < p.N) {
                if (p.epilogue == 0) {
#pragma unroll
                    for (uint32_t j = 0; j < 16; ++j) {
                        const uint32_t n = nb + j;
                        float t = v[j] + ((p.bias != nullptr && n < p.N) ? __ldg(p.bias + n) : 0.0f);
                        if (p.act) t = (t > 0.0f) ? t : t * p.slope;
                        v[j] = t;
                    }
                } else {                                // derivative of the previous layer's leaky ReLU (mask = its saved output)
                    const float* mrow = p.mask + (size_t)m * p.ldm + nb;
                    const bool m_vec = (p.ldm % 4u == 0u) && ((reinterpret_cast<uintptr_t>(p.mask) & 15u) == 0u) &&
                                       nb + 16u <= p.mask_cols && nb + 16u <= p.N;
                    if (m_vec) {
#pragma unroll
                        for (uint32_t j = 0; j < 16; j += 4) {
                            const float4 h = __ldg(reinterpret_cast<const float4*>(mrow + j));
                            v[j] = (h.x > 0.0f) ? v[j] : v[j] * p.slope;
                            v[j + 1] = (h.y > 0.0f) ? v[j + 1] : v[j + 1] * p.slope;
                            v[j + 2] = (h.z > 0.0f) ? v[j + 2] : v[j + 2] * p.slope;
                            v[j + 3] = (h.w > 0.0f) ? v[j + 3] : v[j + 3] * p.slope;
                        }
                    } else {
#pragma unroll
                        for (uint32_t j = 0; j < 16; ++j) {
                            const uint32_t n = nb + j;
                            if (n < p.mask_cols && n < p.N) v[j] = (__ldg(mrow + j) > 0.0f) ? v[j] : v[j] * p.slope;
                        }
                    }
                }
                float* dst = p.C + (size_t)m * p.ldc + nb;
                if (c_vec && nb + 16u <= p.N) {
#pragma unroll
                    for (uint32_t j = 0; j < 16; j += 4)
                        *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (uint32_t j = 0; j < 16; ++j)
                        if (nb + j < p.N) dst[j] = v[j];
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<kTmemCols>(tmem);
}

// ---- host: tensor maps through the driver entry point (no link-time dependency on libcuda) ---------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

// [rows, cols] fp32 row-major with leading dimension ld (floats): boxes of box_rows x 32 columns
static bool make_map(CUtensorMap* map, const float* base, uint32_t rows, uint32_t cols, uint32_t ld, uint32_t box_rows,
                     CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = encode_tiled();
    if (fn == nullptr) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4u};
    const cuuint32_t box[2] = {32u, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Returns SANERF_OK when the product was launched on the TMA path, -1 when this call is not eligible (caller falls back).
// b_trans: B is given as [K, N] row-major (an nn.Linear weight used for a data gradient).  epilogue 0 (bias + activation) or
// 1 (activation-derivative mask).
int launch_gemm_tma(const float* A, uint32_t lda, const float* B, uint32_t ldb, int b_trans, float* C, uint32_t ldc, uint32_t M,
                    uint32_t N, uint32_t K, int epilogue, const float* bias, int act, float slope, const float* mask, uint32_t ldm,
                    uint32_t mask_cols, int precision, cudaStream_t stream) {
    if ((lda & 3u) || (ldb & 3u) || ((uintptr_t)A & 15u) || ((uintptr_t)B & 15u) || (precision != 0 && precision != 1)) return -1;
    CUtensorMap mapA, mapB;
    if (!make_map(&mapA, A, M, K, lda, gtma::kBM, CU_TENSOR_MAP_SWIZZLE_128B)) return -1;
    if (b_trans ? !make_map(&mapB, B, K, N, ldb, gtma::kKC, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                : !make_map(&mapB, B, N, K, ldb, gtma::kBN, CU_TENSOR_MAP_SWIZZLE_128B))
        return -1;
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(gemm_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gtma::kSmem) != cudaSuccess ||
            cudaFuncSetAttribute(gemm_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gtma::kSmem) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
        configured = true;
    }
    GemmTmaParams p{C, bias, mask, ldc, ldm, M, N, K, mask_cols, act, precision, epilogue, slope};
    dim3 grid(div_up(M, gtma::kBM), div_up(N, gtma::kBN), 1);
    if (b_trans) {
        SANERF_LAUNCH(gemm_tma_kernel<true>, grid, gtma::kThreads, gtma::kSmem, stream, mapA, mapB, p);
    } else {
        SANERF_LAUNCH(gemm_tma_kernel<false>, grid, gtma::kThreads, gtma::kSmem, stream, mapA, mapB, p);
    }
    return check_launch("gemm_tma_kernel");
}

}  // namespace sanerf
