// Field head, fused on the 5th-generation tensor cores (sm_100a): hash-grid encode -> Linear(32,64) -> ReLU ->
// Linear(64,64) -> ReLU -> Linear(64,16), and its backward.
//
// Replaces, for the final sampling level, GridEncoder(L16,F2) + grid_mlp of nerf/network.py:102-103, 221-227
// (`h = self.grid(x); f = self.grid_mlp(h)`), which the reference runs as one gather kernel, one permute copy and
// three cuBLAS SGEMMs with three activation round trips through HBM (185 MB per 2^18 samples).
//
// Design
//  * A tile is 128 samples = the 128 TMEM lanes.  Weights live in shared memory for the whole kernel as UMMA
//    B operands; activations go TMEM -> registers (ReLU) -> shared memory (A operand of the next layer); nothing
//    but the 16-wide head (and, for training, the 32-wide encoding) is written to HBM.
//  * fp32 parity: every product is evaluated as hi*hi + hi*lo + lo*hi with hi = the tf32 truncation of the operand
//    and lo = the exact remainder ("3xTF32"), accumulated in fp32 in TMEM: error ~2^-21 relative per product.
//    precision = 1 runs one tf32 pass on round-to-nearest operands (the 1e-2 "fp16-class" tolerance).
//  * Warp-specialised persistent CTA (one per SM): 16 producer warps gather the hash-grid features of tile i+1
//    (one thread = one sample x 4 levels = one 32-byte slice of the encoding; same corner / weight / FMA order as
//    csrc/grid_encode.cu, so the encoding is bit-identical to the reference kernel) while 4 MLP warps run the three
//    dependent MMA stages of tile i.  Two A-operand slots, full/empty mbarriers; tcgen05.commit releases a slot.
#include "grid_common.cuh"
#include "umma.cuh"

namespace sanerf {

namespace head {
constexpr uint32_t kIn = 32, kHid = 64, kOut = 16, kTile = 128;
constexpr uint32_t kLevels = 16, kLevelsPerThread = 4;
constexpr uint32_t kMlpWarps = 4, kProducerWarps = 16;
constexpr uint32_t kThreads = (kMlpWarps + kProducerWarps) * 32;
// shared-memory carve-up (bytes); every operand tile is chunk-major (umma.cuh), hi plane then lo plane
constexpr uint32_t kW1 = kHid * kIn * 4, kW2 = kHid * kHid * 4, kW3 = kOut * kHid * 4;
constexpr uint32_t kA1 = kTile * kIn * 4, kH = kTile * kHid * 4;
constexpr uint32_t oW1 = 0, oW2 = oW1 + 2 * kW1, oW3 = oW2 + 2 * kW2;
constexpr uint32_t oA1 = oW3 + 2 * kW3;            // 2 slots x (hi, lo)
constexpr uint32_t oH = oA1 + 4 * kA1;             // (hi, lo)
constexpr uint32_t kFwdSmem = oH + 2 * kH;
// TMEM columns
constexpr uint32_t cD1 = 0, cD2 = 64, cD3 = 128, kTmemCols = 256;
}  // namespace head

struct HeadFwdParams {
    const float* x01;        // [B,3] in [0,1]^3, or NULL: take the encoding from `enc_in`
    const float* table;      // [rows,2] fp32
    const int32_t* offsets;  // [17]
    const float* enc_in;     // [B,32] when x01 == NULL
    const float* w1;         // [64,32]  nn.Linear layout [out,in]
    const float* w2;         // [64,64]
    const float* w3;         // [16,64]
    float* enc_out;          // [B,32] or NULL
    float* out;              // [B,16]
    uint32_t B, H;
    float S;
    int precision;           // 0: 3xTF32 (fp32 parity), 1: single tf32 pass
};

__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// write 4 consecutive K elements (one 16-byte chunk) of row r into the hi / lo planes of a chunk-major tile
__device__ __forceinline__ void put_chunk(uint8_t* hi_plane, uint8_t* lo_plane, uint32_t rows, uint32_t r, uint32_t chunk,
                                          float a, float b, float c, float d, bool split) {
    const uint32_t off = chunk * (rows * 16u) + r * 16u;
    if (split) {
        float h0, h1, h2, h3, l0, l1, l2, l3;
        umma::split_tf32(a, h0, l0); umma::split_tf32(b, h1, l1); umma::split_tf32(c, h2, l2); umma::split_tf32(d, h3, l3);
        *reinterpret_cast<float4*>(hi_plane + off) = make_float4(h0, h1, h2, h3);
        *reinterpret_cast<float4*>(lo_plane + off) = make_float4(l0, l1, l2, l3);
    } else {
        *reinterpret_cast<float4*>(hi_plane + off) = make_float4(round_tf32(a), round_tf32(b), round_tf32(c), round_tf32(d));
    }
}

// stage an nn.Linear weight [rows=out, cols=in] as a chunk-major B operand (hi, lo planes)
__device__ __forceinline__ void stage_weight(const float* __restrict__ w, uint8_t* plane_hi, uint32_t rows, uint32_t cols,
                                             bool split, uint32_t tid, uint32_t nthreads) {
    uint8_t* plane_lo = plane_hi + rows * cols * 4;
    for (uint32_t i = tid; i < rows * cols / 4; i += nthreads) {
        const uint32_t r = i / (cols / 4), ch = i - r * (cols / 4);
        const float4 v = __ldg(reinterpret_cast<const float4*>(w) + i);
        put_chunk(plane_hi, plane_lo, rows, r, ch, v.x, v.y, v.z, v.w, split);
    }
}

// D[128, N] = A[128, K] . B[N, K]^T over chunk-major K-major operands; one thread issues
template <uint32_t N, uint32_t K, uint32_t B_ROWS>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                           bool split) {
    constexpr uint32_t idesc = umma::idesc_tf32(128, N, 0, 0);
    constexpr uint32_t a_lbo = head::kTile * 16u, b_lbo = B_ROWS * 16u;
    uint32_t acc = 0;
#pragma unroll
    for (uint32_t ks = 0; ks < K / 8; ++ks) {
        const uint32_t ao = ks * 2u * a_lbo, bo = ks * 2u * b_lbo;
        if (split) {
            umma::mma_tf32(tmem_d, umma::smem_desc(a_lo + ao, a_lbo, 128u), umma::smem_desc(b_hi + bo, b_lbo, 128u), idesc, acc);
            umma::mma_tf32(tmem_d, umma::smem_desc(a_hi + ao, a_lbo, 128u), umma::smem_desc(b_lo + bo, b_lbo, 128u), idesc, 1u);
            acc = 1;
        }
        umma::mma_tf32(tmem_d, umma::smem_desc(a_hi + ao, a_lbo, 128u), umma::smem_desc(b_hi + bo, b_lbo, 128u), idesc, acc);
        acc = 1;
    }
}

__device__ __forceinline__ void mbar_arrive(uint32_t saddr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(saddr) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

// TMEM accumulator (64 columns of this thread's lane) -> ReLU -> hi/lo planes of the chunk-major H tile
__device__ __forceinline__ void relu_to_smem(uint32_t tmem_lane_col, uint8_t* h_hi, uint8_t* h_lo, uint32_t row, bool split) {
#pragma unroll
    for (uint32_t c0 = 0; c0 < head::kHid; c0 += 16) {
        float v[16];
        umma::tmem_ld16(tmem_lane_col + c0, v);
#pragma unroll
        for (uint32_t j = 0; j < 16; j += 4)
            put_chunk(h_hi, h_lo, head::kTile, row, (c0 + j) >> 2, fmaxf(v[j], 0.0f), fmaxf(v[j + 1], 0.0f),
                      fmaxf(v[j + 2], 0.0f), fmaxf(v[j + 3], 0.0f), split);
    }
}

__global__ void __launch_bounds__(head::kThreads, 1) head_forward_kernel(const HeadFwdParams p, const uint32_t tiles) {
    using namespace head;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_full[2], s_empty[2], s_mma;
    __shared__ uint32_t s_tmem;
    __shared__ LevelGeom<3> s_geo[kLevels];
    __shared__ uint32_t s_base[kLevels];

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool split = (p.precision == 0);

    if (warp == 0) umma::tmem_alloc<kTmemCols>(umma::smem_u32(&s_tmem));
    if (tid == 32) {
        for (int s = 0; s < 2; ++s) {
            umma::mbar_init(umma::smem_u32(&s_full[s]), kProducerWarps);
            umma::mbar_init(umma::smem_u32(&s_empty[s]), 1);
        }
        umma::mbar_init(umma::smem_u32(&s_mma), 1);
        umma::fence_mbar_init();
    }
    if (p.x01 != nullptr && tid >= 64 && tid < 64 + kLevels) {
        const uint32_t l = tid - 64;
        s_geo[l] = level_geometry<3>(p.offsets, l, p.S, p.H, 0u);
        s_base[l] = (uint32_t)__ldg(p.offsets + l);
    }
    stage_weight(p.w1, smem + oW1, kHid, kIn, split, tid, kThreads);
    stage_weight(p.w2, smem + oW2, kHid, kHid, split, tid, kThreads);
    stage_weight(p.w3, smem + oW3, kOut, kHid, split, tid, kThreads);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s_tmem;

    if (warp >= kMlpWarps) {
        // ===================== producers: encoding of one sample x 4 levels -> A-operand slot =====================
        const uint32_t g = tid - kMlpWarps * 32;
        const uint32_t r = g & (kTile - 1), lg = g >> 7;
        const float* __restrict__ table = p.table;
        for (uint32_t it = 0, tile = blockIdx.x; tile < tiles; ++it, tile += gridDim.x) {
            const uint32_t slot = it & 1u;
            umma::mbar_wait(umma::smem_u32(&s_empty[slot]), ((it >> 1) & 1u) ^ 1u);
            const uint32_t b = tile * kTile + r;
            const bool live = b < p.B;
            float enc[2 * kLevelsPerThread];
            if (p.x01 != nullptr) {
                float x[3] = {0.5f, 0.5f, 0.5f};
                if (live) {
#pragma unroll
                    for (uint32_t d = 0; d < 3; ++d) x[d] = __ldg(p.x01 + (size_t)b * 3 + d);
                }
                const bool zero = !live || out_of_range<3>(x);
                float2 val[kLevelsPerThread][8];
                float frac[kLevelsPerThread][3];
#pragma unroll
                for (uint32_t q = 0; q < kLevelsPerThread; ++q) {
                    const uint32_t level = lg * kLevelsPerThread + q;
                    const LevelGeom<3> geo = s_geo[level];
                    const Cell<3> cell = locate<3>(geo, x, false, 0u);
                    const size_t base = (size_t)s_base[level];
#pragma unroll
                    for (uint32_t d = 0; d < 3; ++d) frac[q][d] = cell.f[d];
#pragma unroll
                    for (uint32_t k = 0; k < 8; ++k)
                        val[q][k] = zero ? make_float2(0.0f, 0.0f)
                                         : __ldg(reinterpret_cast<const float2*>(table + (base + corner_row<3>(geo, cell, k)) * 2));
                }
#pragma unroll
                for (uint32_t q = 0; q < kLevelsPerThread; ++q) {
                    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
                    for (uint32_t k = 0; k < 8; ++k) {
                        float w = 1.0f;
#pragma unroll
                        for (uint32_t d = 0; d < 3; ++d) w *= (k & (1u << d)) ? frac[q][d] : (1.0f - frac[q][d]);
                        a0 = __fmaf_rn(w, val[q][k].x, a0);
                        a1 = __fmaf_rn(w, val[q][k].y, a1);
                    }
                    enc[2 * q] = a0;
                    enc[2 * q + 1] = a1;
                }
                if (p.enc_out != nullptr && live) {
                    float4* dst = reinterpret_cast<float4*>(p.enc_out + (size_t)b * kIn + lg * 8);
                    dst[0] = make_float4(enc[0], enc[1], enc[2], enc[3]);
                    dst[1] = make_float4(enc[4], enc[5], enc[6], enc[7]);
                }
            } else {
                float4 u = make_float4(0.f, 0.f, 0.f, 0.f), v = u;
                if (live) {
                    const float4* src = reinterpret_cast<const float4*>(p.enc_in + (size_t)b * kIn + lg * 8);
                    u = __ldg(src);
                    v = __ldg(src + 1);
                }
                enc[0] = u.x; enc[1] = u.y; enc[2] = u.z; enc[3] = u.w;
                enc[4] = v.x; enc[5] = v.y; enc[6] = v.z; enc[7] = v.w;
            }
            uint8_t* a_hi = smem + oA1 + slot * 2 * kA1;
            put_chunk(a_hi, a_hi + kA1, kTile, r, lg * 2, enc[0], enc[1], enc[2], enc[3], split);
            put_chunk(a_hi, a_hi + kA1, kTile, r, lg * 2 + 1, enc[4], enc[5], enc[6], enc[7], split);
            umma::fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(umma::smem_u32(&s_full[slot]));
        }
    } else {
        // ===================== MLP warps: thread t <-> TMEM lane t <-> sample row t of the tile =====================
        const uint32_t lane_base = umma::tmem_addr(tmem, warp * 32, 0);
        const uint32_t sW1 = umma::smem_u32(smem + oW1), sW2 = umma::smem_u32(smem + oW2), sW3 = umma::smem_u32(smem + oW3);
        const uint32_t sH = umma::smem_u32(smem + oH);
        uint8_t* h_hi = smem + oH;
        uint8_t* h_lo = h_hi + kH;
        uint32_t mma_phase = 0;
        for (uint32_t it = 0, tile = blockIdx.x; tile < tiles; ++it, tile += gridDim.x) {
            const uint32_t slot = it & 1u;
            if (tid == 0) {
                umma::mbar_wait(umma::smem_u32(&s_full[slot]), (it >> 1) & 1u);
                umma::fence_after_sync();
                const uint32_t sA = umma::smem_u32(smem + oA1 + slot * 2 * kA1);
                issue_gemm<kHid, kIn, kHid>(tmem + cD1, sA, sA + kA1, sW1, sW1 + kW1, split);
                umma::commit(umma::smem_u32(&s_empty[slot]));     // A slot free once layer 1 has consumed it
                umma::commit(umma::smem_u32(&s_mma));
            }
            umma::mbar_wait(umma::smem_u32(&s_mma), mma_phase); mma_phase ^= 1u;
            umma::fence_after_sync();
            relu_to_smem(lane_base + cD1, h_hi, h_lo, tid, split);
            umma::fence_proxy_async();
            umma::fence_before_sync();
            named_bar_sync(1, kMlpWarps * 32);
            if (tid == 0) {
                umma::fence_after_sync();
                issue_gemm<kHid, kHid, kHid>(tmem + cD2, sH, sH + kH, sW2, sW2 + kW2, split);
                umma::commit(umma::smem_u32(&s_mma));
            }
            umma::mbar_wait(umma::smem_u32(&s_mma), mma_phase); mma_phase ^= 1u;
            umma::fence_after_sync();
            relu_to_smem(lane_base + cD2, h_hi, h_lo, tid, split);   // layer 2 has completed: H may be overwritten
            umma::fence_proxy_async();
            umma::fence_before_sync();
            named_bar_sync(1, kMlpWarps * 32);
            if (tid == 0) {
                umma::fence_after_sync();
                issue_gemm<kOut, kHid, kOut>(tmem + cD3, sH, sH + kH, sW3, sW3 + kW3, split);
                umma::commit(umma::smem_u32(&s_mma));
            }
            umma::mbar_wait(umma::smem_u32(&s_mma), mma_phase); mma_phase ^= 1u;
            umma::fence_after_sync();
            float v[16];
            umma::tmem_ld16(lane_base + cD3, v);
            const uint32_t b = tile * kTile + tid;
            if (b < p.B) {
                float4* dst = reinterpret_cast<float4*>(p.out + (size_t)b * kOut);
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
            umma::fence_before_sync();   // orders this tile's TMEM reads before the next tile's MMAs (issued after bar 1)
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<head::kTmemCols>(tmem);
}


// ================================================= backward =================================================
// Per 128-sample tile (all operands chunk-major in shared memory, hi/lo planes; X and Y are 128x64 buffers):
//   0. enc -> Y (as A1), g_out -> G3
//   1. D1 = A1 W1^T                  H1 = relu(D1) -> X          (mask1 kept in registers)
//   2. D2 = H1 W2^T                  H2 = relu(D2) -> Y          (mask2)
//   3. dW3^T += H2^T G3 ; DG2 = G3 W3        G2 = DG2 * mask2 -> Y
//   4. dW2   += G2^T H1 ; DG1 = G2 W2        G1 = DG1 * mask1 -> X ;  enc -> Y (as A1 again)
//   5. dW1   += G1^T A1 ; DGE = G1 W1        g_enc = DGE -> HBM
// The three weight gradients accumulate in TMEM (M = 64 accumulators) over all tiles of the persistent CTA and are
// added to HBM once per CTA.  The same shared-memory image of a tile serves as K-major operand (rows = samples) in
// the data-gradient products and as MN-major operand (contraction over the 128 samples) in the weight-gradient
// products; the same image of a weight serves forward (K-major B) and backward (MN-major B).
namespace head {
constexpr uint32_t kG3 = kTile * kOut * 4;
constexpr uint32_t oX = oW3 + 2 * kW3, oY = oX + 2 * kH, oG3 = oY + 2 * kH;
constexpr uint32_t kBwdSmem = oG3 + 2 * kG3;
constexpr uint32_t cDG2 = 128, cDG1 = 192, cDGE = 256, cW1 = 288, cW2 = 320, cW3 = 384, kBwdTmemCols = 512;
constexpr uint32_t kBwdThreads = 128;
}  // namespace head

struct HeadBwdParams {
    const float* enc;     // [B,32]
    const float* g_out;   // [B,16]
    const float* w1;
    const float* w2;
    const float* w3;
    float* g_enc;         // [B,32]
    float* g_w1;          // [64,32] accumulated
    float* g_w2;          // [64,64]
    float* g_w3;          // [16,64]
    uint32_t B;
    int precision;
};

// D[M, N] (+)= At^T . Bt : both operands MN-major views of chunk-major tiles whose ROWS are the contraction index
template <uint32_t M, uint32_t N, uint32_t KROWS>
__device__ __forceinline__ void issue_gemm_mn(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t a_sbo, uint32_t b_hi,
                                              uint32_t b_lo, uint32_t b_sbo, bool split, uint32_t acc) {
    constexpr uint32_t idesc = umma::idesc_tf32(M, N, 1, 1);
#pragma unroll
    for (uint32_t ks = 0; ks < KROWS / 8; ++ks) {
        const uint32_t o = ks * 128u;
        if (split) {
            umma::mma_tf32(tmem_d, umma::smem_desc(a_lo + o, 128u, a_sbo), umma::smem_desc(b_hi + o, 128u, b_sbo), idesc, acc);
            umma::mma_tf32(tmem_d, umma::smem_desc(a_hi + o, 128u, a_sbo), umma::smem_desc(b_lo + o, 128u, b_sbo), idesc, 1u);
            acc = 1;
        }
        umma::mma_tf32(tmem_d, umma::smem_desc(a_hi + o, 128u, a_sbo), umma::smem_desc(b_hi + o, 128u, b_sbo), idesc, acc);
        acc = 1;
    }
}

// D[128, N] = A[128, K] . W[K, N] : A K-major (rows = samples), W = chunk-major weight tile with W_ROWS = K rows,
// read as an MN-major B operand
template <uint32_t N, uint32_t K>
__device__ __forceinline__ void issue_gemm_kn(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, uint32_t w_lo,
                                              bool split) {
    constexpr uint32_t idesc = umma::idesc_tf32(128, N, 0, 1);
    constexpr uint32_t a_lbo = head::kTile * 16u, w_sbo = K * 16u;
    uint32_t acc = 0;
#pragma unroll
    for (uint32_t ks = 0; ks < K / 8; ++ks) {
        const uint32_t ao = ks * 2u * a_lbo, wo = ks * 128u;
        if (split) {
            umma::mma_tf32(tmem_d, umma::smem_desc(a_lo + ao, a_lbo, 128u), umma::smem_desc(w_hi + wo, 128u, w_sbo), idesc, acc);
            umma::mma_tf32(tmem_d, umma::smem_desc(a_hi + ao, a_lbo, 128u), umma::smem_desc(w_lo + wo, 128u, w_sbo), idesc, 1u);
            acc = 1;
        }
        umma::mma_tf32(tmem_d, umma::smem_desc(a_hi + ao, a_lbo, 128u), umma::smem_desc(w_hi + wo, 128u, w_sbo), idesc, acc);
        acc = 1;
    }
}

// accumulator -> ReLU -> tile, remembering the sign pattern
__device__ __forceinline__ uint64_t relu_to_smem_mask(uint32_t tmem_lane_col, uint8_t* h_hi, uint8_t* h_lo, uint32_t row,
                                                      bool split) {
    uint64_t mask = 0;
#pragma unroll
    for (uint32_t c0 = 0; c0 < head::kHid; c0 += 16) {
        float v[16];
        umma::tmem_ld16(tmem_lane_col + c0, v);
#pragma unroll
        for (uint32_t j = 0; j < 16; ++j) {
            mask |= (v[j] > 0.0f) ? (1ull << (c0 + j)) : 0ull;
            v[j] = fmaxf(v[j], 0.0f);
        }
#pragma unroll
        for (uint32_t j = 0; j < 16; j += 4)
            put_chunk(h_hi, h_lo, head::kTile, row, (c0 + j) >> 2, v[j], v[j + 1], v[j + 2], v[j + 3], split);
    }
    return mask;
}

// accumulator * mask -> tile
__device__ __forceinline__ void masked_to_smem(uint32_t tmem_lane_col, uint64_t mask, uint8_t* g_hi, uint8_t* g_lo,
                                               uint32_t row, bool split) {
#pragma unroll
    for (uint32_t c0 = 0; c0 < head::kHid; c0 += 16) {
        float v[16];
        umma::tmem_ld16(tmem_lane_col + c0, v);
#pragma unroll
        for (uint32_t j = 0; j < 16; ++j) v[j] = ((mask >> (c0 + j)) & 1ull) ? v[j] : 0.0f;
#pragma unroll
        for (uint32_t j = 0; j < 16; j += 4)
            put_chunk(g_hi, g_lo, head::kTile, row, (c0 + j) >> 2, v[j], v[j + 1], v[j + 2], v[j + 3], split);
    }
}

__global__ void __launch_bounds__(head::kBwdThreads, 1) head_backward_kernel(const HeadBwdParams p, const uint32_t tiles) {
    using namespace head;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_mma;
    __shared__ uint32_t s_tmem;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool split = (p.precision == 0);

    if (warp == 0) umma::tmem_alloc<kBwdTmemCols>(umma::smem_u32(&s_tmem));
    if (tid == 32) {
        umma::mbar_init(umma::smem_u32(&s_mma), 1);
        umma::fence_mbar_init();
    }
    stage_weight(p.w1, smem + oW1, kHid, kIn, split, tid, kBwdThreads);
    stage_weight(p.w2, smem + oW2, kHid, kHid, split, tid, kBwdThreads);
    stage_weight(p.w3, smem + oW3, kOut, kHid, split, tid, kBwdThreads);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t lane_base = umma::tmem_addr(tmem, warp * 32, 0);
    const uint32_t sW1 = umma::smem_u32(smem + oW1), sW2 = umma::smem_u32(smem + oW2), sW3 = umma::smem_u32(smem + oW3);
    const uint32_t sX = umma::smem_u32(smem + oX), sY = umma::smem_u32(smem + oY), sG3 = umma::smem_u32(smem + oG3);
    uint8_t* X = smem + oX;
    uint8_t* Y = smem + oY;
    uint8_t* G3 = smem + oG3;
    constexpr uint32_t kChunkStride = kTile * 16u;       // bytes between 4-column chunks of a 128-row tile
    uint32_t phase = 0;

    // this thread's rows of the current tile, prefetched one tile ahead
    float4 enc[8], nenc[8], ng[4];
    auto fetch = [&](uint32_t tile, float4 (&e)[8], float4 (&g)[4]) {
        const uint32_t b = tile * kTile + tid;
        if (tile < tiles && b < p.B) {
            const float4* pe = reinterpret_cast<const float4*>(p.enc + (size_t)b * kIn);
            const float4* pg = reinterpret_cast<const float4*>(p.g_out + (size_t)b * kOut);
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) e[j] = __ldg(pe + j);
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) g[j] = __ldg(pg + j);
        } else {
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) e[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto put_enc = [&](const float4 (&e)[8]) {        // A1 image in Y: hi plane at Y, lo plane at Y + kA1
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) put_chunk(Y, Y + kA1, kTile, tid, j, e[j].x, e[j].y, e[j].z, e[j].w, split);
    };
    auto mma_done = [&]() {
        umma::mbar_wait(umma::smem_u32(&s_mma), phase);
        phase ^= 1u;
        umma::fence_after_sync();
    };
    auto publish = [&]() {                             // smem tiles written -> visible to the tensor core; TMEM reads done
        umma::fence_proxy_async();
        umma::fence_before_sync();
        __syncthreads();
    };

    fetch(blockIdx.x, nenc, ng);
    for (uint32_t it = 0, tile = blockIdx.x; tile < tiles; ++it, tile += gridDim.x) {
        const uint32_t first = (it == 0) ? 0u : 1u;
        // ---- 0: stage this tile's rows, prefetch the next tile's
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) enc[j] = nenc[j];
        put_enc(enc);
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) put_chunk(G3, G3 + kG3, kTile, tid, j, ng[j].x, ng[j].y, ng[j].z, ng[j].w, split);
        fetch(tile + gridDim.x, nenc, ng);
        publish();
        // ---- 1
        if (tid == 0) {
            umma::fence_after_sync();
            issue_gemm<kHid, kIn, kHid>(tmem + cD1, sY, sY + kA1, sW1, sW1 + kW1, split);
            umma::commit(umma::smem_u32(&s_mma));
        }
        mma_done();
        const uint64_t mask1 = relu_to_smem_mask(lane_base + cD1, X, X + kH, tid, split);
        publish();
        // ---- 2
        if (tid == 0) {
            umma::fence_after_sync();
            issue_gemm<kHid, kHid, kHid>(tmem + cD2, sX, sX + kH, sW2, sW2 + kW2, split);
            umma::commit(umma::smem_u32(&s_mma));
        }
        mma_done();
        const uint64_t mask2 = relu_to_smem_mask(lane_base + cD2, Y, Y + kH, tid, split);
        publish();
        // ---- 3
        if (tid == 0) {
            umma::fence_after_sync();
            issue_gemm_mn<64, kOut, kTile>(tmem + cW3, sY, sY + kH, kChunkStride, sG3, sG3 + kG3, kChunkStride, split, first);
            issue_gemm_kn<kHid, kOut>(tmem + cDG2, sG3, sG3 + kG3, sW3, sW3 + kW3, split);
            umma::commit(umma::smem_u32(&s_mma));
        }
        mma_done();
        masked_to_smem(lane_base + cDG2, mask2, Y, Y + kH, tid, split);
        publish();
        // ---- 4
        if (tid == 0) {
            umma::fence_after_sync();
            issue_gemm_mn<64, kHid, kTile>(tmem + cW2, sY, sY + kH, kChunkStride, sX, sX + kH, kChunkStride, split, first);
            issue_gemm_kn<kHid, kHid>(tmem + cDG1, sY, sY + kH, sW2, sW2 + kW2, split);
            umma::commit(umma::smem_u32(&s_mma));
        }
        mma_done();
        masked_to_smem(lane_base + cDG1, mask1, X, X + kH, tid, split);
        put_enc(enc);
        publish();
        // ---- 5
        if (tid == 0) {
            umma::fence_after_sync();
            issue_gemm_mn<64, kIn, kTile>(tmem + cW1, sX, sX + kH, kChunkStride, sY, sY + kA1, kChunkStride, split, first);
            issue_gemm_kn<kIn, kHid>(tmem + cDGE, sX, sX + kH, sW1, sW1 + kW1, split);
            umma::commit(umma::smem_u32(&s_mma));
        }
        mma_done();
        {
            const uint32_t b = tile * kTile + tid;
#pragma unroll
            for (uint32_t c0 = 0; c0 < kIn; c0 += 16) {
                float v[16];
                umma::tmem_ld16(lane_base + cDGE + c0, v);
                if (b < p.B) {
                    float4* dst = reinterpret_cast<float4*>(p.g_enc + (size_t)b * kIn + c0);
#pragma unroll
                    for (uint32_t j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            }
        }
        umma::fence_before_sync();
        __syncthreads();        // every read of X / Y / G3 by step-5 MMAs is complete (mma_done) before the next tile stages
    }

    // ---- weight gradients: M = 64 accumulators, row (16 w + l) lives in TMEM lane (32 w + l), l < 16
    const uint32_t row = warp * 16 + lane;
    const bool owns = lane < 16;
#pragma unroll
    for (uint32_t c0 = 0; c0 < kIn; c0 += 16) {          // dW1[out=row][in=c]
        float v[16];
        umma::tmem_ld16(lane_base + cW1 + c0, v);
        if (owns) {
#pragma unroll
            for (uint32_t j = 0; j < 16; j += 4) red_add_v4_f32(p.g_w1 + row * kIn + c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    }
#pragma unroll
    for (uint32_t c0 = 0; c0 < kHid; c0 += 16) {         // dW2[out=row][in=c]
        float v[16];
        umma::tmem_ld16(lane_base + cW2 + c0, v);
        if (owns) {
#pragma unroll
            for (uint32_t j = 0; j < 16; j += 4) red_add_v4_f32(p.g_w2 + row * kHid + c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    }
    {                                                    // dW3^T[in=row][out=c] -> g_w3[out][in]
        float v[16];
        umma::tmem_ld16(lane_base + cW3, v);
        if (owns) {
#pragma unroll
            for (uint32_t j = 0; j < 16; ++j) red_add_f32(p.g_w3 + j * kHid + row, v[j]);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<head::kBwdTmemCols>(tmem);
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_field_head_forward(const float* x01, const float* table, const int32_t* offsets, float S, uint32_t H,
                                         const float* enc_in, const float* w1, const float* w2, const float* w3, uint32_t B,
                                         float* enc_out, float* out, int precision, void* stream) {
    if (B == 0) return SANERF_OK;
    if (x01 != nullptr) {
        SANERF_REQUIRE_PTR(table); SANERF_REQUIRE_PTR(offsets);
    } else {
        SANERF_REQUIRE_PTR(enc_in);
    }
    SANERF_REQUIRE_PTR(w1); SANERF_REQUIRE_PTR(w2); SANERF_REQUIRE_PTR(w3); SANERF_REQUIRE_PTR(out);
    if (precision != 0 && precision != 1) return fail(SANERF_ERR_INVALID_ARG, "field_head: precision 0 (3xTF32) or 1 (TF32)");
    HeadFwdParams p{x01, table, offsets, enc_in, w1, w2, w3, enc_out, out, B, H, S, precision};
    const uint32_t tiles = div_up(B, head::kTile);
    const uint32_t blocks = tiles < (uint32_t)kNumSMs ? tiles : (uint32_t)kNumSMs;
    cudaError_t e = cudaFuncSetAttribute(head_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, head::kFwdSmem);
    if (e != cudaSuccess) return fail(SANERF_ERR_CUDA, "field_head_forward: %s", cudaGetErrorString(e));
    head_forward_kernel<<<blocks, head::kThreads, head::kFwdSmem, static_cast<cudaStream_t>(stream)>>>(p, tiles);
    return check_launch("head_forward_kernel");
}

extern "C" int sanerf_field_head_backward(const float* enc, const float* g_out, const float* w1, const float* w2,
                                          const float* w3, uint32_t B, float* g_enc, float* g_w1, float* g_w2, float* g_w3,
                                          int precision, void* stream) {
    if (B == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(enc); SANERF_REQUIRE_PTR(g_out); SANERF_REQUIRE_PTR(w1); SANERF_REQUIRE_PTR(w2); SANERF_REQUIRE_PTR(w3);
    SANERF_REQUIRE_PTR(g_enc); SANERF_REQUIRE_PTR(g_w1); SANERF_REQUIRE_PTR(g_w2); SANERF_REQUIRE_PTR(g_w3);
    if (precision != 0 && precision != 1) return fail(SANERF_ERR_INVALID_ARG, "field_head: precision 0 (3xTF32) or 1 (TF32)");
    HeadBwdParams p{enc, g_out, w1, w2, w3, g_enc, g_w1, g_w2, g_w3, B, precision};
    const uint32_t tiles = div_up(B, head::kTile);
    const uint32_t blocks = tiles < (uint32_t)kNumSMs ? tiles : (uint32_t)kNumSMs;
    cudaError_t e = cudaFuncSetAttribute(head_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, head::kBwdSmem);
    if (e != cudaSuccess) return fail(SANERF_ERR_CUDA, "field_head_backward: %s", cudaGetErrorString(e));
    head_backward_kernel<<<blocks, head::kBwdThreads, head::kBwdSmem, static_cast<cudaStream_t>(stream)>>>(p, tiles);
    return check_launch("head_backward_kernel");
}
