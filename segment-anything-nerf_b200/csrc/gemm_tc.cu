// fp32-parity GEMM on the 5th-generation tensor cores (sm_100a) for the wide MLP head of the SAM feature field:
//   samvit_mlp = SkipConnMLP(163 -> 256 x4 -> 256, skip at layer 2, bias, leaky ReLU) (nerf/network.py:36-75, 120-123).
// The reference runs it as nn.Linear -> cuBLAS SIMT SGEMMs + separate bias / activation / concat kernels, and autograd
// adds one data-gradient GEMM, one weight-gradient GEMM, one column-sum and one activation-backward kernel per layer
// (55 launches, 650 us per 4096-ray step on a B200).  Here every product of the forward AND the backward is the same
// kernel with its elementwise neighbours folded into the operand load / the epilogue:
//
//   C[M,N] (op)= A . B^T          A: [M,K] row-major, or given transposed as [K,M]   (a_trans)
//                                 B: [N,K] row-major, or given transposed as [K,N]   (b_trans)
//   epilogue 0:  C  = act(acc + bias[n])                       forward layer (act = leaky ReLU or identity)
//   epilogue 1:  C  = acc * act'(mask[m,n])  for n < mask_cols  data gradient of a layer, times the derivative of the
//                                                              PREVIOUS layer's activation (mask = its saved output);
//                colsum[n] += sum_m C[m,n]                     = that layer's bias gradient, reduced in the same epilogue
//   epilogue 2:  C += acc   (vector reductions)                weight gradient, split over the 4096 rows (gridDim.z)
//
// Arithmetic: tcgen05.mma.kind::tf32 with fp32 accumulators in tensor memory; precision 0 evaluates every product as
// hi*hi + hi*lo + lo*hi (hi = tf32 truncation, lo = exact remainder): ~2^-21 relative per product, the same scheme as
// the field head (mlp_tc.cu); precision 1 = one pass on round-to-nearest tf32 operands.
//
// One CTA = one 128 x 64 tile of C.  K is consumed in chunks of 32 through a three-stage shared-memory ring: eight loader
// warps fetch a chunk from global memory into registers two chunks ahead (float4 loads; transposed operands through a 4x4
// register transpose), split it into hi / lo planes of the no-swizzle K-major UMMA layout and arrive on the stage's `full`
// mbarrier; a ninth warp waits for it, issues the 12 MMAs of the chunk from one elected lane and tcgen05.commit-s to the
// stage's `empty` mbarrier.  No CTA-wide barrier inside the K loop: loads, splits / stores and tensor-core work of different
// chunks overlap (with a __syncthreads per chunk they added up: 1.2 us per chunk, tools/time_gemm.py).
#include "common.cuh"
#include "umma.cuh"

namespace sanerf {

namespace gemm {
constexpr uint32_t kBM = 128, kBN = 64, kKC = 32, kThreads = 256 /* loader threads */, kStages = 3;
constexpr uint32_t kCtaThreads = kThreads + 32;                    // + one warp that only issues the MMAs
// Byte stride between consecutive 4-element K chunks of a tile with `rows` rows (the descriptors' leading byte offset):
// one 16-byte slot of padding per chunk column rotates the bank a (row, chunk) slot lands in, so a warp that writes
// 4 rows x 8 chunks (coalesced global loads: 8 lanes per 128-byte row segment) stores without bank conflicts.
__host__ __device__ constexpr uint32_t chunk_stride(uint32_t rows) { return rows * 16u + 16u; }
constexpr uint32_t kAPlane = chunk_stride(kBM) * (kKC / 4) , kBPlane = chunk_stride(kBN) * (kKC / 4);
static_assert(kAPlane % 128 == 0 && kBPlane % 128 == 0, "operand planes must stay 128-byte aligned");
static_assert(kKC == 32, "the loaders map 8 lanes to the 8 chunks of a 32-wide K slice");
constexpr uint32_t kStageBytes = 2 * kAPlane + 2 * kBPlane;           // A hi | A lo | B hi | B lo
constexpr uint32_t kSmem = kStages * kStageBytes + 128;
constexpr uint32_t kTmemCols = 64;
}  // namespace gemm

struct GemmParams {
    const float* A;
    const float* B;
    float* C;
    const float* bias;
    const float* mask;
    float* colsum;
    uint32_t lda, ldb, ldc, ldm;
    uint32_t M, N, K;
    uint32_t chunks_per_split;
    uint32_t stages;         // 2 or 3 (3 = one CTA per SM; 2 = two CTAs per SM when the grid has more CTAs than SMs)
    uint32_t mask_cols;
    int a_trans, b_trans, epilogue, act, precision;
    float slope;
};

__device__ __forceinline__ float gemm_round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void gemm_mbar_arrive(uint32_t saddr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(saddr) : "memory");
}

// 4 consecutive K elements of row r -> hi / lo planes of a chunk-major K-major tile with `rows` rows
__device__ __forceinline__ void gemm_put_chunk(uint8_t* hi_plane, uint8_t* lo_plane, uint32_t rows, uint32_t r, uint32_t chunk,
                                               float a, float b, float c, float d, bool split) {
    const uint32_t off = chunk * gemm::chunk_stride(rows) + r * 16u;
    if (split) {
        float h0, h1, h2, h3, l0, l1, l2, l3;
        umma::split_tf32(a, h0, l0); umma::split_tf32(b, h1, l1); umma::split_tf32(c, h2, l2); umma::split_tf32(d, h3, l3);
        *reinterpret_cast<float4*>(hi_plane + off) = make_float4(h0, h1, h2, h3);
        *reinterpret_cast<float4*>(lo_plane + off) = make_float4(l0, l1, l2, l3);
    } else {
        *reinterpret_cast<float4*>(hi_plane + off) =
            make_float4(gemm_round_tf32(a), gemm_round_tf32(b), gemm_round_tf32(c), gemm_round_tf32(d));
    }
}

// 4 consecutive floats starting at p[0] (element index `first` of a run whose valid indices are < limit); vec = the
// address is known to be 16-byte aligned
__device__ __forceinline__ float4 gemm_load4(const float* __restrict__ p, uint32_t first, uint32_t limit, bool vec) {
    if (vec && first + 4u <= limit) return __ldg(reinterpret_cast<const float4*>(p));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (first < limit) v.x = __ldg(p);
    if (first + 1u < limit) v.y = __ldg(p + 1);
    if (first + 2u < limit) v.z = __ldg(p + 2);
    if (first + 3u < limit) v.w = __ldg(p + 3);
    return v;
}

// One operand's share of a K chunk, held in registers between the global loads and the shared-memory stores.
//   normal  (src [rows][K], ld): item = (row, 4-element K chunk), 8 consecutive lanes = the 128 contiguous bytes of one
//           row segment (4 lines per warp load instead of 32); ROWS*8 items, ROWS*8/256 per thread
//   transposed (src [K][rows], ld): item = (4 rows, 4 K) block = 4 float4 loads along the contiguous row index; lanes
//           vary fastest along K so that the four transposed 16-byte stores of a warp spread over all banks
template <uint32_t ROWS>
struct OperandRegs {
    static constexpr uint32_t kItems = ROWS * 8u / gemm::kThreads;       // normal orientation: 4 (A) or 2 (B)
    float4 v[4];

    __device__ __forceinline__ void load(const float* __restrict__ src, uint32_t ld, bool trans, bool vec, uint32_t row0,
                                         uint32_t row_limit, uint32_t k0, uint32_t k_limit, uint32_t tid) {
        if (!trans) {
#pragma unroll
            for (uint32_t j = 0; j < kItems; ++j) {
                const uint32_t id = tid + gemm::kThreads * j;
                const uint32_t row = id >> 3, chunk = id & 7u;
                const uint32_t r = row0 + row, k = k0 + chunk * 4u;
                v[j] = (r < row_limit) ? gemm_load4(src + (size_t)r * ld + k, k, k_limit, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else {
            const uint32_t k4 = tid & 7u, row4 = tid >> 3;
            const bool mine = row4 < ROWS / 4u;
#pragma unroll
            for (uint32_t i = 0; i < 4; ++i) {
                const uint32_t k = k0 + k4 * 4u + i, r = row0 + row4 * 4u;
                v[i] = (mine && k < k_limit) ? gemm_load4(src + (size_t)k * ld + r, r, row_limit, vec)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }

    __device__ __forceinline__ void store(uint8_t* hi, uint8_t* lo, bool trans, bool split, uint32_t tid) const {
        if (!trans) {
#pragma unroll
            for (uint32_t j = 0; j < kItems; ++j) {
                const uint32_t id = tid + gemm::kThreads * j;
                gemm_put_chunk(hi, lo, ROWS, id >> 3, id & 7u, v[j].x, v[j].y, v[j].z, v[j].w, split);
            }
        } else {
            const uint32_t k4 = tid & 7u, row4 = tid >> 3;
            if (row4 < ROWS / 4u) {
                gemm_put_chunk(hi, lo, ROWS, row4 * 4u + 0u, k4, v[0].x, v[1].x, v[2].x, v[3].x, split);
                gemm_put_chunk(hi, lo, ROWS, row4 * 4u + 1u, k4, v[0].y, v[1].y, v[2].y, v[3].y, split);
                gemm_put_chunk(hi, lo, ROWS, row4 * 4u + 2u, k4, v[0].z, v[1].z, v[2].z, v[3].z, split);
                gemm_put_chunk(hi, lo, ROWS, row4 * 4u + 3u, k4, v[0].w, v[1].w, v[2].w, v[3].w, split);
            }
        }
    }
};

__global__ void __launch_bounds__(gemm::kCtaThreads) gemm_tc_kernel(const GemmParams p) {
    pdl_begin();
    using namespace gemm;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages], s_done;
    __shared__ uint32_t s_tmem;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127u) & ~uintptr_t(127));

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;
    const uint32_t total_chunks = (p.K + kKC - 1u) / kKC;
    const uint32_t c_begin = blockIdx.z * p.chunks_per_split;
    const uint32_t c_end = min(c_begin + p.chunks_per_split, total_chunks);
    if (c_begin >= c_end) return;                                    // empty split of a reduction epilogue (uniform)
    const uint32_t nc = c_end - c_begin;
    const bool split = ((p.precision & 15) == 0);
    const int dbg = p.precision >> 4;      // diagnostics (tools/time_gemm.py): 1 = no MMA, 2 = no shared stores, 4 = no global loads

    if (warp == 0) umma::tmem_alloc<kTmemCols>(umma::smem_u32(&s_tmem));
    if (tid == 32) {
        for (uint32_t s = 0; s < kStages; ++s) {
            umma::mbar_init(umma::smem_u32(&s_full[s]), kThreads / 32);     // one arrival per loader warp
            umma::mbar_init(umma::smem_u32(&s_empty[s]), 1);                // tcgen05.commit of the chunk that used the stage
        }
        umma::mbar_init(umma::smem_u32(&s_done), 1);
        umma::fence_mbar_init();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s_tmem;

    const bool a_vec = (p.lda % 4u == 0u) && ((reinterpret_cast<uintptr_t>(p.A) & 15u) == 0u);
    const bool b_vec = (p.ldb % 4u == 0u) && ((reinterpret_cast<uintptr_t>(p.B) & 15u) == 0u);
    const bool a_tr = p.a_trans != 0, b_tr = p.b_trans != 0;

    constexpr uint32_t idesc = umma::idesc_tf32(kBM, kBN, 0, 0);
    constexpr uint32_t dhi = umma::desc_hi(128u, umma::kLayoutNone);
    constexpr uint32_t a_step = (2u * chunk_stride(kBM)) >> 4, b_step = (2u * chunk_stride(kBN)) >> 4;   // 8 K elements = 2 chunks

    if (warp == kThreads / 32) {
        // ===================== issuing warp: waits for a full stage, issues the chunk's MMAs, releases the stage ===========
        for (uint32_t c = 0; c < nc; ++c) {
            const uint32_t s = c % p.stages;
            umma::mbar_wait(umma::smem_u32(&s_full[s]), (c / p.stages) & 1u);
            if (umma::elect_one()) {
                umma::fence_after_sync();
                const uint32_t sa = umma::smem_u32(smem + s * kStageBytes);
                const uint32_t dAh = umma::desc_lo(sa, chunk_stride(kBM)), dAl = umma::desc_lo(sa + kAPlane, chunk_stride(kBM));
                const uint32_t dBh = umma::desc_lo(sa + 2 * kAPlane, chunk_stride(kBN)), dBl = umma::desc_lo(sa + 2 * kAPlane + kBPlane, chunk_stride(kBN));
                uint32_t acc = (c > 0u) ? 1u : 0u;
#pragma unroll
                for (uint32_t ks = 0; ks < kKC / 8u; ++ks) {
                    if ((dbg & 1) && !(c == 0u && ks == 0u)) break;
                    const uint32_t ao = ks * a_step, bo = ks * b_step;
                    if (split) {
                        umma::mma_tf32_ss2(tmem, dAl + ao, dhi, dBh + bo, dhi, idesc, acc);
                        umma::mma_tf32_ss2(tmem, dAh + ao, dhi, dBl + bo, dhi, idesc, 1u);
                        acc = 1u;
                    }
                    umma::mma_tf32_ss2(tmem, dAh + ao, dhi, dBh + bo, dhi, idesc, acc);
                    acc = 1u;
                }
                umma::commit(umma::smem_u32(&s_empty[s]));
                if (c + 1u == nc) umma::commit(umma::smem_u32(&s_done));
            }
            __syncwarp();
        }
    } else {
        // ===================== loader warps: global -> registers (two chunks ahead) -> hi / lo planes of a free stage =====
        // No CTA-wide barrier per chunk: a warp that has stored its share arrives on the stage's `full` barrier and moves on;
        // it blocks only when the ring is full (`empty` = the MMAs that read the stage have completed).
        OperandRegs<kBM> ra0, ra1;
        OperandRegs<kBN> rb0, rb1;
        auto fetch = [&](OperandRegs<kBM>& ra, OperandRegs<kBN>& rb, uint32_t c) {
            if ((dbg & 4) && c >= 2u) return;
            ra.load(p.A, p.lda, a_tr, a_vec, m0, p.M, (c_begin + c) * kKC, p.K, tid);
            rb.load(p.B, p.ldb, b_tr, b_vec, n0, p.N, (c_begin + c) * kKC, p.K, tid);
        };
        auto consume = [&](const OperandRegs<kBM>& ra, const OperandRegs<kBN>& rb, uint32_t c) {
            const uint32_t s = c % p.stages;
            uint8_t* stage = smem + s * kStageBytes;
            if (c >= p.stages) umma::mbar_wait(umma::smem_u32(&s_empty[s]), ((c / p.stages) - 1u) & 1u);
            if (!(dbg & 2)) {
                ra.store(stage, stage + kAPlane, a_tr, split, tid);
                rb.store(stage + 2 * kAPlane, stage + 2 * kAPlane + kBPlane, b_tr, split, tid);
            }
            umma::fence_proxy_async();                   // this thread's generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) gemm_mbar_arrive(umma::smem_u32(&s_full[s]));
        };
        fetch(ra0, rb0, 0u);
        if (nc > 1u) fetch(ra1, rb1, 1u);
        for (uint32_t c = 0; c < nc; c += 2u) {
            consume(ra0, rb0, c);
            if (c + 2u < nc) fetch(ra0, rb0, c + 2u);
            if (c + 1u < nc) {
                consume(ra1, rb1, c + 1u);
                if (c + 3u < nc) fetch(ra1, rb1, c + 3u);
            }
        }
    }
    umma::mbar_wait(umma::smem_u32(&s_done), 0u);        // every MMA of this CTA has completed
    umma::fence_after_sync();

    // ---- epilogue: warp w < 8 reads TMEM lanes 32 (w & 3) .. +31 (rows), columns 32 (w >> 2) .. +31
    if (warp < kThreads / 32) {
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
            } else if (p.epilogue == 1) {
                const float* mrow = p.mask + (size_t)m * p.ldm + nb;
                const bool m_vec = (p.ldm % 4u == 0u) && ((reinterpret_cast<uintptr_t>(p.mask) & 15u) == 0u) &&
                                   nb + 16u <= p.mask_cols && nb + 16u <= p.N;
                if (m_vec) {                                    // a thread owns a row: 4 x 16-byte loads instead of 16 scalar ones
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
            if (p.epilogue == 2) {
                if (c_vec && nb + 16u <= p.N) {
#pragma unroll
                    for (uint32_t j = 0; j < 16; j += 4) red_add_v4_f32(dst + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (uint32_t j = 0; j < 16; ++j)
                        if (nb + j < p.N) red_add_f32(dst + j, v[j]);
                }
            } else {
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
        if (p.colsum != nullptr && p.epilogue == 1 && nb < p.mask_cols) {     // warp-uniform: bias gradient of the layer below
#pragma unroll
            for (uint32_t j = 0; j < 16; ++j) {
                const float t = warp_sum((m < p.M) ? v[j] : 0.0f);
                if (lane == 0 && nb + j < p.mask_cols && nb + j < p.N) red_add_f32(p.colsum + nb + j, t);
            }
        }
    }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<kTmemCols>(tmem);
}

// out[n] += sum_m act'(mask) is NOT applied here: X is already the pre-activation gradient.  out[n] += sum_m X[m, n].
__global__ void __launch_bounds__(256) colsum_add_kernel(const float* __restrict__ X, uint32_t ld, uint32_t M, uint32_t N,
                                                         uint32_t rows_per_block, float* __restrict__ out) {
    pdl_begin();
    __shared__ float part[8][33];
    const uint32_t cx = threadIdx.x & 31u, ry = threadIdx.x >> 5;
    const uint32_t n = blockIdx.x * 32u + cx;
    const uint32_t r0 = blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, M);
    float acc = 0.0f;
    if (n < N)
        for (uint32_t r = r0 + ry; r < r1; r += 8u) acc += __ldg(X + (size_t)r * ld + n);
    part[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && n < N) {
        float t = 0.0f;
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) t += part[i][cx];
        red_add_f32(out + n, t);
    }
}

// csrc/gemm_tma.cu: the TMA-fed variant for plain forward products (both operands row-major, 16-byte aligned rows);
// returns -1 when the call is not eligible
int launch_gemm_tma(const float* A, uint32_t lda, const float* B, uint32_t ldb, int b_trans, float* C, uint32_t ldc, uint32_t M,
                    uint32_t N, uint32_t K, int epilogue, const float* bias, int act, float slope, const float* mask, uint32_t ldm,
                    uint32_t mask_cols, int precision, cudaStream_t stream);

static bool gemm_tma_enabled() {
    static const bool on = [] {
        const char* e = getenv("SANERF_GEMM_TMA");
        return e == nullptr || e[0] != '0';
    }();
    return on;
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_gemm_tc(const float* A, uint32_t lda, int a_trans, const float* B, uint32_t ldb, int b_trans, float* C,
                              uint32_t ldc, uint32_t M, uint32_t N, uint32_t K, uint32_t k_splits, int epilogue,
                              const float* bias, int act, float slope, const float* mask, uint32_t ldm, uint32_t mask_cols,
                              float* colsum, int precision, void* stream) {
    if (M == 0 || N == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(A);
    SANERF_REQUIRE_PTR(B);
    SANERF_REQUIRE_PTR(C);
    if (K == 0) return fail(SANERF_ERR_INVALID_ARG, "gemm_tc: K must be positive");
    if (epilogue < 0 || epilogue > 2) return fail(SANERF_ERR_INVALID_ARG, "gemm_tc: epilogue must be 0, 1 or 2");
    if (epilogue == 1 && mask == nullptr) return fail(SANERF_ERR_NULL_POINTER, "gemm_tc: mask is NULL");
    if ((precision & 15) != 0 && (precision & 15) != 1) return fail(SANERF_ERR_INVALID_ARG, "gemm_tc: precision must be 0 or 1");
    if (k_splits == 0) k_splits = 1;
    if (k_splits > 1 && epilogue != 2)
        return fail(SANERF_ERR_INVALID_ARG, "gemm_tc: K can only be split with the accumulating epilogue");
    if (!a_trans && epilogue != 2 && k_splits == 1 && colsum == nullptr && (precision >> 4) == 0 && gemm_tma_enabled()) {
        const int rc = launch_gemm_tma(A, lda, B, ldb, b_trans, C, ldc, M, N, K, epilogue, bias, act, slope, mask, ldm, mask_cols,
                                       precision, static_cast<cudaStream_t>(stream));
        if (rc != -1) return rc;                      // -1: operands not TMA-addressable (row stride not a multiple of 16 bytes)
    }
    const uint32_t chunks = (K + gemm::kKC - 1u) / gemm::kKC;
    GemmParams p;
    p.A = A; p.B = B; p.C = C; p.bias = bias; p.mask = mask; p.colsum = colsum;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.ldm = ldm;
    p.M = M; p.N = N; p.K = K;
    p.chunks_per_split = (chunks + k_splits - 1u) / k_splits;
    p.mask_cols = mask_cols;
    p.a_trans = a_trans; p.b_trans = b_trans; p.epilogue = epilogue; p.act = act; p.precision = precision;
    p.slope = slope;
    dim3 grid(div_up(M, gemm::kBM), div_up(N, gemm::kBN), k_splits);
    p.stages = ((size_t)grid.x * grid.y * grid.z > (size_t)kNumSMs) ? 2u : gemm::kStages;     // second wave -> co-residency
    const size_t smem = (size_t)p.stages * gemm::kStageBytes + 128;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm::kSmem);
        if (e != cudaSuccess) return fail(SANERF_ERR_CUDA, "gemm_tc: %s", cudaGetErrorString(e));
        configured = true;
    }
    SANERF_LAUNCH(gemm_tc_kernel, grid, gemm::kCtaThreads, smem, static_cast<cudaStream_t>(stream), p);
    return check_launch("gemm_tc_kernel");
}

extern "C" int sanerf_colsum_add(const float* X, uint32_t ld, uint32_t M, uint32_t N, float* out, void* stream) {
    if (M == 0 || N == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(X);
    SANERF_REQUIRE_PTR(out);
    const uint32_t rows_per_block = 256;
    dim3 grid(div_up(N, 32u), div_up(M, rows_per_block), 1);
    SANERF_LAUNCH(colsum_add_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), X, ld, M, N, rows_per_block, out);
    return check_launch("colsum_add_kernel");
}
