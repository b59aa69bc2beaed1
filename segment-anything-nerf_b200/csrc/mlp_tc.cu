// Field head, fused on the 5th-generation tensor cores (sm_100a): hash-grid encode -> Linear(32,64) -> ReLU ->
// Linear(64,64) -> ReLU -> Linear(64,16), and its backward.
//
// Replaces, for the final sampling level, GridEncoder(L16,F2) + grid_mlp of nerf/network.py:102-103, 221-227
// (`h = self.grid(x); f = self.grid_mlp(h)`), which the reference runs as one gather kernel, one permute copy and
// three cuBLAS SGEMMs with three activation round trips through HBM (185 MB per 2^18 samples).
//
// Design
//  * A tile is 128 samples = the 128 TMEM lanes.  Weights live in shared memory for the whole kernel as UMMA
//    B operands; every A operand lives in TENSOR MEMORY: the producers write the encoding there (tcgen05.st), the
//    activations go TMEM accumulator -> registers (ReLU, hi/lo split) -> TMEM operand planes.  The unified L1 / shared
//    data path is what the hash-grid gather is bound by (one 128-byte wavefront per lane and load), so no activation
//    tile travels through it, and with 56 KB of shared memory the L1 keeps ~170 KB for the table.
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
constexpr uint32_t kFwdSmem = oW3 + 2 * kW3;       // forward: only the weights live in shared memory
// forward TMEM columns: two encoding slots (hi 32 | lo 32 each), hidden activations (hi 64 | lo 64), three accumulators
constexpr uint32_t cEnc = 0, cHhi = 128, cHlo = 192, cD1 = 256, cD2 = 320, cD3 = 384, kTmemCols = 512;
}  // namespace head

// float offset of the 4-element chunk c of sample b in a "tile-chunk-major" [B, W] matrix: [tile][chunk][row][4].
// The saved activations (enc, H1, H2) use it: a warp that writes / reads one chunk of 32 consecutive samples touches
// 512 contiguous bytes, whether it works sample-per-thread (forward epilogue) or any-lane (backward staging).
__host__ __device__ __forceinline__ size_t tcm_off(uint32_t b, uint32_t chunk, uint32_t chunks_per_row) {
    return (((size_t)(b >> 7) * chunks_per_row + chunk) * 128u + (b & 127u)) * 4u;
}

struct HeadFwdParams {
    const float* x01;        // [B,3] in [0,1]^3, or NULL: take the encoding from `enc_in`
    const float* table;      // [rows,2] fp32
    const int32_t* offsets;  // [17]
    const float* enc_in;     // [B,32] when x01 == NULL
    const float* w1;         // [64,32]  nn.Linear layout [out,in]
    const float* w2;         // [64,64]
    const float* w3;         // [16,64]
    float* enc_out;          // [B,32] or NULL
    float* h1_out;           // [B,64] or NULL: relu(layer 1), saved for the backward
    float* h2_out;           // [B,64] or NULL: relu(layer 2)
    float* out;              // [B,16]
    uint32_t B, H;
    float S;
    int precision;           // 0: 3xTF32 (fp32 parity), 1: single tf32 pass
    // Chunked front-to-back evaluation (early ray termination): when ray_list != NULL the kernel evaluates samples
    // [chunk * chunk_len, (chunk + 1) * chunk_len) of the rays ray_list[0 .. *list_count) only; compact index b' maps to
    // sample (ray_list[b' / chunk_len], chunk * chunk_len + b' % chunk_len) of the [N, T] layout of x01 / out.
    const uint32_t* ray_list;
    const uint32_t* list_count;
    uint32_t chunk, chunk_len, T;
};

// compact index -> row of x01 / out; returns false past the end of the work list
__device__ __forceinline__ bool head_row(const HeadFwdParams& p, uint32_t work, uint32_t bprime, uint32_t& row) {
    if (bprime >= work) return false;
    if (p.ray_list == nullptr) { row = bprime; return true; }
    const uint32_t q = bprime / p.chunk_len, j = bprime - q * p.chunk_len;
    row = __ldg(p.ray_list + q) * p.T + p.chunk * p.chunk_len + j;
    return true;
}

__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// 256-bit store of eight consecutive floats (one full 32-byte sector)
__device__ __forceinline__ void stg_f8(float* p, const float (&v)[8]) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
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

// 16 / 8 consecutive columns of this thread's TMEM lane <- hi / lo parts of v (A-operand planes of the next product)
__device__ __forceinline__ void split_to_tmem16(uint32_t t_hi, uint32_t t_lo, const float (&v)[16], bool split) {
    float hi[16], lo[16];
#pragma unroll
    for (uint32_t j = 0; j < 16; ++j) {
        if (split) umma::split_tf32(v[j], hi[j], lo[j]);
        else hi[j] = round_tf32(v[j]);
    }
    umma::tmem_st16(t_hi, hi);
    if (split) umma::tmem_st16(t_lo, lo);
}
__device__ __forceinline__ void split_to_tmem8(uint32_t t_hi, uint32_t t_lo, const float (&v)[8], bool split) {
    float hi[8], lo[8];
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j) {
        if (split) umma::split_tf32(v[j], hi[j], lo[j]);
        else hi[j] = round_tf32(v[j]);
    }
    umma::tmem_st8(t_hi, hi);
    if (split) umma::tmem_st8(t_lo, lo);
}

// D[128, N] = A[tmem: 128 lanes x K columns] . B[N, K]^T ; B = chunk-major K-major weight tile with N rows; b_* are low
// descriptor halves (umma::desc_lo), the high halves are compile-time constants
template <uint32_t N, uint32_t K>
__device__ __forceinline__ void issue_gemm_ts(uint32_t tmem_d, uint32_t tmem_a_hi, uint32_t tmem_a_lo, uint32_t b_hi,
                                              uint32_t b_lo, bool split) {
    constexpr uint32_t idesc = umma::idesc_tf32(128, N, 0, 0);
    constexpr uint32_t b_step = (2u * N * 16u) >> 4;
    constexpr uint32_t hi = umma::desc_hi(128u, umma::kLayoutNone);
    uint32_t acc = 0;
#pragma unroll
    for (uint32_t ks = 0; ks < K / 8; ++ks) {
        const uint32_t bo = ks * b_step;
        if (split) {
            umma::mma_tf32_ts2(tmem_d, tmem_a_lo + ks * 8u, b_hi + bo, hi, idesc, acc);
            umma::mma_tf32_ts2(tmem_d, tmem_a_hi + ks * 8u, b_lo + bo, hi, idesc, 1u);
            acc = 1;
        }
        umma::mma_tf32_ts2(tmem_d, tmem_a_hi + ks * 8u, b_hi + bo, hi, idesc, acc);
        acc = 1;
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

// D[128, N] = A[128, K] . B[N, K]^T over chunk-major K-major operands; one thread issues.  a_* / b_* are the LOW
// descriptor halves of the hi / lo planes (umma::desc_lo); the high halves are compile-time constants.
template <uint32_t N, uint32_t K, uint32_t B_ROWS>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                           bool split) {
    constexpr uint32_t idesc = umma::idesc_tf32(128, N, 0, 0);
    constexpr uint32_t a_step = (2u * head::kTile * 16u) >> 4, b_step = (2u * B_ROWS * 16u) >> 4;   // 8 K elements = 2 chunks
    constexpr uint32_t hi = umma::desc_hi(128u, umma::kLayoutNone);
    uint32_t acc = 0;
#pragma unroll
    for (uint32_t ks = 0; ks < K / 8; ++ks) {
        const uint32_t ao = ks * a_step, bo = ks * b_step;
        if (split) {
            umma::mma_tf32_ss2(tmem_d, a_lo + ao, hi, b_hi + bo, hi, idesc, acc);
            umma::mma_tf32_ss2(tmem_d, a_hi + ao, hi, b_lo + bo, hi, idesc, 1u);
            acc = 1;
        }
        umma::mma_tf32_ss2(tmem_d, a_hi + ao, hi, b_hi + bo, hi, idesc, acc);
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
// TMEM accumulator (64 columns of this thread's lane) -> ReLU -> hi / lo operand planes of the next layer in TMEM.
// `save` = the tile-chunk-major [B,64] activation kept for the backward (or NULL): chunk c of sample b at tcm_off(b, c, 16)
__device__ __forceinline__ void relu_to_tmem(uint32_t lane_base, uint32_t c_acc, bool split, float* save, uint32_t b) {
#pragma unroll
    for (uint32_t c0 = 0; c0 < head::kHid; c0 += 16) {
        float v[16];
        umma::tmem_ld16(lane_base + c_acc + c0, v);
#pragma unroll
        for (uint32_t j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
        split_to_tmem16(lane_base + head::cHhi + c0, lane_base + head::cHlo + c0, v, split);
        if (save != nullptr) {
#pragma unroll
            for (uint32_t j = 0; j < 16; j += 4)
                __stcs(reinterpret_cast<float4*>(save + tcm_off(b, (c0 + j) >> 2, 16)), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        }
    }
}

// Phase timestamps of CTA 0 (diagnostics; compiled in only with -DSANERF_HEAD_TRACE): [thread 0][tile 0..3][11]
#ifdef SANERF_HEAD_TRACE
static int g_head_dbg_host = 0;
#define HEAD_DBG(bit) ((p.dbg & (bit)) != 0)
__device__ long long g_head_trace[2 * 4 * 16];
__device__ long long g_head_marks[8];          // CTA 0, thread 0: entry, prologue done, loop done, exit
#define HEAD_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_head_marks[i] = clock64(); } while (0)
#define HEAD_TRACE(slot)                                                                                   \
    do {                                                                                                   \
        if (blockIdx.x == 0 && it < 4 && (tid == 0 || tid == 128))                                         \
            g_head_trace[((tid == 0 ? 0 : 1) * 4 + it) * 16 + (slot)] = clock64();                          \
    } while (0)
#else
#define HEAD_TRACE(slot) do {} while (0)
#define HEAD_MARK(i) do {} while (0)
#define HEAD_DBG(bit) false
#endif

__global__ void __launch_bounds__(head::kThreads, 1) head_forward_kernel(const HeadFwdParams p, const uint32_t tiles_arg) {
    HEAD_MARK(0);
    pdl_begin();
    HEAD_MARK(1);
    using namespace head;
    // number of samples to evaluate: all B, or chunk_len per ray still alive (the count lives on the device: no host sync)
    const uint32_t work = (p.ray_list != nullptr) ? __ldg(p.list_count) * p.chunk_len : p.B;
    const uint32_t tiles = (p.ray_list != nullptr) ? (work + kTile - 1) / kTile : tiles_arg;
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
    HEAD_MARK(2);

    if (warp >= kMlpWarps) {
        // ===================== producers: encoding of one sample x 4 levels -> A-operand slot =====================
        const uint32_t g = tid - kMlpWarps * 32;
        const uint32_t r = g & (kTile - 1), lg = g >> 7;
        const uint32_t enc_lane = umma::tmem_addr(tmem, (warp & 3u) * 32u, 0);      // r == 32 (warp & 3) + lane
        const float* __restrict__ table = p.table;
        for (uint32_t it = 0, tile = blockIdx.x; tile < tiles; ++it, tile += gridDim.x) {
            const uint32_t slot = it & 1u;
            umma::mbar_wait(umma::smem_u32(&s_empty[slot]), ((it >> 1) & 1u) ^ 1u);
            uint32_t b = 0;
            const bool live = head_row(p, work, tile * kTile + r, b);
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
                    for (uint32_t k = 0; k < 8; k += 2) {
                        if (zero) {
                            val[q][k] = val[q][k + 1] = make_float2(0.0f, 0.0f);
                        } else {
                            gather_pair_f2(table + base * 2, corner_row<3>(geo, cell, k), corner_row<3>(geo, cell, k + 1),
                                           val[q][k], val[q][k + 1]);
                        }
                    }
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
                if (p.enc_out != nullptr && live) {          // tile-chunk-major: 32 consecutive samples = 512 contiguous bytes
                    // streaming (evict-first) stores: 640 B/sample of saved activations must not push the table out of L2
                    __stcs(reinterpret_cast<float4*>(p.enc_out + tcm_off(b, lg * 2, 8)), make_float4(enc[0], enc[1], enc[2], enc[3]));
                    __stcs(reinterpret_cast<float4*>(p.enc_out + tcm_off(b, lg * 2 + 1, 8)), make_float4(enc[4], enc[5], enc[6], enc[7]));
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
            // this sample's 8 features -> TMEM operand planes of slot `slot` (this warp owns TMEM lanes 32 (warp & 3) ..)
            split_to_tmem8(enc_lane + cEnc + slot * 64u + lg * 8u, enc_lane + cEnc + slot * 64u + 32u + lg * 8u, enc, split);
            umma::tmem_st_wait();
            umma::fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(umma::smem_u32(&s_full[slot]));
        }
    } else {
        // ===================== MLP warps: thread t <-> TMEM lane t <-> sample row t of the tile =====================
        const uint32_t lane_base = umma::tmem_addr(tmem, warp * 32, 0);
        // low descriptor halves (K-major: LBO = rows * 16 = next 4-element K chunk)
        const uint32_t dW1h = umma::desc_lo(umma::smem_u32(smem + oW1), kHid * 16u), dW1l = umma::desc_lo(umma::smem_u32(smem + oW1 + kW1), kHid * 16u);
        const uint32_t dW2h = umma::desc_lo(umma::smem_u32(smem + oW2), kHid * 16u), dW2l = umma::desc_lo(umma::smem_u32(smem + oW2 + kW2), kHid * 16u);
        const uint32_t dW3h = umma::desc_lo(umma::smem_u32(smem + oW3), kOut * 16u), dW3l = umma::desc_lo(umma::smem_u32(smem + oW3 + kW3), kOut * 16u);
        uint32_t mma_phase = 0;
        for (uint32_t it = 0, tile = blockIdx.x; tile < tiles; ++it, tile += gridDim.x) {
            const uint32_t slot = it & 1u;
            if (warp == 0) {              // warp-uniform branch + elected lane: the issue loop stays on the uniform datapath
                umma::mbar_wait(umma::smem_u32(&s_full[slot]), (it >> 1) & 1u);
                if (it == 0) HEAD_MARK(3);
                if (umma::elect_one()) {
                    umma::fence_after_sync();
                    issue_gemm_ts<kHid, kIn>(tmem + cD1, tmem + cEnc + slot * 64u, tmem + cEnc + slot * 64u + 32u, dW1h, dW1l, split);
                    umma::commit(umma::smem_u32(&s_empty[slot]));     // A slot free once layer 1 has consumed it
                    umma::commit(umma::smem_u32(&s_mma));
                }
                __syncwarp();
            }
            umma::mbar_wait(umma::smem_u32(&s_mma), mma_phase); mma_phase ^= 1u;
            umma::fence_after_sync();
            uint32_t b = 0;
            const bool row_live = head_row(p, work, tile * kTile + tid, b);
            const bool keep = (p.h1_out != nullptr) && row_live;
            relu_to_tmem(lane_base, cD1, split, keep ? p.h1_out : nullptr, b);
            umma::tmem_st_wait();
            umma::fence_before_sync();
            named_bar_sync(1, kMlpWarps * 32);
            if (warp == 0) {
                if (umma::elect_one()) {
                    umma::fence_after_sync();
                    issue_gemm_ts<kHid, kHid>(tmem + cD2, tmem + cHhi, tmem + cHlo, dW2h, dW2l, split);
                    umma::commit(umma::smem_u32(&s_mma));
                }
                __syncwarp();
            }
            umma::mbar_wait(umma::smem_u32(&s_mma), mma_phase); mma_phase ^= 1u;
            umma::fence_after_sync();
            relu_to_tmem(lane_base, cD2, split, keep ? p.h2_out : nullptr, b);     // layer 2 has completed: the H planes are free
            umma::tmem_st_wait();
            umma::fence_before_sync();
            named_bar_sync(1, kMlpWarps * 32);
            if (warp == 0) {
                if (umma::elect_one()) {
                    umma::fence_after_sync();
                    issue_gemm_ts<kOut, kHid>(tmem + cD3, tmem + cHhi, tmem + cHlo, dW3h, dW3l, split);
                    umma::commit(umma::smem_u32(&s_mma));
                }
                __syncwarp();
            }
            umma::mbar_wait(umma::smem_u32(&s_mma), mma_phase); mma_phase ^= 1u;
            umma::fence_after_sync();
            float v[16];
            umma::tmem_ld16(lane_base + cD3, v);
            if (row_live) {                 // two 256-bit stores: every lane fills whole 32-byte sectors
                const float lo8[8] = {v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]};
                const float hi8[8] = {v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]};
                stg_f8(p.out + (size_t)b * kOut, lo8);
                stg_f8(p.out + (size_t)b * kOut + 8, hi8);
            }
            umma::fence_before_sync();   // orders this tile's TMEM reads before the next tile's MMAs (issued after bar 1)
            if (it == 0) HEAD_MARK(4);
        }
        HEAD_MARK(5);
    }
    umma::fence_before_sync();
    __syncthreads();
    HEAD_MARK(6);
    if (warp == 0) umma::tmem_dealloc<head::kTmemCols>(tmem);
}


// ================================================= backward =================================================
// Inputs per sample: enc [32], H1 = relu(D1) [64], H2 = relu(D2) [64] (saved by the forward), g_out [16].
// Per 128-sample tile:
//   3. dW3^T += H2^T G3 ;  DG2 = G3 W3        G2 = DG2 * [H2 > 0]
//   4. dW2   += G2^T H1 ;  DG1 = G2 W2        G1 = DG1 * [H1 > 0]
//   5. dW1   += G1^T A1 ;  DGE = G1 W1        g_enc = DGE -> HBM
// Operand placement
//   * data-gradient products (contraction over features): A = the gradient tile in TENSOR MEMORY (lane = sample,
//     columns = features, hi/lo planes written with tcgen05.st), B = the TRANSPOSED weight, K-major in shared memory;
//   * weight-gradient products (contraction over the 128 samples): both operands MN-major in shared memory.  32-bit
//     MN-major operands exist only in the 128-byte-swizzle / 32-byte-atom layout (umma::mn32_off), so every
//     activation / gradient tile is written once more in that layout: a row of the tile is a contraction row, each
//     thread writes its own sample's row with 16-byte stores.
//   * the three weight gradients accumulate in tensor memory (M = 64 accumulators) over all tiles of the persistent
//     CTA and are added to HBM once per CTA.
// 16 symmetric worker warps: each prefetches its share of the NEXT tile's rows into registers (coalesced,
// tile-chunk-major saved activations) while the current tile computes, stages them at the tile boundary, and runs the
// epilogue of one 16-column group of its TMEM lane quadrant; one elected lane of the last warp issues the MMAs.
namespace head {
constexpr uint32_t kMn64 = kTile * 128u * 2u;      // one plane of a 64-wide MN-major tile (two 32-wide blocks)
constexpr uint32_t kMn32 = kTile * 128u;           // one plane of a 32-wide (or padded 16-wide) MN-major tile
constexpr uint32_t kMnLbo = kTile * 128u, kMnSbo = 512u;
// transposed weights (K-major, chunk-major, hi then lo): W3^T [64][16], W2^T [64][64], W1^T [32][64]
constexpr uint32_t oT3 = 0, oT2 = oT3 + 2 * kW3, oT1 = oT2 + 2 * kW2;
constexpr uint32_t oBufA = oT1 + 2 * kW1;          // H1 (64-wide), later A1 (32-wide)
constexpr uint32_t oBufB = oBufA + 2 * kMn64;      // H2, then G2, then G1 (64-wide)
constexpr uint32_t oBufG = oBufB + 2 * kMn64;      // G3 (16-wide, padded)
constexpr uint32_t kBwdSmem = oBufG + 2 * kMn32 + 1024;     // + slack to align the base to 1024 bytes
static_assert(oBufA % 1024 == 0 && oBufB % 1024 == 0 && oBufG % 1024 == 0, "swizzled tiles need aligned bases");
static_assert(kBwdSmem <= 227 * 1024, "backward tiles exceed shared memory");
// TMEM columns: A-operand staging (hi, lo), data-gradient accumulator, weight-gradient accumulators (M = 64)
constexpr uint32_t cAhi = 0, cAlo = 64, cDG = 128, cW2 = 192, cW1 = 256, cW3 = 288, cDGE = 304 /* two 32-column buffers */,
                   kBwdTmemCols = 512;
constexpr uint32_t kBwdThreads = 512;                   // worker threads (and the whole lock-step kernel)
constexpr uint32_t kBwdThreads2 = kBwdThreads + 32;      // + one issuing warp
}  // namespace head

struct HeadBwdParams {
    const float* enc;     // [B,32]
    const float* h1;      // [B,64]
    const float* h2;      // [B,64]
    const float* g_out;   // [B,16]
    const float* w1;
    const float* w2;
    const float* w3;
    float* g_enc;         // [B,32] or NULL
    // fused hash-grid scatter (all NULL / 0: g_enc is written instead)
    const float* x01;        // [B,3]
    const int32_t* offsets;  // [17]
    float* g_table;          // [rows,2] accumulated
    float S;
    uint32_t H;
    float* g_w1;          // [64,32] accumulated
    float* g_w2;          // [64,64]
    float* g_w3;          // [16,64]
    uint32_t B;
    int precision;
    int dbg;              // diagnostics (SANERF_HEAD_TRACE builds): phases to leave out when timing
};

// stage the TRANSPOSE of an nn.Linear weight [rows=out, cols=in] as a K-major B operand with `cols` rows
__device__ __forceinline__ void stage_weight_t(const float* __restrict__ w, uint8_t* plane_hi, uint32_t rows, uint32_t cols,
                                               bool split, uint32_t tid, uint32_t nthreads) {
    uint8_t* plane_lo = plane_hi + rows * cols * 4;
    for (uint32_t i = tid; i < rows * cols; i += nthreads) {
        const uint32_t r = i / cols, c = i - r * cols;          // W[r][c] -> tile row c, contraction index r
        const uint32_t off = umma::tile_off(cols, c, r);
        const float v = __ldg(w + i);
        if (split) {
            float hi, lo;
            umma::split_tf32(v, hi, lo);
            *reinterpret_cast<float*>(plane_hi + off) = hi;
            *reinterpret_cast<float*>(plane_lo + off) = lo;
        } else {
            *reinterpret_cast<float*>(plane_hi + off) = round_tf32(v);
        }
    }
}

// four consecutive features f..f+3 (f % 4 == 0) of contraction row k -> hi / lo planes of an MN-major tile
__device__ __forceinline__ void put_mn(uint8_t* hi_plane, uint8_t* lo_plane, uint32_t k, uint32_t f, float a, float b,
                                       float c, float d, bool split) {
    const uint32_t off = umma::mn32_off(head::kTile, k, f);
    if (split) {
        float h0, h1, h2, h3, l0, l1, l2, l3;
        umma::split_tf32(a, h0, l0); umma::split_tf32(b, h1, l1); umma::split_tf32(c, h2, l2); umma::split_tf32(d, h3, l3);
        *reinterpret_cast<float4*>(hi_plane + off) = make_float4(h0, h1, h2, h3);
        *reinterpret_cast<float4*>(lo_plane + off) = make_float4(l0, l1, l2, l3);
    } else {
        *reinterpret_cast<float4*>(hi_plane + off) = make_float4(round_tf32(a), round_tf32(b), round_tf32(c), round_tf32(d));
    }
}

// 16 consecutive features of this thread's sample -> TMEM A-operand planes (hi at col, lo at col + 64)
__device__ __forceinline__ void put_tmem16(uint32_t lane_base, uint32_t col, const float (&v)[16], bool split) {
    float t[16];
#pragma unroll
    for (uint32_t j = 0; j < 16; ++j) t[j] = split ? __uint_as_float(__float_as_uint(v[j]) & 0xffffe000u) : round_tf32(v[j]);
    umma::tmem_st16(lane_base + head::cAhi + col, t);
    if (split) {
#pragma unroll
        for (uint32_t j = 0; j < 16; ++j) t[j] = v[j] - __uint_as_float(__float_as_uint(v[j]) & 0xffffe000u);
        umma::tmem_st16(lane_base + head::cAlo + col, t);
    }
}

// D[64, N] (+)= At^T . Bt : MN-major swizzled tiles whose 128 rows are the contraction index (low descriptor halves)
template <uint32_t N>
__device__ __forceinline__ void issue_gemm_mn(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                              bool split, uint32_t acc) {
    constexpr uint32_t idesc = umma::idesc_tf32(64, N, 1, 1);
    using namespace head;
    constexpr uint32_t hi = umma::desc_hi(kMnSbo, umma::kLayoutMn32);
    constexpr uint32_t step = (2u * kMnSbo) >> 4;                 // 8 contraction rows = two 512-byte atoms
#pragma unroll
    for (uint32_t ks = 0; ks < kTile / 8; ++ks) {
        const uint32_t o = ks * step;
        if (split) {
            umma::mma_tf32_ss2(tmem_d, a_lo + o, hi, b_hi + o, hi, idesc, acc);
            umma::mma_tf32_ss2(tmem_d, a_hi + o, hi, b_lo + o, hi, idesc, 1u);
            acc = 1;
        }
        umma::mma_tf32_ss2(tmem_d, a_hi + o, hi, b_hi + o, hi, idesc, acc);
        acc = 1;
    }
}

__device__ __forceinline__ float4 ldg_nc_alloc_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_plain_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_ef_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::evict_first.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// stage the TRANSPOSE of an nn.Linear weight (as stage_weight_t) with all of a thread's loads in flight before its first store
template <uint32_t ROWS, uint32_t COLS, uint32_t NTHREADS>
__device__ __forceinline__ void stage_weight_t_batched(const float* __restrict__ w, uint8_t* plane_hi, bool split, uint32_t tid) {
    constexpr uint32_t kPer = (ROWS * COLS + NTHREADS - 1) / NTHREADS;
    uint8_t* plane_lo = plane_hi + ROWS * COLS * 4;
    float v[kPer];
#pragma unroll
    for (uint32_t j = 0; j < kPer; ++j) {
        const uint32_t i = tid + j * NTHREADS;
        v[j] = (i < ROWS * COLS) ? __ldg(w + i) : 0.0f;
    }
#pragma unroll
    for (uint32_t j = 0; j < kPer; ++j) {
        const uint32_t i = tid + j * NTHREADS;
        if (i < ROWS * COLS) {
            const uint32_t r = i / COLS, c = i - r * COLS;          // W[r][c] -> tile row c, contraction index r
            const uint32_t off = umma::tile_off(COLS, c, r);
            if (split) {
                float hi, lo;
                umma::split_tf32(v[j], hi, lo);
                *reinterpret_cast<float*>(plane_hi + off) = hi;
                *reinterpret_cast<float*>(plane_lo + off) = lo;
            } else {
                *reinterpret_cast<float*>(plane_hi + off) = round_tf32(v[j]);
            }
        }
    }
}
// 1-D bulk copy global -> shared memory (TMA, no tensor map); completion = transaction bytes on an mbarrier
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_saddr, const void* src, uint32_t bytes, uint32_t bar_saddr) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_saddr), "l"(src), "r"(bytes), "r"(bar_saddr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar_saddr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_saddr), "r"(bytes) : "memory");
}
// 256-bit read-only load (two consecutive float4)
__device__ __forceinline__ void ldg_nc_f8(const float* p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(gptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ float4 ldg_nc_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// Hash-grid scatter of ONE level for this thread's sample, fused into the backward (replaces the matching slice of
// sanerf_grid_encode_backward): the two gradient values come straight from tensor memory, consecutive lanes that share
// a cell are merged first (grid_common.cuh: warp_run_reduce), the run heads issue the eight red.global.add.v2.f32.
__device__ __noinline__ void scatter_level(const float* __restrict__ x01, float* __restrict__ g_table, const LevelGeom<3> geo,
                                           uint32_t base_row, uint32_t taddr, uint32_t b, uint32_t B, uint32_t lane) {
    float g0, g1;
    umma::tmem_ld2(taddr, g0, g1);
    float x[3] = {0.5f, 0.5f, 0.5f};
    const bool live = b < B;
    if (live) {
#pragma unroll
        for (uint32_t d = 0; d < 3; ++d) x[d] = __ldg(x01 + (size_t)b * 3 + d);
    }
    const bool contributes = live && !out_of_range<3>(x) && (g0 != 0.0f || g1 != 0.0f);
    const Cell<3> cell = locate<3>(geo, x, false, 0u);
    float v[16];
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) {
        const float w = contributes ? corner_weight<3>(cell, k) : 0.0f;
        v[2 * k] = w * g0;
        v[2 * k + 1] = w * g1;
    }
    const uint32_t key[3] = {contributes ? cell.lo[0] : 0xffffffffu - lane, cell.lo[1], cell.lo[2]};
    const bool head_lane = warp_run_reduce<16, 3>(v, key, lane);
    if (head_lane && contributes) {
        float* slice = g_table + (size_t)base_row * 2;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) red_add_v2_f32(slice + (size_t)corner_row<3>(geo, cell, k) * 2, v[2 * k], v[2 * k + 1]);
    }
}

// 16 symmetric worker warps.  Warp w serves TMEM lane quadrant q = w & 3 (tile rows 32 q .. 32 q + 31) and, in the
// epilogues, the 16 accumulator columns 16 (w >> 2) ..; in the staging phases it moves the 8 tile rows 8 w .. 8 w + 7.
// Every phase therefore runs with four warps per scheduler instead of one.
__global__ void __launch_bounds__(head::kBwdThreads, 1) head_backward_lockstep_kernel(const HeadBwdParams p, const uint32_t tiles) {
    pdl_begin();
    using namespace head;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_mma;
    __shared__ __align__(16) uint8_t s_mask[2][kTile][16];     // ReLU sign patterns of H1 / H2: one nibble per 4-feature chunk
    __shared__ uint32_t s_tmem;
    __shared__ LevelGeom<3> s_geo[kLevels];
    __shared__ uint32_t s_base[kLevels];
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool split = (p.precision == 0);
    const uint32_t q = warp & 3u, cg = warp >> 2;           // TMEM lane quadrant, 16-column group
    const bool scatter = (p.x01 != nullptr);
    if (scatter && tid >= 64 && tid < 64 + kLevels) {
        s_geo[tid - 64] = level_geometry<3>(p.offsets, tid - 64, p.S, p.H, 0u);
        s_base[tid - 64] = (uint32_t)__ldg(p.offsets + (tid - 64));
    }
    const uint32_t row = q * 32u + lane;                    // tile row of this thread's TMEM lane
    uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);

    if (warp == 0) umma::tmem_alloc<kBwdTmemCols>(umma::smem_u32(&s_tmem));
    if (tid == 32) {
        umma::mbar_init(umma::smem_u32(&s_mma), 1);
        umma::fence_mbar_init();
    }
    stage_weight_t(p.w3, smem + oT3, kOut, kHid, split, tid, kBwdThreads);
    stage_weight_t(p.w2, smem + oT2, kHid, kHid, split, tid, kBwdThreads);
    stage_weight_t(p.w1, smem + oT1, kHid, kIn, split, tid, kBwdThreads);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s_tmem;
    const uint32_t lane_base = umma::tmem_addr(tmem, q * 32u, 0);
    // low descriptor halves: transposed weights K-major (LBO = rows * 16), tiles MN-major (LBO = next 32-wide block)
    const uint32_t dT3h = umma::desc_lo(umma::smem_u32(smem + oT3), kHid * 16u), dT3l = dT3h + (kW3 >> 4);
    const uint32_t dT2h = umma::desc_lo(umma::smem_u32(smem + oT2), kHid * 16u), dT2l = dT2h + (kW2 >> 4);
    const uint32_t dT1h = umma::desc_lo(umma::smem_u32(smem + oT1), kIn * 16u), dT1l = dT1h + (kW1 >> 4);
    const uint32_t dA = umma::desc_lo(umma::smem_u32(smem + oBufA), kMnLbo), dB = umma::desc_lo(umma::smem_u32(smem + oBufB), kMnLbo);
    const uint32_t dG = umma::desc_lo(umma::smem_u32(smem + oBufG), kMnLbo);
    const bool issuer = (warp == kBwdThreads / 32 - 1);
    uint8_t* bufA = smem + oBufA;
    uint8_t* bufB = smem + oBufB;
    uint8_t* bufG = smem + oBufG;
    uint32_t phase = 0;

    auto mma_done = [&]() {
        umma::mbar_wait(umma::smem_u32(&s_mma), phase);
        phase ^= 1u;
        umma::fence_after_sync();
    };
    auto publish = [&]() {       // tiles written (shared: generic -> async proxy; tensor memory: st complete), TMEM reads done
        umma::tmem_st_wait();
        umma::fence_proxy_async();
        umma::fence_before_sync();
        __syncthreads();
    };

    // ---- staging registers of this thread: 4 chunks of H2, 4 of H1, 2 of enc, 1 of g_out (any-lane pieces of the tile)
    // plus, in the warps with cg == 0, this lane's own g_out row for the TMEM A planes.  Loaded one tile ahead.
    float4 rh2[4], rh1[4], re[2], rg, rgrow[4];
    // tile-chunk-major pieces: one instruction = one chunk of 32 consecutive rows (512 contiguous bytes)
    const uint32_t h_c = warp;                                             // H1 / H2: chunk = warp, rows 32 i + lane
    const uint32_t e_c = warp >> 1, e_row = (warp & 1u) * 64u + lane;      // enc: chunk = warp / 2, rows e_row + 32 i
    const uint32_t g_row = warp * 8u + (lane >> 2), g_c = lane & 3u;       // g_out (row-major): 8 rows per instruction
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ld_tcm = [&](const float* base, uint32_t chunks, uint32_t tile, uint32_t r, uint32_t c) -> float4 {
        const uint32_t bb = tile * kTile + r;
        return (tile < tiles && bb < p.B) ? ldg_nc_f4(base + tcm_off(bb, c, chunks)) : zero4;
    };
    auto ld_row = [&](const float* base, uint32_t width, uint32_t tile, uint32_t r, uint32_t c) -> float4 {
        const uint32_t bb = tile * kTile + r;
        return (tile < tiles && bb < p.B) ? ldg_nc_f4(base + (size_t)bb * width + 4u * c) : zero4;
    };
    auto put_mask = [&](uint32_t layer, uint32_t r, uint32_t c, const float4& v) {
        s_mask[layer][r][c] = (uint8_t)((v.x > 0.f) | ((v.y > 0.f) << 1) | ((v.z > 0.f) << 2) | ((v.w > 0.f) << 3));
    };
    auto load_h2_g = [&](uint32_t tile) {
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) rh2[i] = ld_tcm(p.h2, 16, tile, 32 * i + lane, h_c);
        rg = ld_row(p.g_out, kOut, tile, g_row, g_c);
        if (cg == 0) {
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) rgrow[j] = ld_row(p.g_out, kOut, tile, row, j);
        }
    };
    auto load_h1 = [&](uint32_t tile) {
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) rh1[i] = ld_tcm(p.h1, 16, tile, 32 * i + lane, h_c);
    };
    auto load_enc = [&](uint32_t tile) {
#pragma unroll
        for (uint32_t i = 0; i < 2; ++i) re[i] = ld_tcm(p.enc, 8, tile, e_row + 32 * i, e_c);
    };
    // masked data gradient: accumulator columns 16 cg .. of this thread's row -> TMEM A planes + MN-major tile
    auto masked_epilogue = [&](uint32_t layer) {
        const uint32_t c0 = cg * 16u;
        const uint32_t nib = *reinterpret_cast<const uint32_t*>(&s_mask[layer][row][cg * 4u]);      // 4 chunks = 16 columns
        const uint32_t bits = (nib & 0xfu) | ((nib >> 4) & 0xf0u) | ((nib >> 8) & 0xf00u) | ((nib >> 12) & 0xf000u);
        float v[16];
        umma::tmem_ld16(lane_base + cDG + c0, v);
#pragma unroll
        for (uint32_t j = 0; j < 16; ++j) v[j] = ((bits >> j) & 1u) ? v[j] : 0.0f;
        put_tmem16(lane_base, c0, v, split);
        if (!HEAD_DBG(4))
#pragma unroll
        for (uint32_t j = 0; j < 16; j += 4) put_mn(bufB, bufB + kMn64, row, c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3], split);
    };

    // One level (of the four this warp serves: 4 cg + lq) of the PREVIOUS tile's scatter.  Spread over the tile's four
    // tensor-core waits: the L2 reduction rate (~0.7 per clock and SM), not the issue rate, bounds the scatter, so the
    // reductions have to drain while the tensor core and the other phases are busy, not in one burst.
    auto scatter_slot = [&](uint32_t prev_it, uint32_t prev_tile, uint32_t lq) {
        const uint32_t level = cg * 4u + lq;
        scatter_level(p.x01, p.g_table, s_geo[level], s_base[level],
                      lane_base + cDGE + (prev_it & 1u) * 32u + cg * 8u + lq * 2u, prev_tile * kTile + row, p.B, lane);
    };

    load_h2_g(blockIdx.x);
    load_h1(blockIdx.x);
    load_enc(blockIdx.x);

    for (uint32_t it = 0, tile = blockIdx.x; tile < tiles; ++it, tile += gridDim.x) {
        const uint32_t first = (it == 0) ? 0u : 1u;
        const uint32_t next = tile + gridDim.x;
        HEAD_TRACE(0);
        // ---- stage what step 3 needs: H2 -> bufB (+ sign mask), G3 -> bufG and TMEM A planes
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) {
            put_mask(1, 32 * i + lane, h_c, rh2[i]);
            put_mn(bufB, bufB + kMn64, 32 * i + lane, h_c * 4u, rh2[i].x, rh2[i].y, rh2[i].z, rh2[i].w, split);
        }
        put_mn(bufG, bufG + kMn32, g_row, g_c * 4u, rg.x, rg.y, rg.z, rg.w, split);
        if (cg == 0) {
            float g16[16];
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j) {
                g16[4 * j] = rgrow[j].x; g16[4 * j + 1] = rgrow[j].y; g16[4 * j + 2] = rgrow[j].z; g16[4 * j + 3] = rgrow[j].w;
            }
            put_tmem16(lane_base, 0, g16, split);
        }
        HEAD_TRACE(1);
        publish();
        HEAD_TRACE(2);
        // ---- 3
        if (issuer) {
            if (umma::elect_one()) {
                umma::fence_after_sync();
                if (!HEAD_DBG(16)) issue_gemm_ts<kHid, kOut>(tmem + cDG, tmem + cAhi, tmem + cAlo, dT3h, dT3l, split);
                if (!HEAD_DBG(8)) issue_gemm_mn<kOut>(tmem + cW3, dB, dB + (kMn64 >> 4), dG, dG + (kMn32 >> 4), split, first);
                umma::commit(umma::smem_u32(&s_mma));
            }
            __syncwarp();
        }
        // H1 -> bufA is needed by step 4 only: written while step 3 runs on the tensor core; the freed registers are
        // refilled with the next tile's rows
        if (!HEAD_DBG(2))
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) {
            put_mask(0, 32 * i + lane, h_c, rh1[i]);
            put_mn(bufA, bufA + kMn64, 32 * i + lane, h_c * 4u, rh1[i].x, rh1[i].y, rh1[i].z, rh1[i].w, split);
        }
        if (!HEAD_DBG(1)) load_h2_g(next);
        if (scatter && it > 0) scatter_slot(it - 1, tile - gridDim.x, 0);
        mma_done();
        HEAD_TRACE(3);
        masked_epilogue(1);                                  // G2 = DG2 * [H2 > 0]
        HEAD_TRACE(4);
        publish();
        HEAD_TRACE(5);
        // ---- 4
        if (issuer) {
            if (umma::elect_one()) {
                umma::fence_after_sync();
                if (!HEAD_DBG(16)) issue_gemm_ts<kHid, kHid>(tmem + cDG, tmem + cAhi, tmem + cAlo, dT2h, dT2l, split);
                if (!HEAD_DBG(8)) issue_gemm_mn<kHid>(tmem + cW2, dB, dB + (kMn64 >> 4), dA, dA + (kMn64 >> 4), split, first);
                umma::commit(umma::smem_u32(&s_mma));
            }
            __syncwarp();
        }
        if (!HEAD_DBG(1)) load_h1(next);
        if (scatter && it > 0) { scatter_slot(it - 1, tile - gridDim.x, 1); scatter_slot(it - 1, tile - gridDim.x, 2); }
        mma_done();
        HEAD_TRACE(6);
        masked_epilogue(0);                                  // G1 = DG1 * [H1 > 0]; step 4 has released bufA and bufB
#pragma unroll
        for (uint32_t i = 0; i < 2; ++i)
            put_mn(bufA, bufA + kMn32, e_row + 32 * i, e_c * 4u, re[i].x, re[i].y, re[i].z, re[i].w, split);
        HEAD_TRACE(7);
        publish();
        HEAD_TRACE(8);
        // ---- 5
        if (issuer) {
            if (umma::elect_one()) {
                umma::fence_after_sync();
                if (!HEAD_DBG(16)) issue_gemm_ts<kIn, kHid>(tmem + (scatter ? cDGE + (it & 1u) * 32u : cDG), tmem + cAhi, tmem + cAlo, dT1h, dT1l, split);
                if (!HEAD_DBG(8)) issue_gemm_mn<kIn>(tmem + cW1, dB, dB + (kMn64 >> 4), dA, dA + (kMn32 >> 4), split, first);
                umma::commit(umma::smem_u32(&s_mma));
            }
            __syncwarp();
        }
        if (!HEAD_DBG(1)) load_enc(next);
        if (scatter && it > 0) scatter_slot(it - 1, tile - gridDim.x, 3);
        mma_done();
        HEAD_TRACE(9);
        if (scatter) {
            // the gradient of the encoding stays in tensor memory (buffer it & 1); it is scattered during the next tile
        } else if (cg < 2) {                                 // g_enc [B,32] row-major: 64 contiguous bytes per thread
            const uint32_t b = tile * kTile + row;
            float v[16];
            umma::tmem_ld16(lane_base + cDG + cg * 16u, v);
            if (b < p.B) {
                float4* dst = reinterpret_cast<float4*>(p.g_enc + (size_t)b * kIn + cg * 16u);
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        }
        HEAD_TRACE(10);
        // the next tile's staging may start at once: every MMA that read bufA / bufB / bufG / the A planes has completed;
        // this tile's reads of cDG are ordered before the next MMAs by the fence in the next publish()
    }
    if (scatter) {                                        // the last tile's scatter
        const uint32_t n_it = (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
        if (n_it > 0) {
            const uint32_t last_tile = blockIdx.x + (n_it - 1) * gridDim.x;
#pragma unroll 1
            for (uint32_t lq = 0; lq < 4; ++lq) scatter_slot(n_it - 1, last_tile, lq);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();

    // ---- weight gradients: M = 64 accumulators, row (16 q + l) lives in TMEM lane (32 q + l), l < 16
    {
        const uint32_t wrow = q * 16 + lane;
        const bool owns = lane < 16;
        float v[16];
        umma::tmem_ld16(lane_base + cW2 + cg * 16u, v);      // dW2[out=wrow][in = 16 cg ..]
        if (owns) {
#pragma unroll
            for (uint32_t j = 0; j < 16; j += 4) red_add_v4_f32(p.g_w2 + wrow * kHid + cg * 16u + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (cg < 2) {                                        // dW1[out=wrow][in = 16 cg ..]
            umma::tmem_ld16(lane_base + cW1 + cg * 16u, v);
            if (owns) {
#pragma unroll
                for (uint32_t j = 0; j < 16; j += 4) red_add_v4_f32(p.g_w1 + wrow * kIn + cg * 16u + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
        if (cg == 2) {                                       // dW3^T[in=wrow][out=c] -> g_w3[out][in]
            umma::tmem_ld16(lane_base + cW3, v);
            if (owns) {
#pragma unroll
                for (uint32_t j = 0; j < 16; ++j) red_add_f32(p.g_w3 + j * kHid + wrow, v[j]);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<head::kBwdTmemCols>(tmem);
}


// ---------------------------------------------------------------------------------------------------------------------
// The backward without the fused scatter (the form the training steps use).  Same operand placement as above, but
//  * the data-gradient chain (DG2 -> G2 -> DG1 -> G1 -> DGE) and the weight-gradient products complete on SEPARATE
//    mbarriers: an epilogue reads its accumulator as soon as the chain product is done and writes the next A operand
//    into tensor memory while the weight-gradient product of the previous stage still runs; the shared-memory copy of
//    the same gradient tile (the MN-major operand of the NEXT weight-gradient product) is written afterwards, beside
//    the next chain product.  Per 128-sample tile the tensor pipe sees DG2 | dW3 | DG1 | dW2 | DGE | dW1 back to back
//    instead of three issue -> drain -> epilogue rounds (measured phase times: profiles/r2b_head_backward.md);
//  * DG2 of a tile only needs G3 in tensor memory, so it is issued before the tile's activations are staged;
//  * every 16-byte store into an MN-major tile is bank-conflict-free: in the 128-byte-swizzle / 32-byte-atom layout the
//    32 rows a warp writes land in only four 32-byte slots of a 128-byte line, so lanes alternate (by bit 2 of the lane)
//    between the two 4-feature chunks that share such a slot - half of the lanes write the lower 16 bytes, half the upper.
__global__ void __launch_bounds__(head::kBwdThreads2, 1) head_backward_kernel(const HeadBwdParams p, const uint32_t tiles) {
    HEAD_MARK(0);
    pdl_begin();
    HEAD_MARK(1);
    using namespace head;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t s_chain, s_wgrad, s_tmem_ready, s_smem_ready, s_land;
    __shared__ uint32_t s_tmem;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool split = (p.precision == 0);
    const uint32_t q = warp & 3u, cg = warp >> 2;           // TMEM lane quadrant, 16-column group
    const uint32_t row = q * 32u + lane;                    // tile row of this thread's TMEM lane
    const uint32_t sw = (lane >> 2) & 1u;                   // which chunk of a pair this lane takes first (conflict-free stores)
    uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);

    if (warp == 0) umma::tmem_alloc<kBwdTmemCols>(umma::smem_u32(&s_tmem));
    if (tid == 32) {
        umma::mbar_init(umma::smem_u32(&s_chain), 1);
        umma::mbar_init(umma::smem_u32(&s_wgrad), 1);
        umma::mbar_init(umma::smem_u32(&s_tmem_ready), 1);
        umma::mbar_init(umma::smem_u32(&s_smem_ready), 1);
        umma::mbar_init(umma::smem_u32(&s_land), 1);
        umma::fence_mbar_init();
    }
    stage_weight_t_batched<kOut, kHid, kBwdThreads2>(p.w3, smem + oT3, split, tid);
    stage_weight_t_batched<kHid, kHid, kBwdThreads2>(p.w2, smem + oT2, split, tid);
    stage_weight_t_batched<kHid, kIn, kBwdThreads2>(p.w1, smem + oT1, split, tid);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    HEAD_MARK(2);
    const uint32_t tmem = s_tmem;
    const uint32_t lane_base = umma::tmem_addr(tmem, q * 32u, 0);
    const uint32_t dT3h = umma::desc_lo(umma::smem_u32(smem + oT3), kHid * 16u), dT3l = dT3h + (kW3 >> 4);
    const uint32_t dT2h = umma::desc_lo(umma::smem_u32(smem + oT2), kHid * 16u), dT2l = dT2h + (kW2 >> 4);
    const uint32_t dT1h = umma::desc_lo(umma::smem_u32(smem + oT1), kIn * 16u), dT1l = dT1h + (kW1 >> 4);
    const uint32_t dA = umma::desc_lo(umma::smem_u32(smem + oBufA), kMnLbo), dB = umma::desc_lo(umma::smem_u32(smem + oBufB), kMnLbo);
    const uint32_t dG = umma::desc_lo(umma::smem_u32(smem + oBufG), kMnLbo);
    uint8_t* bufA = smem + oBufA;
    uint8_t* bufB = smem + oBufB;
    uint8_t* bufG = smem + oBufG;
    const uint32_t bar_chain = umma::smem_u32(&s_chain), bar_wgrad = umma::smem_u32(&s_wgrad);
    const uint32_t bar_tmem_ready = umma::smem_u32(&s_tmem_ready), bar_smem_ready = umma::smem_u32(&s_smem_ready);
    uint32_t ph_chain = 0, ph_wgrad = 0;

    if (warp == kBwdThreads / 32) {
        // ===================== issuing warp: six products per tile, each as soon as its operands are published ==========
        // (the ~200 tcgen05.mma of a tile take 3-4 thousand clocks to ISSUE; on a worker warp that time sat on every
        // worker's critical path through the next barrier)
        uint32_t ph_t = 0, ph_s = 0;
        auto ready = [&](uint32_t bar, uint32_t& ph) { umma::mbar_wait(bar, ph); ph ^= 1u; };
        // The workers' staging loads go through an L1 that the 222 KB of tiles leave ~30 KB of: the lines it can keep in
        // flight times the HBM latency bound the loads (measured: 14 k -> 9 k clocks per tile without them).  This warp
        // therefore pulls the tile after next into L2 with bulk prefetches (no registers, no shared memory), so the
        // loads meet the L2 latency instead.
        auto prefetch_tile = [&](uint32_t t) {
            if (t >= tiles || HEAD_DBG(2)) return;
            const uint32_t rows = (p.B - t * kTile) < kTile ? (p.B - t * kTile) : kTile;
            bulk_prefetch_l2(p.h2 + (size_t)t * kTile * kHid, kTile * kHid * 4u);
            bulk_prefetch_l2(p.h1 + (size_t)t * kTile * kHid, kTile * kHid * 4u);
            bulk_prefetch_l2(p.enc + (size_t)t * kTile * kIn, kTile * kIn * 4u);
            bulk_prefetch_l2(p.g_out + (size_t)t * kTile * kOut, rows * kOut * 4u);
        };
        if (umma::elect_one()) { prefetch_tile(blockIdx.x); prefetch_tile(blockIdx.x + gridDim.x); }
        __syncwarp();
        for (uint32_t it = 0, tile = blockIdx.x; tile < tiles; ++it, tile += gridDim.x) {
            const uint32_t first = (it == 0) ? 0u : 1u;
            if (umma::elect_one()) prefetch_tile(tile + 2u * gridDim.x);
            __syncwarp();
            ready(bar_tmem_ready, ph_t);
            if (umma::elect_one()) {
                umma::fence_after_sync();
                issue_gemm_ts<kHid, kOut>(tmem + cDG, tmem + cAhi, tmem + cAlo, dT3h, dT3l, split);              // DG2 = G3 W3
                umma::commit(bar_chain);
            }
            __syncwarp();
            ready(bar_smem_ready, ph_s);
            if (umma::elect_one()) {
                umma::fence_proxy_async();
                umma::fence_after_sync();
                issue_gemm_mn<kOut>(tmem + cW3, dB, dB + (kMn64 >> 4), dG, dG + (kMn32 >> 4), split, first);     // dW3^T += H2^T G3
                umma::commit(bar_wgrad);
            }
            __syncwarp();
            ready(bar_tmem_ready, ph_t);
            if (umma::elect_one()) {
                umma::fence_after_sync();
                issue_gemm_ts<kHid, kHid>(tmem + cDG, tmem + cAhi, tmem + cAlo, dT2h, dT2l, split);              // DG1 = G2 W2
                umma::commit(bar_chain);
            }
            __syncwarp();
            ready(bar_smem_ready, ph_s);
            if (umma::elect_one()) {
                umma::fence_proxy_async();
                umma::fence_after_sync();
                issue_gemm_mn<kHid>(tmem + cW2, dB, dB + (kMn64 >> 4), dA, dA + (kMn64 >> 4), split, first);     // dW2 += G2^T H1
                umma::commit(bar_wgrad);
            }
            __syncwarp();
            ready(bar_tmem_ready, ph_t);
            if (umma::elect_one()) {
                umma::fence_after_sync();
                issue_gemm_ts<kIn, kHid>(tmem + cDG, tmem + cAhi, tmem + cAlo, dT1h, dT1l, split);               // DGE = G1 W1
                umma::commit(bar_chain);
            }
            __syncwarp();
            ready(bar_smem_ready, ph_s);
            if (umma::elect_one()) {
                umma::fence_proxy_async();
                umma::fence_after_sync();
                issue_gemm_mn<kIn>(tmem + cW1, dB, dB + (kMn64 >> 4), dA, dA + (kMn32 >> 4), split, first);      // dW1 += G1^T enc
                umma::commit(bar_wgrad);
            }
            __syncwarp();
        }
    } else {
    // ===================== 16 worker warps =====================

    auto wait_chain = [&]() { umma::mbar_wait(bar_chain, ph_chain); ph_chain ^= 1u; umma::fence_after_sync(); };
    auto wait_wgrad = [&]() { umma::mbar_wait(bar_wgrad, ph_wgrad); ph_wgrad ^= 1u; umma::fence_after_sync(); };
    // A planes written (tcgen05.st complete) and this thread's reads of the accumulator done
    auto publish_tmem = [&]() {
        umma::tmem_st_wait(); umma::fence_before_sync(); named_bar_sync(1, kBwdThreads);
        if (tid == 0) mbar_arrive(bar_tmem_ready);
    };
    // MN-major tiles written: generic-proxy stores -> visible to the tensor core
    // The generic-proxy -> async-proxy fence for the tiles is executed by the ISSUING thread, after it has observed the
    // workers' barrier (a proxy fence anywhere on the causality path from the stores to the tensor-core reads orders them).
    // On the writers' side fence.proxy.async is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the MEMBAR waits for every global
    // load the thread has in flight: with the staging registers loaded one phase ahead that exposed the whole HBM latency
    // three times per tile (measured: 14.1 k -> 10.0 k clocks per tile without the loads, profiles/r2b_head_backward.md).
    auto publish_smem = [&]() {
        umma::fence_before_sync(); named_bar_sync(1, kBwdThreads);
        if (tid == 0) mbar_arrive(bar_smem_ready);
    };

    // ---- staging registers, loaded two phases before they are staged (tile-chunk-major pieces; one instruction = two runs of 16 rows x 16 B
    // of the two chunks of a pair: lanes alternate between the chunks in groups of four)
    float4 rh[4], re[2], rg, rgq;     // rh: H2 of the next tile, then H1 of the current one
    const uint32_t h_cp = warp >> 1, h_r0 = (warp & 1u) * 64u + lane;      // H1 / H2: chunk pair, rows h_r0 + 32 i
    const uint32_t e_cp = warp >> 2, e_row = (warp & 3u) * 32u + lane;     // enc: chunk pair, one row per lane
    // g_out (row-major): 8 rows per instruction; rows 1 <-> 2 and 5 <-> 6 trade places so that the two rows a quarter-warp
    // stores differ by 2 mod 4 (conflict-free in the swizzled tile)
    const uint32_t g_i = lane >> 2, g_row = warp * 8u + ((g_i & 4u) | ((g_i & 1u) << 1) | ((g_i >> 1) & 1u)), g_c = lane & 3u;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ld_tcm = [&](const float* base, uint32_t chunks, uint32_t tile, uint32_t r, uint32_t c) -> float4 {
        const uint32_t bb = tile * kTile + r;
#ifdef SANERF_HEAD_TRACE
        if (!(tile < tiles && bb < p.B)) return zero4;
        const float* src = base + tcm_off(bb, c, chunks);
        return HEAD_DBG(4) ? ldg_nc_alloc_f4(src) : HEAD_DBG(8) ? ldg_plain_f4(src) : HEAD_DBG(16) ? ldg_ef_f4(src) : ldg_nc_f4(src);
#else
        return (tile < tiles && bb < p.B) ? ldg_nc_alloc_f4(base + tcm_off(bb, c, chunks)) : zero4;
#endif
    };
    auto ld_row = [&](const float* base, uint32_t width, uint32_t tile, uint32_t r, uint32_t c) -> float4 {
        const uint32_t bb = tile * kTile + r;
        return (tile < tiles && bb < p.B) ? ldg_nc_f4(base + (size_t)bb * width + 4u * c) : zero4;
    };
    // an activation chunk as saved by the forward -> MN-major planes.  The hi plane takes the RAW values (the tensor core
    // ignores the 13 low mantissa bits of a tf32 operand itself: bit-identical products, see gemm_tma.cu), so the epilogue
    // reads the ReLU signs of its 16 columns straight from that plane - no separate sign table (whose byte stores were
    // 8-way bank-conflicted: 1.8 M excessive wavefronts per launch, profiles/r2b_head_backward.md)
    auto put_act = [&](uint8_t* hi_plane, uint8_t* lo_plane, uint32_t kk, uint32_t f, const float4& x) {
        const uint32_t off = umma::mn32_off(kTile, kk, f);
        if (split) {
            float h0, h1, h2, h3, l0, l1, l2, l3;
            umma::split_tf32(x.x, h0, l0); umma::split_tf32(x.y, h1, l1); umma::split_tf32(x.z, h2, l2); umma::split_tf32(x.w, h3, l3);
            *reinterpret_cast<float4*>(hi_plane + off) = x;
            *reinterpret_cast<float4*>(lo_plane + off) = make_float4(l0, l1, l2, l3);
        } else {
            *reinterpret_cast<float4*>(hi_plane + off) = make_float4(round_tf32(x.x), round_tf32(x.y), round_tf32(x.z), round_tf32(x.w));
        }
    };
    // H2 / H1 do not come through the L1 (which 217 KB of tiles leave ~30 KB of): thread 0 has the raw 32 KB tile-chunk-major
    // tile LANDED by a bulk copy in the one region that is idle at that time - the lo plane of bufA (H1's), free from the end
    // of dW2 of one tile to the H1 staging of the next - H2 of the next tile during E4b / E5, then H1 of the tile between its
    // H2 staging and its own staging; the workers read their pieces from there (conflict-free LDS.128).
    uint8_t* landing = bufA + kMn64;
    const uint32_t bar_land = umma::smem_u32(&s_land);
    uint32_t ph_land = 0;
    auto land = [&](const float* base, uint32_t tile) {          // thread 0, after a barrier behind every reader of the region
        umma::fence_proxy_async();                               // earlier generic-proxy accesses of the region -> async-proxy write
        mbar_arrive_expect_tx(bar_land, kTile * kHid * 4u);
        bulk_copy_g2s(umma::smem_u32(landing), base + (size_t)tile * kTile * kHid, kTile * kHid * 4u, bar_land);
    };
    auto load_h = [&](uint32_t tile) {                           // rh[2 i + t] = row h_r0 + 32 i, chunk 2 h_cp + (t ^ sw)
        umma::mbar_wait(bar_land, ph_land); ph_land ^= 1u;
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) {
            const uint32_t rr = h_r0 + 32u * (i >> 1), c = 2u * h_cp + ((i & 1u) ^ sw);
            rh[i] = *reinterpret_cast<const float4*>(landing + (c * kTile + rr) * 16u);
            if (tile * kTile + rr >= p.B) rh[i] = zero4;         // rows past B of the last tile are not initialised
        }
    };
    auto load_g = [&](uint32_t tile) {
        rg = ld_row(p.g_out, kOut, tile, g_row, g_c);        // for the MN-major tile: 8 rows per instruction
        rgq = ld_row(p.g_out, kOut, tile, row, cg);          // for the A planes: this lane's row, columns 4 cg ..
    };
    auto load_enc = [&](uint32_t tile) {
#pragma unroll
        for (uint32_t i = 0; i < 2; ++i) re[i] = ld_tcm(p.enc, 8, tile, e_row, 2u * e_cp + (i ^ sw));
    };
    // a 64-wide activation tile -> MN-major hi (raw) / lo planes; lanes alternate between the two chunks of a pair
    auto stage_h = [&](uint8_t* buf) {
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) {
            const uint32_t rr = h_r0 + 32u * (i >> 1), c = 2u * h_cp + ((i & 1u) ^ sw);
            put_act(buf, buf + kMn64, rr, c * 4u, rh[i]);
        }
    };
    // masked data gradient, part a: accumulator columns 16 cg .. of this thread's row -> registers -> TMEM A planes
    float v[16];
    // `act` = hi plane of the layer's own activation tile (this row, these 16 columns: the ReLU derivative); the four
    // 16-byte reads of a lane go in the order j ^ sw, like its stores: bank-conflict-free
    auto epilogue_tmem = [&](const uint8_t* act) {
        const uint32_t c0 = cg * 16u;
        uint32_t bits = 0;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t ch = j ^ sw;
            const float4 a = *reinterpret_cast<const float4*>(act + umma::mn32_off(kTile, row, c0 + 4u * ch));
            bits |= (uint32_t)((a.x > 0.f) | ((a.y > 0.f) << 1) | ((a.z > 0.f) << 2) | ((a.w > 0.f) << 3)) << (4u * ch);
        }
        umma::tmem_ld16(lane_base + cDG + c0, v);
#pragma unroll
        for (uint32_t j = 0; j < 16; ++j) v[j] = ((bits >> j) & 1u) ? v[j] : 0.0f;
        put_tmem16(lane_base, c0, v, split);
    };
    // part b: the same 16 values -> MN-major tile bufB; instruction j of a lane carries chunk j ^ sw
    auto epilogue_smem = [&]() {
        const uint32_t c0 = cg * 16u;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            const uint32_t o = j ^ 1u;       // the other chunk of the pair
            const float a = sw ? v[4 * o] : v[4 * j], b = sw ? v[4 * o + 1] : v[4 * j + 1];
            const float c = sw ? v[4 * o + 2] : v[4 * j + 2], d = sw ? v[4 * o + 3] : v[4 * j + 3];
            put_mn(bufB, bufB + kMn64, row, c0 + 4u * (j ^ sw), a, b, c, d, split);
        }
    };

    if (tid == 0 && blockIdx.x < tiles) land(p.h2, blockIdx.x);
    load_g(blockIdx.x);
    load_enc(blockIdx.x);

    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < tiles; ++it, tile += gridDim.x) {
        const uint32_t next = tile + gridDim.x;
        HEAD_TRACE(0);
        // ---- G3 -> TMEM A planes; DG2 = G3 W3 starts before anything else of the tile is staged
        {
            float h0, h1, h2, h3, l0, l1, l2, l3;
            if (split) {
                umma::split_tf32(rgq.x, h0, l0); umma::split_tf32(rgq.y, h1, l1);
                umma::split_tf32(rgq.z, h2, l2); umma::split_tf32(rgq.w, h3, l3);
                umma::tmem_st4(lane_base + cAlo + cg * 4u, l0, l1, l2, l3);
            } else {
                h0 = round_tf32(rgq.x); h1 = round_tf32(rgq.y); h2 = round_tf32(rgq.z); h3 = round_tf32(rgq.w);
            }
            umma::tmem_st4(lane_base + cAhi + cg * 4u, h0, h1, h2, h3);
        }
        publish_tmem();
        HEAD_TRACE(1);
        // ---- H2 -> bufB, G3 -> bufG (the previous tile's dW1 has to be done with bufA / bufB first)
        if (it > 0) wait_wgrad();
        HEAD_TRACE(2);
        load_h(tile);                                        // H2, landed during the previous tile's last phases
        stage_h(bufB);
        put_mn(bufG, bufG + kMn32, g_row, g_c * 4u, rg.x, rg.y, rg.z, rg.w, split);
        publish_smem();                                      // (also: every worker has read the landing region)
        if (tid == 0) land(p.h1, tile);                      // consumed two phases further down
        HEAD_TRACE(3);
        wait_chain();                                        // DG2
        HEAD_TRACE(4);
        epilogue_tmem(bufB);                                    // G2 = DG2 * [H2 > 0]
        publish_tmem();
        HEAD_TRACE(5);
        wait_wgrad();                                        // dW3 has released bufB
        HEAD_TRACE(6);
        epilogue_smem();
        load_h(tile);                                        // H1: raw tile in the landing region = the lo plane it is about to fill
        named_bar_sync(2, kBwdThreads);                      // every worker holds its pieces before anyone overwrites the region
        stage_h(bufA);                                       // bufA was released by the previous tile's dW1
        publish_smem();
        HEAD_TRACE(7);
        wait_chain();                                        // DG1
        HEAD_TRACE(8);
        epilogue_tmem(bufA);                                    // G1 = DG1 * [H1 > 0]
        publish_tmem();
        HEAD_TRACE(9);
        wait_wgrad();                                        // dW2 has released bufA and bufB
        HEAD_TRACE(10);
        if (tid == 0 && next < tiles) land(p.h2, next);      // dW2 was the last reader of bufA's lo plane
        epilogue_smem();
#pragma unroll
        for (uint32_t i = 0; i < 2; ++i)
            put_mn(bufA, bufA + kMn32, e_row, (2u * e_cp + (i ^ sw)) * 4u, re[i].x, re[i].y, re[i].z, re[i].w, split);
        HEAD_TRACE(14);
        publish_smem();
        HEAD_TRACE(15);
        if (!HEAD_DBG(1)) { load_g(next); load_enc(next); }                   // next tile's G3 and encoding: a tile ahead
        HEAD_TRACE(11);
        wait_chain();                                        // DGE
        HEAD_TRACE(12);
        {                                                    // g_enc [B,32] row-major: one full 32-byte sector per thread
            const uint32_t b = tile * kTile + row;
            float e8[8];
            umma::tmem_ld8(lane_base + cDG + cg * 8u, e8);
            if (b < p.B) stg_f8(p.g_enc + (size_t)b * kIn + cg * 8u, e8);
        }
        HEAD_TRACE(13);
        // the next tile's G3 goes into the A planes at once: DGE, the last reader, has completed; this tile's reads of
        // cDG are ordered before the next DG2 by the fence in publish_tmem()
    }
    if (it > 0) wait_wgrad();                              // the last dW1: every product of this CTA has completed
    HEAD_MARK(3);

    // ---- weight gradients: M = 64 accumulators, row (16 q + l) lives in TMEM lane (32 q + l), l < 16
    {
        const uint32_t wrow = q * 16 + lane;
        const bool owns = lane < 16;
        float w[16];
        umma::tmem_ld16(lane_base + cW2 + cg * 16u, w);      // dW2[out=wrow][in = 16 cg ..]
        if (owns && it > 0) {
#pragma unroll
            for (uint32_t j = 0; j < 16; j += 4) red_add_v4_f32(p.g_w2 + wrow * kHid + cg * 16u + j, w[j], w[j + 1], w[j + 2], w[j + 3]);
        }
        if (cg < 2) {                                        // dW1[out=wrow][in = 16 cg ..]
            umma::tmem_ld16(lane_base + cW1 + cg * 16u, w);
            if (owns && it > 0) {
#pragma unroll
                for (uint32_t j = 0; j < 16; j += 4) red_add_v4_f32(p.g_w1 + wrow * kIn + cg * 16u + j, w[j], w[j + 1], w[j + 2], w[j + 3]);
            }
        }
        if (cg == 2) {                                       // dW3^T[in=wrow][out=c] -> g_w3[out][in]
            umma::tmem_ld16(lane_base + cW3, w);
            if (owns && it > 0) {
#pragma unroll
                for (uint32_t j = 0; j < 16; ++j) red_add_f32(p.g_w3 + j * kHid + wrow, w[j]);
            }
        }
    }
    HEAD_MARK(4);
    }   // worker warps
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc<head::kBwdTmemCols>(tmem);
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_field_head_forward(const float* x01, const float* table, const int32_t* offsets, float S, uint32_t H,
                                         const float* enc_in, const float* w1, const float* w2, const float* w3, uint32_t B,
                                         float* enc_out, float* h1_out, float* h2_out, float* out, int precision,
                                         void* stream) {
    if (B == 0) return SANERF_OK;
    if (x01 != nullptr) {
        SANERF_REQUIRE_PTR(table); SANERF_REQUIRE_PTR(offsets);
    } else {
        SANERF_REQUIRE_PTR(enc_in);
    }
    SANERF_REQUIRE_PTR(w1); SANERF_REQUIRE_PTR(w2); SANERF_REQUIRE_PTR(w3); SANERF_REQUIRE_PTR(out);
    if (precision != 0 && precision != 1) return fail(SANERF_ERR_INVALID_ARG, "field_head: precision 0 (3xTF32) or 1 (TF32)");
    if ((h1_out == nullptr) != (h2_out == nullptr)) return fail(SANERF_ERR_INVALID_ARG, "field_head: h1_out and h2_out go together");
    HeadFwdParams p{x01, table, offsets, enc_in, w1, w2, w3, enc_out, h1_out, h2_out, out, B, H, S, precision,
                    nullptr, nullptr, 0u, 0u, 0u};
    const uint32_t tiles = div_up(B, head::kTile);
    const uint32_t blocks = tiles < (uint32_t)kNumSMs ? tiles : (uint32_t)kNumSMs;
    cudaError_t e = cudaFuncSetAttribute(head_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, head::kFwdSmem);
    if (e != cudaSuccess) return fail(SANERF_ERR_CUDA, "field_head_forward: %s", cudaGetErrorString(e));
    SANERF_LAUNCH(head_forward_kernel, blocks, head::kThreads, head::kFwdSmem, static_cast<cudaStream_t>(stream), p, tiles);
    return check_launch("head_forward_kernel");
}

extern "C" int sanerf_field_head_forward_chunk(const float* x01, const float* table, const int32_t* offsets, float S, uint32_t H,
                                               const float* w1, const float* w2, const float* w3, float* out, int precision,
                                               const uint32_t* ray_list, const uint32_t* list_count, uint32_t max_rays,
                                               uint32_t chunk, uint32_t chunk_len, uint32_t T, void* stream) {
    if (max_rays == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(x01); SANERF_REQUIRE_PTR(table); SANERF_REQUIRE_PTR(offsets);
    SANERF_REQUIRE_PTR(w1); SANERF_REQUIRE_PTR(w2); SANERF_REQUIRE_PTR(w3); SANERF_REQUIRE_PTR(out);
    SANERF_REQUIRE_PTR(ray_list); SANERF_REQUIRE_PTR(list_count);
    if (precision != 0 && precision != 1) return fail(SANERF_ERR_INVALID_ARG, "field_head: precision 0 (3xTF32) or 1 (TF32)");
    if (chunk_len == 0 || T % chunk_len != 0 || (chunk + 1) * chunk_len > T)
        return fail(SANERF_ERR_INVALID_ARG, "field_head_forward_chunk: chunk_len must divide T and the chunk lie inside the ray");
    HeadFwdParams p{x01, table, offsets, nullptr, w1, w2, w3, nullptr, nullptr, nullptr, out, max_rays * T, H, S, precision,
                    ray_list, list_count, chunk, chunk_len, T};
    const uint32_t tiles = div_up(max_rays * chunk_len, head::kTile);           // upper bound: the kernel reads the live count
    const uint32_t blocks = tiles < (uint32_t)kNumSMs ? tiles : (uint32_t)kNumSMs;
    cudaError_t e = cudaFuncSetAttribute(head_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, head::kFwdSmem);
    if (e != cudaSuccess) return fail(SANERF_ERR_CUDA, "field_head_forward_chunk: %s", cudaGetErrorString(e));
    SANERF_LAUNCH(head_forward_kernel, blocks, head::kThreads, head::kFwdSmem, static_cast<cudaStream_t>(stream), p, tiles);
    return check_launch("head_forward_kernel");
}

#ifdef SANERF_HEAD_TRACE
extern "C" __attribute__((visibility("default"))) void sanerf_debug_head_flags(int flags) { g_head_dbg_host = flags; }
extern "C" __attribute__((visibility("default"))) int sanerf_debug_head_marks(long long* out) {
    return cudaMemcpyFromSymbol(out, g_head_marks, sizeof(long long) * 8) == cudaSuccess ? 0 : 1;
}
extern "C" __attribute__((visibility("default"))) int sanerf_debug_head_trace(long long* out) {
    return cudaMemcpyFromSymbol(out, g_head_trace, sizeof(long long) * 2 * 4 * 16) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int sanerf_field_head_backward(const float* enc, const float* h1, const float* h2, const float* g_out,
                                          const float* w1, const float* w2, const float* w3, uint32_t B, float* g_enc,
                                          const float* x01, const int32_t* offsets, float S, uint32_t H, float* g_table,
                                          float* g_w1, float* g_w2, float* g_w3, int precision, void* stream) {
    if (B == 0) return SANERF_OK;
    if (x01 != nullptr) {
        SANERF_REQUIRE_PTR(offsets); SANERF_REQUIRE_PTR(g_table);
    } else {
        SANERF_REQUIRE_PTR(g_enc);
    }
    SANERF_REQUIRE_PTR(enc); SANERF_REQUIRE_PTR(h1); SANERF_REQUIRE_PTR(h2); SANERF_REQUIRE_PTR(g_out);
    SANERF_REQUIRE_PTR(w1); SANERF_REQUIRE_PTR(w2); SANERF_REQUIRE_PTR(w3);
    SANERF_REQUIRE_PTR(g_w1); SANERF_REQUIRE_PTR(g_w2); SANERF_REQUIRE_PTR(g_w3);
    if (precision != 0 && precision != 1) return fail(SANERF_ERR_INVALID_ARG, "field_head: precision 0 (3xTF32) or 1 (TF32)");
#ifdef SANERF_HEAD_TRACE
    const int dbg = g_head_dbg_host;
#else
    const int dbg = 0;
#endif
    HeadBwdParams p{enc, h1, h2, g_out, w1, w2, w3, g_enc, x01, offsets, g_table, S, H, g_w1, g_w2, g_w3, B, precision, dbg};
    const uint32_t tiles = div_up(B, head::kTile);
    const uint32_t blocks = tiles < (uint32_t)kNumSMs ? tiles : (uint32_t)kNumSMs;
    // the fused-scatter form keeps the lock-step kernel (SANERF_HEAD_BWD_LOCKSTEP=1 selects it everywhere: A/B timing)
    static const bool lockstep = [] { const char* s = getenv("SANERF_HEAD_BWD_LOCKSTEP"); return s != nullptr && atoi(s) != 0; }();
    auto kernel = (x01 != nullptr || lockstep) ? head_backward_lockstep_kernel : head_backward_kernel;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, head::kBwdSmem);
    if (e != cudaSuccess) return fail(SANERF_ERR_CUDA, "field_head_backward: %s", cudaGetErrorString(e));
    const uint32_t threads = (x01 != nullptr || lockstep) ? head::kBwdThreads : head::kBwdThreads2;
    SANERF_LAUNCH(kernel, blocks, threads, head::kBwdSmem, static_cast<cudaStream_t>(stream), p, tiles);
    return check_launch("head_backward_kernel");
}
