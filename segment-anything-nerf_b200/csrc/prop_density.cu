// Proposal density: hash-grid encode -> 2-layer MLP -> trunc_exp, fused (sm_100a).
//
// Replaces, for the proposal levels, GridEncoder (L<=8, F=2) + MLP(in,16,1 no bias, ReLU) + trunc_exp of
// nerf/network.py:211-219, 248-252: the [B, L*C] encoding, the [B,16] hidden layer and the pre-activation never
// touch HBM — the forward reads 12 B and writes 4 B per sample (plus the gathers, which hit L2: the proposal
// tables are 3 MB).  The backward recomputes the forward, scatters the table gradient with vector reductions and
// reduces the weight gradients through shared memory (each thread of the CTA owns one entry of dW and sweeps the
// tile's samples), then issues one atomic per entry per CTA.
//
// MLP (proposal levels are 1.5 M samples x 352 flop = 0.5 GFLOP) stays on the FP32 pipe: too small and too
// narrow (10 -> 16 -> 1) for a tensor-core tile; the wide heads use csrc/mlp_tc.cu.
#include "grid_common.cuh"

namespace sanerf {

constexpr uint32_t kPropHidden = 16;
constexpr uint32_t kPropMaxIn = 16;
constexpr uint32_t kPropThreads = 256;

struct PropParams {
    const float* x01;        // [B,3]
    const float* table;      // [rows,2]
    const int32_t* offsets;  // [L+1]
    const float* w1;         // [16, IN]  (nn.Linear weight layout: [out, in])
    const float* w2;         // [1, 16]
    uint32_t B, L, H;
    float S;
};

// `geo` / `base`: the per-level geometry, computed once per CTA (shared memory; every lane reads the same words: broadcast).
// Evaluated per thread it was ~50 of the ~280 instructions a sample spends per level (exp2f, the stride loop, two dependent
// offset loads in front of the gathers) in a kernel that ncu shows issue-bound (issue slots 55 % busy at 4 warps per scheduler).
template <uint32_t L>
__device__ __forceinline__ void prop_encode(const PropParams& p, const LevelGeom<3>* __restrict__ s_geo, const uint32_t* __restrict__ s_base,
                                            const float (&x)[3], bool oob, float (&enc)[2 * L]) {
    const float* __restrict__ table = p.table;
    float2 val[L][8];
    float frac[L][3];
#pragma unroll
    for (uint32_t l = 0; l < L; ++l) {
        if (!oob) {
            const LevelGeom<3> geo = s_geo[l];
            const Cell<3> cell = locate<3>(geo, x, false, 0u);
            const size_t base = (size_t)s_base[l];
#pragma unroll
            for (uint32_t d = 0; d < 3; ++d) frac[l][d] = cell.f[d];
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k)   // plain 8-byte gathers: these 3 MB tables hit L1/L2; the kernel is issue-bound and
                                               // the pair-merged form (grid_common.cuh: gather_pair_f2) costs it 30 %
                val[l][k] = __ldg(reinterpret_cast<const float2*>(table + (base + corner_row<3>(geo, cell, k)) * 2));
        } else {
#pragma unroll
            for (uint32_t d = 0; d < 3; ++d) frac[l][d] = 0.0f;
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) val[l][k] = make_float2(0.0f, 0.0f);
        }
    }
#pragma unroll
    for (uint32_t l = 0; l < L; ++l) {
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            float w = 1.0f;
#pragma unroll
            for (uint32_t d = 0; d < 3; ++d) w *= (k & (1u << d)) ? frac[l][d] : (1.0f - frac[l][d]);
            a0 = __fmaf_rn(w, val[l][k].x, a0);
            a1 = __fmaf_rn(w, val[l][k].y, a1);
        }
        enc[2 * l] = a0;
        enc[2 * l + 1] = a1;
    }
}

template <uint32_t L>
__global__ void __launch_bounds__(kPropThreads) prop_forward_kernel(const PropParams p, float* __restrict__ sigma,
                                                                    float* __restrict__ enc_out) {
    pdl_begin();
    constexpr uint32_t IN = 2 * L, INP = (IN + 3u) & ~3u;          // weight rows padded to whole float4: 128-bit broadcast reads
    __shared__ __align__(16) float s_w1[kPropHidden * INP];
    __shared__ float s_w2[kPropHidden];
    __shared__ LevelGeom<3> s_geo[L];
    __shared__ uint32_t s_base[L];
    for (uint32_t i = threadIdx.x; i < kPropHidden * IN; i += blockDim.x) s_w1[(i / IN) * INP + (i % IN)] = __ldg(p.w1 + i);
    if (threadIdx.x < kPropHidden) s_w2[threadIdx.x] = __ldg(p.w2 + threadIdx.x);
    if (threadIdx.x >= 32 && threadIdx.x < 32 + L) {
        s_geo[threadIdx.x - 32] = level_geometry<3>(p.offsets, threadIdx.x - 32, p.S, p.H, 0u);
        s_base[threadIdx.x - 32] = (uint32_t)__ldg(p.offsets + (threadIdx.x - 32));
    }
    __syncthreads();
    // persistent over 256-sample tiles: the weights / geometry prologue (every thread waits for five threads' dependent offset
    // loads) is paid once per resident CTA instead of once per tile (ncu: 5.5 % of the warp samples sat at that barrier)
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < p.B; b += gridDim.x * blockDim.x) {
    float x[3];
#pragma unroll
    for (uint32_t d = 0; d < 3; ++d) x[d] = __ldg(p.x01 + (size_t)b * 3 + d);
    float enc[IN];
    prop_encode<L>(p, s_geo, s_base, x, out_of_range<3>(x), enc);
    if (enc_out != nullptr) {                // kept for the backward: 8 L bytes per sample instead of 8 L gathers
#pragma unroll
        for (uint32_t l = 0; l < L; ++l)
            *reinterpret_cast<float2*>(enc_out + (size_t)b * IN + 2 * l) = make_float2(enc[2 * l], enc[2 * l + 1]);
    }
    float pre = 0.0f;
#pragma unroll
    for (uint32_t j = 0; j < kPropHidden; ++j) {
        float wj[INP];
#pragma unroll
        for (uint32_t i = 0; i < INP; i += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(s_w1 + j * INP + i);
            wj[i] = w4.x; wj[i + 1] = w4.y; wj[i + 2] = w4.z; wj[i + 3] = w4.w;
        }
        float h = 0.0f;
#pragma unroll
        for (uint32_t i = 0; i < IN; ++i) h = __fmaf_rn(wj[i], enc[i], h);      // same order as before: bit-identical
        pre = __fmaf_rn(s_w2[j], fmaxf(h, 0.0f), pre);
    }
    sigma[b] = expf(pre);                    // trunc_exp forward (activation.py:10)
    }
}

// Backward.  One thread = one sample (persistent grid-stride over 256-sample tiles); everything is warp-synchronous,
// there is no CTA-wide barrier inside the loop:
//  1. recompute the MLP forward and the gradient of the encoding from the encoding the forward saved (SAVED; 8 L
//     bytes per sample) or, without it, from a full re-gather;
//  2. weight gradients: the warp's 32 samples are staged in a warp-private shared-memory tile and every lane
//     accumulates the L+1 entries of dW it owns (lane -> hidden unit j = lane % 16, input half = lane / 16) in
//     registers ACROSS tiles; one cross-warp reduction and one atomic per entry per CTA at the very end;
//  3. table gradient: per level, consecutive lanes in the same cell are merged first (warp_run_reduce), then the
//     head lanes issue one red.global.add.v2.f32 per corner.
constexpr uint32_t kPropRow = 36;                        // h[16] | enc half 0 (8) | enc half 1 (8) | dpre (4)

template <uint32_t L, bool SAVED>
__global__ void __launch_bounds__(kPropThreads, 2) prop_backward_kernel(const PropParams p, const float* __restrict__ enc_saved,
                                                                        const float* __restrict__ g_sigma,
                                                                        float* __restrict__ grad_table,
                                                                        float* __restrict__ grad_w1,
                                                                        float* __restrict__ grad_w2, uint32_t tiles) {
    pdl_begin();
    constexpr uint32_t IN = 2 * L;
    constexpr uint32_t kWarps = kPropThreads / 32;
    __shared__ float s_w1[kPropHidden * IN];
    __shared__ float s_w2[kPropHidden];
    __shared__ LevelGeom<3> s_geo[L];
    __shared__ uint32_t s_base[L];
    __shared__ __align__(16) float s_rows[kWarps * 32 * kPropRow];
    __shared__ uint32_t s_next;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid == 0) s_next = 0u;
    for (uint32_t i = tid; i < kPropHidden * IN; i += blockDim.x) s_w1[i] = __ldg(p.w1 + i);
    if (tid < kPropHidden) s_w2[tid] = __ldg(p.w2 + tid);
    if (tid >= 32 && tid < 32 + L) {
        s_geo[tid - 32] = level_geometry<3>(p.offsets, tid - 32, p.S, p.H, 0u);
        s_base[tid - 32] = (uint32_t)__ldg(p.offsets + (tid - 32));
    }
    __syncthreads();

    float* rows = s_rows + (size_t)warp * 32 * kPropRow;
    float* row = rows + (size_t)lane * kPropRow;
    const uint32_t own_j = lane & 15u, own_half = lane >> 4;
    float acc1[L], acc2 = 0.0f;                         // dW1[own_j][own_half*L + q] / W2[own_j] ; dW2[own_j] (half 0)
#pragma unroll
    for (uint32_t q = 0; q < L; ++q) acc1[q] = 0.0f;
    const float* __restrict__ table = p.table;

    // Work distribution inside the CTA is dynamic: its tiles (blockIdx.x, + gridDim.x, ...) are 8 slices of 32 samples each, and
    // every warp pulls the next slice from a shared counter.  Samples are ray-ordered (a tile = 2-4 whole rays) and the skip below
    // depends on the position along the ray: with the fixed assignment warp w <-> slice w the same warps of a CTA did all the work
    // while the others waited at the final barrier (ncu source page: 42-64 % of all warp samples were stall_barrier at the first
    // instruction behind it).
    const uint32_t my_tiles = (tiles > blockIdx.x) ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    const uint32_t my_slices = my_tiles * kWarps;
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = atomicAdd(&s_next, 1u);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= my_slices) break;
        const uint32_t tile = blockIdx.x + (idx / kWarps) * gridDim.x;
        const uint32_t b = tile * kPropThreads + (idx % kWarps) * 32u + lane;
        const bool live = b < p.B;
        // Most proposal samples receive no gradient at all (the proposal loss only pushes where the final level's
        // weight exceeds the proposal's bound): a warp whose 32 incoming gradients are all zero contributes nothing to
        // the table, to dW1 or to dW2 and skips the tile outright.
        const float gs = live ? __ldg(g_sigma + b) : 0.0f;
        if (__ballot_sync(0xffffffffu, gs != 0.0f) == 0u) continue;
        float x[3] = {0.5f, 0.5f, 0.5f};
        if (live) {
#pragma unroll
            for (uint32_t d = 0; d < 3; ++d) x[d] = __ldg(p.x01 + (size_t)b * 3 + d);
        }
        const bool ok = live && !out_of_range<3>(x);
        // ---- 1. forward recompute
        float enc[IN];
        if constexpr (SAVED) {
#pragma unroll
            for (uint32_t l = 0; l < L; ++l) {
                const float2 e = live ? __ldg(reinterpret_cast<const float2*>(enc_saved + (size_t)b * IN + 2 * l))
                                      : make_float2(0.0f, 0.0f);
                enc[2 * l] = e.x;
                enc[2 * l + 1] = e.y;
            }
        } else {
#pragma unroll
        for (uint32_t l = 0; l < L; ++l) {
            const LevelGeom<3> geo = s_geo[l];
            const Cell<3> cell = locate<3>(geo, x, false, 0u);
            const size_t base = (size_t)s_base[l];
            float a0 = 0.0f, a1 = 0.0f;
            float2 val[8];
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k)
                val[k] = ok ? __ldg(reinterpret_cast<const float2*>(table + (base + corner_row<3>(geo, cell, k)) * 2))
                            : make_float2(0.0f, 0.0f);
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
                const float w = corner_weight<3>(cell, k);
                a0 = __fmaf_rn(w, val[k].x, a0);
                a1 = __fmaf_rn(w, val[k].y, a1);
            }
            enc[2 * l] = a0;
            enc[2 * l + 1] = a1;
        }
        }
        // one pass over the weights: h_j and, for the backward, u_i = sum_j [h_j > 0] W2_j W1_ji  (d enc_i = dpre * u_i)
        float h[kPropHidden], u[IN], pre = 0.0f;
#pragma unroll
        for (uint32_t i = 0; i < IN; ++i) u[i] = 0.0f;
#pragma unroll
        for (uint32_t j = 0; j < kPropHidden; ++j) {
            float wj[IN], a = 0.0f;
#pragma unroll
            for (uint32_t i = 0; i < IN; ++i) { wj[i] = s_w1[j * IN + i]; a = __fmaf_rn(wj[i], enc[i], a); }
            h[j] = fmaxf(a, 0.0f);
            const float w2j = s_w2[j];
            pre = __fmaf_rn(w2j, h[j], pre);
            const float gate = (a > 0.0f) ? w2j : 0.0f;                  // ReLU mask
#pragma unroll
            for (uint32_t i = 0; i < IN; ++i) u[i] = __fmaf_rn(gate, wj[i], u[i]);
        }
        // trunc_exp backward (activation.py:16)
        const float dpre = gs * expf(fminf(fmaxf(pre, -15.0f), 15.0f));
        float denc[IN];
#pragma unroll
        for (uint32_t i = 0; i < IN; ++i) denc[i] = dpre * u[i];
        // ---- 2. weight gradients:  dW1[j][i] = W2[j] * sum_s [h_sj > 0] dpre_s enc_si ;  dW2[j] = sum_s dpre_s h_sj
#pragma unroll
        for (uint32_t j = 0; j < kPropHidden; j += 4)
            *reinterpret_cast<float4*>(row + j) = make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]);
#pragma unroll
        for (uint32_t half = 0; half < 2; ++half) {
            float e[8];
#pragma unroll
            for (uint32_t q = 0; q < 8; ++q) e[q] = (q < L) ? enc[half * L + q] : 0.0f;
            *reinterpret_cast<float4*>(row + 16 + half * 8) = make_float4(e[0], e[1], e[2], e[3]);
            if (L > 4) *reinterpret_cast<float4*>(row + 20 + half * 8) = make_float4(e[4], e[5], e[6], e[7]);
        }
        row[32] = dpre;
        __syncwarp();
#pragma unroll 4
        for (uint32_t s = 0; s < 32; ++s) {
            const float* r = rows + s * kPropRow;
            const float dp = r[32];
            if (dp == 0.0f) continue;                       // warp-uniform (broadcast read): this sample contributes nothing
            const float hs = r[own_j];
            const float m = (hs > 0.0f) ? dp : 0.0f;
            const float4 e0 = *reinterpret_cast<const float4*>(r + 16 + own_half * 8);
            float e[8] = {e0.x, e0.y, e0.z, e0.w, 0.0f, 0.0f, 0.0f, 0.0f};
            if (L > 4) {
                const float4 e1 = *reinterpret_cast<const float4*>(r + 20 + own_half * 8);
                e[4] = e1.x; e[5] = e1.y; e[6] = e1.z; e[7] = e1.w;
            }
#pragma unroll
            for (uint32_t q = 0; q < L; ++q) acc1[q] = __fmaf_rn(m, e[q], acc1[q]);
            acc2 = __fmaf_rn(dp, hs, acc2);
        }
        __syncwarp();
        // ---- 3. table gradient (gridencoder.cu:313-347), warp-aggregated
        const bool contributes = ok && dpre != 0.0f;
        // the level loop stays rolled (register pressure): the gradient of the encoding is parked in this lane's row
#pragma unroll
        for (uint32_t i = 0; i < IN; ++i) row[i] = denc[i];
#pragma unroll 1
        for (uint32_t l = 0; l < L; ++l) {
            const LevelGeom<3> geo = s_geo[l];
            const Cell<3> cell = locate<3>(geo, x, false, 0u);
            const float d0 = row[2 * l], d1 = row[2 * l + 1];
            float v[16];
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
                const float w = contributes ? corner_weight<3>(cell, k) : 0.0f;
                v[2 * k] = w * d0;
                v[2 * k + 1] = w * d1;
            }
            // lanes that do not contribute get a key no cell has (their own runs, which are skipped)
            const uint32_t key[3] = {contributes ? cell.lo[0] : 0xffffffffu - lane, cell.lo[1], cell.lo[2]};
            const bool head = warp_run_reduce<16, 3>(v, key, lane);
            if (head && contributes) {
                float* slice = grad_table + (size_t)s_base[l] * 2;
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k)
                    red_add_v2_f32(slice + (size_t)corner_row<3>(geo, cell, k) * 2, v[2 * k], v[2 * k + 1]);
            }
        }
    }
    // ---- cross-warp reduction of the weight gradients, one atomic per entry per CTA
    __syncthreads();
    float* scratch = s_rows;                              // [kWarps][32][L+1]
#pragma unroll
    for (uint32_t q = 0; q < L; ++q) scratch[(warp * 32 + lane) * (L + 1) + q] = acc1[q];
    scratch[(warp * 32 + lane) * (L + 1) + L] = acc2;
    __syncthreads();
    for (uint32_t e = tid; e < 32 * (L + 1); e += blockDim.x) {
        float a = 0.0f;
#pragma unroll
        for (uint32_t w = 0; w < kWarps; ++w) a += scratch[w * 32 * (L + 1) + e];
        const uint32_t ln = e / (L + 1), q = e - ln * (L + 1);
        const uint32_t j = ln & 15u, half = ln >> 4;
        if (q < L) red_add_f32(grad_w1 + j * IN + half * L + q, a * s_w2[j]);
        else if (half == 0) red_add_f32(grad_w2 + j, a);
    }
}

static int check_prop(const PropParams& p) {
    SANERF_REQUIRE_PTR(p.x01); SANERF_REQUIRE_PTR(p.table); SANERF_REQUIRE_PTR(p.offsets);
    SANERF_REQUIRE_PTR(p.w1); SANERF_REQUIRE_PTR(p.w2);
    if (p.L < 1 || p.L > 8) return fail(SANERF_ERR_INVALID_ARG, "prop_density: 1 <= L <= 8 levels of F=2 features");
    return SANERF_OK;
}

}  // namespace sanerf

using namespace sanerf;

#define SANERF_PROP_DISPATCH(L_, CALL)                                  \
    switch (L_) {                                                       \
        case 1: { constexpr uint32_t LL = 1; CALL; } break;             \
        case 2: { constexpr uint32_t LL = 2; CALL; } break;             \
        case 3: { constexpr uint32_t LL = 3; CALL; } break;             \
        case 4: { constexpr uint32_t LL = 4; CALL; } break;             \
        case 5: { constexpr uint32_t LL = 5; CALL; } break;             \
        case 6: { constexpr uint32_t LL = 6; CALL; } break;             \
        case 7: { constexpr uint32_t LL = 7; CALL; } break;             \
        default: { constexpr uint32_t LL = 8; CALL; } break;            \
    }

extern "C" int sanerf_prop_density_forward(const float* x01, const float* table, const int32_t* offsets,
                                           const float* w1, const float* w2, uint32_t B, uint32_t L, float S,
                                           uint32_t H, float* sigma, float* enc_out, void* stream) {
    if (B == 0) return SANERF_OK;
    PropParams p{x01, table, offsets, w1, w2, B, L, H, S};
    int rc = check_prop(p);
    if (rc != SANERF_OK) return rc;
    SANERF_REQUIRE_PTR(sigma);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t tiles = div_up(B, kPropThreads);
    static const uint32_t per_sm = [] { const char* e = getenv("SANERF_PROP_FWD_CTAS_PER_SM"); return e ? (uint32_t)atoi(e) : 4u; }();
    const uint32_t cap = (uint32_t)kNumSMs * per_sm;
    const uint32_t blocks = tiles < cap ? tiles : cap;
    SANERF_PROP_DISPATCH(L, (SANERF_LAUNCH((prop_forward_kernel<LL>), blocks, kPropThreads, 0, st, p, sigma, enc_out)));
    return check_launch("prop_forward_kernel");
}

extern "C" int sanerf_prop_density_backward(const float* x01, const float* table, const int32_t* offsets,
                                            const float* w1, const float* w2, uint32_t B, uint32_t L, float S,
                                            uint32_t H, const float* enc, const float* g_sigma, float* grad_table,
                                            float* grad_w1, float* grad_w2, void* stream) {
    if (B == 0) return SANERF_OK;
    PropParams p{x01, table, offsets, w1, w2, B, L, H, S};
    int rc = check_prop(p);
    if (rc != SANERF_OK) return rc;
    SANERF_REQUIRE_PTR(g_sigma); SANERF_REQUIRE_PTR(grad_table); SANERF_REQUIRE_PTR(grad_w1); SANERF_REQUIRE_PTR(grad_w2);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t tiles = div_up(B, kPropThreads);
    const uint32_t cap = (uint32_t)kNumSMs * 2u;                              // persistent: resident CTAs only
    const uint32_t blocks = tiles < cap ? tiles : cap;
    if (enc != nullptr) {
        SANERF_PROP_DISPATCH(L, (SANERF_LAUNCH((prop_backward_kernel<LL, true>), blocks, kPropThreads, 0, st, p, enc, g_sigma, grad_table,
                                                                                                grad_w1, grad_w2, tiles)));
    } else {
        SANERF_PROP_DISPATCH(L, (SANERF_LAUNCH((prop_backward_kernel<LL, false>), blocks, kPropThreads, 0, st, p, enc, g_sigma, grad_table,
                                                                                                 grad_w1, grad_w2, tiles)));
    }
    return check_launch("prop_backward_kernel");
}
