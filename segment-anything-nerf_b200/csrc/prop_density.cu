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

template <uint32_t L>
__device__ __forceinline__ void prop_encode(const PropParams& p, const float (&x)[3], bool oob, float (&enc)[2 * L]) {
    const float* __restrict__ table = p.table;
    float2 val[L][8];
    float frac[L][3];
#pragma unroll
    for (uint32_t l = 0; l < L; ++l) {
        if (!oob) {
            const LevelGeom<3> geo = level_geometry<3>(p.offsets, l, p.S, p.H, 0u);
            const Cell<3> cell = locate<3>(geo, x, false, 0u);
            const size_t base = (size_t)(uint32_t)__ldg(p.offsets + l);
#pragma unroll
            for (uint32_t d = 0; d < 3; ++d) frac[l][d] = cell.f[d];
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k)
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
__global__ void __launch_bounds__(kPropThreads) prop_forward_kernel(const PropParams p, float* __restrict__ sigma) {
    constexpr uint32_t IN = 2 * L;
    __shared__ float s_w1[kPropHidden * IN];
    __shared__ float s_w2[kPropHidden];
    for (uint32_t i = threadIdx.x; i < kPropHidden * IN; i += blockDim.x) s_w1[i] = __ldg(p.w1 + i);
    if (threadIdx.x < kPropHidden) s_w2[threadIdx.x] = __ldg(p.w2 + threadIdx.x);
    __syncthreads();
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    float x[3];
#pragma unroll
    for (uint32_t d = 0; d < 3; ++d) x[d] = __ldg(p.x01 + (size_t)b * 3 + d);
    float enc[IN];
    prop_encode<L>(p, x, out_of_range<3>(x), enc);
    float pre = 0.0f;
#pragma unroll
    for (uint32_t j = 0; j < kPropHidden; ++j) {
        float h = 0.0f;
#pragma unroll
        for (uint32_t i = 0; i < IN; ++i) h = __fmaf_rn(s_w1[j * IN + i], enc[i], h);
        pre = __fmaf_rn(s_w2[j], fmaxf(h, 0.0f), pre);
    }
    sigma[b] = expf(pre);                    // trunc_exp forward (activation.py:10)
}

template <uint32_t L>
__global__ void __launch_bounds__(kPropThreads) prop_backward_kernel(const PropParams p, const float* __restrict__ g_sigma,
                                                                     float* __restrict__ grad_table,
                                                                     float* __restrict__ grad_w1,
                                                                     float* __restrict__ grad_w2, uint32_t tiles) {
    constexpr uint32_t IN = 2 * L;
    constexpr uint32_t ROW = kPropHidden + IN + 1;       // h[16] | enc[IN] | dpre (odd length: bank-conflict free)
    __shared__ float s_w1[kPropHidden * IN];
    __shared__ float s_w2[kPropHidden];
    __shared__ float s_rows[kPropThreads * ROW];
    for (uint32_t i = threadIdx.x; i < kPropHidden * IN; i += blockDim.x) s_w1[i] = __ldg(p.w1 + i);
    if (threadIdx.x < kPropHidden) s_w2[threadIdx.x] = __ldg(p.w2 + threadIdx.x);
    __syncthreads();

    // entries of dW owned by this thread in the reduction phase: e < 16*IN -> dW1[j][i]; then dW2[j]
    constexpr uint32_t kEntries = kPropHidden * IN + kPropHidden;
    constexpr uint32_t kOwn = (kEntries + kPropThreads - 1) / kPropThreads;
    float acc[kOwn];
#pragma unroll
    for (uint32_t q = 0; q < kOwn; ++q) acc[q] = 0.0f;

    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t b = tile * kPropThreads + threadIdx.x;
        float* row = s_rows + (size_t)threadIdx.x * ROW;     // h[16] | enc[IN] | dpre
        const bool live = b < p.B;
        float x[3] = {0.5f, 0.5f, 0.5f};
        if (live) {
#pragma unroll
            for (uint32_t d = 0; d < 3; ++d) x[d] = __ldg(p.x01 + (size_t)b * 3 + d);
        }
        const bool oob = out_of_range<3>(x);
        float enc[IN], h[kPropHidden];
        prop_encode<L>(p, x, oob || !live, enc);
        float pre = 0.0f;
#pragma unroll
        for (uint32_t j = 0; j < kPropHidden; ++j) {
            float a = 0.0f;
#pragma unroll
            for (uint32_t i = 0; i < IN; ++i) a = __fmaf_rn(s_w1[j * IN + i], enc[i], a);
            h[j] = fmaxf(a, 0.0f);
            pre = __fmaf_rn(s_w2[j], h[j], pre);
        }
        // trunc_exp backward (activation.py:16)
        const float dpre = live ? __ldg(g_sigma + b) * expf(fminf(fmaxf(pre, -15.0f), 15.0f)) : 0.0f;
        float denc[IN];
#pragma unroll
        for (uint32_t i = 0; i < IN; ++i) denc[i] = 0.0f;
#pragma unroll
        for (uint32_t j = 0; j < kPropHidden; ++j) {
            const float dh = (h[j] > 0.0f) ? dpre * s_w2[j] : 0.0f;      // ReLU mask
            row[j] = h[j];
#pragma unroll
            for (uint32_t i = 0; i < IN; ++i) denc[i] = __fmaf_rn(s_w1[j * IN + i], dh, denc[i]);
        }
#pragma unroll
        for (uint32_t i = 0; i < IN; ++i) row[kPropHidden + i] = enc[i];
        row[ROW - 1] = dpre;

        // table gradient: scatter denc through the interpolation weights (gridencoder.cu:313-347)
        if (live && !oob && dpre != 0.0f) {
#pragma unroll
            for (uint32_t l = 0; l < L; ++l) {
                const LevelGeom<3> geo = level_geometry<3>(p.offsets, l, p.S, p.H, 0u);
                const Cell<3> cell = locate<3>(geo, x, false, 0u);
                float* slice = grad_table + (size_t)(uint32_t)__ldg(p.offsets + l) * 2;
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k) {
                    const float w = corner_weight<3>(cell, k);
                    red_add_v2_f32(slice + (size_t)corner_row<3>(geo, cell, k) * 2, w * denc[2 * l], w * denc[2 * l + 1]);
                }
            }
        }
        __syncthreads();
        // weight gradients: each thread sweeps the tile's samples for the entries it owns
        //   dW1[j][i] = W2[j] * sum_s [h_sj > 0] dpre_s enc_si      dW2[j] = sum_s dpre_s h_sj
#pragma unroll
        for (uint32_t q = 0; q < kOwn; ++q) {
            const uint32_t e = threadIdx.x + q * kPropThreads;
            if (e < kPropHidden * IN) {
                const uint32_t ej = e / IN, ei = e - ej * IN;
                float a = 0.0f;
#pragma unroll 8
                for (uint32_t s = 0; s < kPropThreads; ++s) {
                    const float* r = s_rows + s * ROW;
                    a = __fmaf_rn((r[ej] > 0.0f) ? r[ROW - 1] : 0.0f, r[kPropHidden + ei], a);
                }
                acc[q] += a * s_w2[ej];
            } else if (e < kEntries) {
                const uint32_t ej = e - kPropHidden * IN;
                float a = 0.0f;
#pragma unroll 8
                for (uint32_t s = 0; s < kPropThreads; ++s) {
                    const float* r = s_rows + s * ROW;
                    a = __fmaf_rn(r[ROW - 1], r[ej], a);
                }
                acc[q] += a;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (uint32_t q = 0; q < kOwn; ++q) {
        const uint32_t e = threadIdx.x + q * kPropThreads;
        if (e < kPropHidden * IN) red_add_f32(grad_w1 + e, acc[q]);
        else if (e < kEntries) red_add_f32(grad_w2 + (e - kPropHidden * IN), acc[q]);
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
                                           uint32_t H, float* sigma, void* stream) {
    if (B == 0) return SANERF_OK;
    PropParams p{x01, table, offsets, w1, w2, B, L, H, S};
    int rc = check_prop(p);
    if (rc != SANERF_OK) return rc;
    SANERF_REQUIRE_PTR(sigma);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t blocks = div_up(B, kPropThreads);
    SANERF_PROP_DISPATCH(L, (prop_forward_kernel<LL><<<blocks, kPropThreads, 0, st>>>(p, sigma)));
    return check_launch("prop_forward_kernel");
}

extern "C" int sanerf_prop_density_backward(const float* x01, const float* table, const int32_t* offsets,
                                            const float* w1, const float* w2, uint32_t B, uint32_t L, float S,
                                            uint32_t H, const float* g_sigma, float* grad_table, float* grad_w1,
                                            float* grad_w2, void* stream) {
    if (B == 0) return SANERF_OK;
    PropParams p{x01, table, offsets, w1, w2, B, L, H, S};
    int rc = check_prop(p);
    if (rc != SANERF_OK) return rc;
    SANERF_REQUIRE_PTR(g_sigma); SANERF_REQUIRE_PTR(grad_table); SANERF_REQUIRE_PTR(grad_w1); SANERF_REQUIRE_PTR(grad_w2);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t tiles = div_up(B, kPropThreads);
    const uint32_t blocks = tiles < (uint32_t)(kNumSMs * 4) ? tiles : (uint32_t)(kNumSMs * 4);
    SANERF_PROP_DISPATCH(L, (prop_backward_kernel<LL><<<blocks, kPropThreads, 0, st>>>(p, g_sigma, grad_table, grad_w1,
                                                                                      grad_w2, tiles)));
    return check_launch("prop_backward_kernel");
}
