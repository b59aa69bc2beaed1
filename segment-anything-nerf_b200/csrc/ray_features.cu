// Ray-composited hash-grid features (stage 2, SAM feature field) and their gradient scatter (sm_100a).
//
// Replaces, for the feature branch of NeRFRenderer.run (nerf/renderer.py:302-303 `features = self.s_grid(xyzs)`,
// :377 `f_sam = torch.sum(weights.unsqueeze(-1) * features, dim=-2)`), the chain
//   grid encode [N*T, L*C]  ->  broadcast multiply [N, T, L*C]  ->  sum over T            (forward)
//   expand g_out to [N, T, L*C] * weights  ->  permute copy  ->  zeros_like(table)  ->  scatter   (backward)
// by one kernel per direction that never materialises a per-sample feature or gradient matrix:
//   out[r, l*C + c]      = sum_i w[r,i] * enc_l,c(x[r,i])
//   g_table[row(l,x,k)] += cw_k(x[r,i]) * (w[r,i] * g_out[r, l*C + c])
//
// Work decomposition: one warp owns (ray, level); a lane owns one sample of the ray (chunks of 32).  Levels are the slow
// grid dimension, so the resident CTAs work on one level's slice of the table at a time (L2 / L1 locality, as in
// grid_encode.cu).  All 2^3 corner rows of a lane are in flight before the first FMA; the encoding of a sample is
// evaluated in the reference's order (corner 0..7, weight ((1*a0)*a1)*a2, one FMA per corner and channel), so
// w * enc is the same fp32 value the unfused path multiplies.
//
// Backward: the gradient of every sample of a ray is the SAME vector g_out[r, level] scaled by the sample's weight, so
// consecutive samples that fall into the same cell (coarse levels, or samples concentrated at a surface) are merged by
// a segmented warp reduction over just 2^3 scalars (w * corner weight) before any reduction is issued; the run head
// then fires one red.global.add.v4.f32 per 4 channels and corner.  Rays with an all-zero gradient row and samples
// with zero weight (early-terminated) issue nothing.
#include "grid_common.cuh"

namespace sanerf {

struct RayFeatParams {
    const float* x01;        // [N*T, 3] in [0,1]^3
    const float* weights;    // [N*T]
    const float* table;      // [rows, C]   (forward)
    const float* g_out;      // [N, L*C]    (backward)
    const int32_t* offsets;  // [L+1]
    float* out;              // [N, L*C]    (forward)
    float* g_table;          // [rows, C]   (backward, accumulated into)
    uint32_t N, T, L, H;
    float S;
    uint32_t level_begin, level_end;   // backward: levels [level_begin, level_end) of this launch (forward: all)
    uint32_t row_stride;               // floats between consecutive rays of out / g_out (>= L*C, multiple of 4)
};

constexpr uint32_t kRayWarps = 8;

template <uint32_t C>
__global__ void __launch_bounds__(kRayWarps * 32) ray_features_forward_kernel(const RayFeatParams p) {
    pdl_begin();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t ray = blockIdx.x * kRayWarps + warp, level = blockIdx.y;
    if (ray >= p.N) return;
    const LevelGeom<3> geo = level_geometry<3>(p.offsets, level, p.S, p.H, 0u);
    const float* __restrict__ slice = p.table + (size_t)(uint32_t)__ldg(p.offsets + level) * C;
    float sum[C];
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) sum[c] = 0.0f;
    for (uint32_t i0 = 0; i0 < p.T; i0 += 32u) {
        const uint32_t i = i0 + lane;
        const size_t s = (size_t)ray * p.T + i;
        const float wgt = (i < p.T) ? __ldg(p.weights + s) : 0.0f;
        float x[3] = {0.5f, 0.5f, 0.5f};
        if (i < p.T) {
#pragma unroll
            for (uint32_t d = 0; d < 3; ++d) x[d] = __ldg(p.x01 + s * 3 + d);
        }
        if (wgt == 0.0f || out_of_range<3>(x)) continue;     // OOB samples encode to zero (gridencoder.cu:105-130)
        const Cell<3> cell = locate<3>(geo, x, false, 0u);
        float val[8][C];
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) RowIO<float, C>::load(slice + (size_t)corner_row<3>(geo, cell, k) * C, val[k]);
        float acc[C];
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) acc[c] = 0.0f;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            const float cw = corner_weight<3>(cell, k);
#pragma unroll
            for (uint32_t c = 0; c < C; ++c) acc[c] = __fmaf_rn(cw, val[k][c], acc[c]);
        }
#pragma unroll
        for (uint32_t c = 0; c < C; ++c) sum[c] = __fmaf_rn(wgt, acc[c], sum[c]);
    }
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) sum[c] = warp_sum(sum[c]);
    if (lane == 0) RowIO<float, C>::store(p.out + (size_t)ray * p.row_stride + level * C, sum);
}

template <uint32_t C>
__global__ void __launch_bounds__(kRayWarps * 32) ray_features_backward_kernel(const RayFeatParams p) {
    pdl_begin();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t ray = blockIdx.x * kRayWarps + warp, level = p.level_begin + blockIdx.y;
    if (ray >= p.N) return;
    float g[C];
    RowIO<float, C>::load(p.g_out + (size_t)ray * p.row_stride + level * C, g);     // same address in every lane: one broadcast
    bool any = false;
#pragma unroll
    for (uint32_t c = 0; c < C; ++c) any |= (g[c] != 0.0f);
    if (!any) return;
    const LevelGeom<3> geo = level_geometry<3>(p.offsets, level, p.S, p.H, 0u);
    float* __restrict__ slice = p.g_table + (size_t)(uint32_t)__ldg(p.offsets + level) * C;
    for (uint32_t i0 = 0; i0 < p.T; i0 += 32u) {
        const uint32_t i = i0 + lane;
        const size_t s = (size_t)ray * p.T + i;
        const float wgt = (i < p.T) ? __ldg(p.weights + s) : 0.0f;
        float x[3] = {0.5f, 0.5f, 0.5f};
        if (i < p.T) {
#pragma unroll
            for (uint32_t d = 0; d < 3; ++d) x[d] = __ldg(p.x01 + s * 3 + d);
        }
        const bool contributes = (wgt != 0.0f) && !out_of_range<3>(x);      // gridencoder.cu:279-284: OOB gradient is dropped
        const Cell<3> cell = locate<3>(geo, x, false, 0u);
        float sc[8];
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) sc[k] = contributes ? corner_weight<3>(cell, k) * wgt : 0.0f;
        uint32_t key[3];
#pragma unroll
        for (uint32_t d = 0; d < 3; ++d) key[d] = cell.lo[d];
        if (!contributes) key[0] = 0xffffffffu - lane;                      // a key no cell has: its own (skipped) run
        const bool head = warp_run_reduce<8, 3>(sc, key, lane);
        if (head && contributes) {
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
                float upd[C];
#pragma unroll
                for (uint32_t c = 0; c < C; ++c) upd[c] = sc[k] * g[c];
                RowIO<float, C>::red(slice + (size_t)corner_row<3>(geo, cell, k) * C, upd);
            }
        }
    }
}

template <bool kBackward>
static int launch_ray_features(const RayFeatParams& p, uint32_t C, cudaStream_t st) {
    if (p.N == 0 || p.T == 0 || p.L == 0) return SANERF_OK;
    dim3 grid(div_up(p.N, kRayWarps), kBackward ? p.level_end - p.level_begin : p.L, 1);
    const uint32_t threads = kRayWarps * 32;
    switch (C) {
        case 2:
            if (kBackward) SANERF_LAUNCH((ray_features_backward_kernel<2>), grid, threads, 0, st, p);
            else SANERF_LAUNCH((ray_features_forward_kernel<2>), grid, threads, 0, st, p);
            break;
        case 4:
            if (kBackward) SANERF_LAUNCH((ray_features_backward_kernel<4>), grid, threads, 0, st, p);
            else SANERF_LAUNCH((ray_features_forward_kernel<4>), grid, threads, 0, st, p);
            break;
        case 8:
            if (kBackward) SANERF_LAUNCH((ray_features_backward_kernel<8>), grid, threads, 0, st, p);
            else SANERF_LAUNCH((ray_features_forward_kernel<8>), grid, threads, 0, st, p);
            break;
        default: return fail(SANERF_ERR_INVALID_ARG, "ray features: C must be 2, 4 or 8");
    }
    return check_launch(kBackward ? "ray_features_backward_kernel" : "ray_features_forward_kernel");
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_ray_features_forward(const float* x01, const float* weights, const float* embeddings,
                                           const int32_t* offsets, uint32_t N, uint32_t T, uint32_t C, uint32_t L, float S,
                                           uint32_t H, float* out, uint32_t out_stride, void* stream) {
    if (N == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(x01);
    SANERF_REQUIRE_PTR(weights);
    SANERF_REQUIRE_PTR(embeddings);
    SANERF_REQUIRE_PTR(offsets);
    SANERF_REQUIRE_PTR(out);
    if (out_stride == 0) out_stride = L * C;
    if (out_stride < L * C || (out_stride & 3u)) return fail(SANERF_ERR_INVALID_ARG, "ray_features: row stride must be >= L*C and a multiple of 4");
    RayFeatParams p{x01, weights, embeddings, nullptr, offsets, out, nullptr, N, T, L, H, S, 0u, L, out_stride};
    return launch_ray_features<false>(p, C, static_cast<cudaStream_t>(stream));
}

extern "C" int sanerf_ray_features_backward(const float* x01, const float* weights, const float* g_out,
                                            const int32_t* offsets, uint32_t N, uint32_t T, uint32_t C, uint32_t L, float S,
                                            uint32_t H, float* grad_embeddings, uint32_t level_begin, uint32_t level_end,
                                            uint32_t g_stride, void* stream) {
    if (level_end > L) level_end = L;
    if (N == 0 || level_begin >= level_end) return SANERF_OK;
    SANERF_REQUIRE_PTR(x01);
    SANERF_REQUIRE_PTR(weights);
    SANERF_REQUIRE_PTR(g_out);
    SANERF_REQUIRE_PTR(offsets);
    SANERF_REQUIRE_PTR(grad_embeddings);
    if (g_stride == 0) g_stride = L * C;
    if (g_stride < L * C || (g_stride & 3u)) return fail(SANERF_ERR_INVALID_ARG, "ray_features: row stride must be >= L*C and a multiple of 4");
    RayFeatParams p{x01, weights, nullptr, g_out, offsets, nullptr, grad_embeddings, N, T, L, H, S, level_begin, level_end, g_stride};
    return launch_ray_features<true>(p, C, static_cast<cudaStream_t>(stream));
}
