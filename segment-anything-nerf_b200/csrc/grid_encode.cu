// Multiresolution hash-grid encoder, forward (sm_100a).
//
// Replaces kernel_grid / grid_encode_forward of the reference
// (gridencoder/src/gridencoder.cu:82-249, 467-490).  Not a port: the work decomposition is
// different.
//
//  * One thread owns one sample and a GROUP of G consecutive levels, chosen so that the
//    group's features are one full 32-byte sector of the [B, L*C] output row (G*C*sizeof(T)
//    == 32 B where possible).  All G*2^D table gathers of a thread are issued before the
//    first FMA (G*2^D independent 8/16/32-byte loads in flight per thread).
//  * Level groups are the slow grid dimension, so at any moment the resident CTAs work on
//    a few levels only and that slice of the table stays L2 / L1 resident (the same reason the
//    reference puts `level` in blockIdx.y), while the output goes straight to the [B, L*C]
//    layout GridEncoder.forward returns — no [L,B,C] intermediate + permute copy
//    (grid.py:63), which doubled the output traffic.
//  * Table rows are fetched as single 32/64/128-bit read-only loads.
//
// fp32 results are bit-identical to the reference kernel: same corner order, same weight
// product order, one FMA per (corner, channel).  fp16 tables accumulate in fp32 and round
// once (the reference rounds to half after every corner).
#include "grid_common.cuh"

namespace sanerf {

struct GridFwdParams {
    const float* inputs;
    const void* table;
    const int32_t* offsets;
    void* outputs;
    void* dy_dx;
    uint32_t B, L, max_level, H;
    float S;
    uint32_t gridtype, interp;
    int align_corners, blc, zero_tail;
};

template <typename T, uint32_t D, uint32_t C, uint32_t G, uint32_t CH>
__global__ void __launch_bounds__(256) grid_forward_kernel(const GridFwdParams p) {
    static_assert(C % CH == 0, "channel chunk must divide C");
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const uint32_t level0 = blockIdx.y * G;

    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) x[d] = __ldg(p.inputs + (size_t)b * D + d);
    const bool oob = out_of_range<D>(x);

    const T* __restrict__ table = static_cast<const T*>(p.table);
    T* __restrict__ out = static_cast<T*>(p.outputs);

    // phase 0: per-level geometry -> element offsets of the 2^D corner rows + fractions
    size_t elem[G][1u << D];
    float frac[G][D];
    bool live[G];
#pragma unroll
    for (uint32_t g = 0; g < G; ++g) {
        const uint32_t level = level0 + g;
        live[g] = (level < p.max_level) && !oob;
        if (live[g]) {
            const LevelGeom<D> geo = level_geometry<D>(p.offsets, level, p.S, p.H, p.gridtype);
            const Cell<D> cell = locate<D>(geo, x, p.align_corners != 0, p.interp);
            const size_t base = (size_t)(uint32_t)__ldg(p.offsets + level);
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) frac[g][d] = cell.f[d];
#pragma unroll
            for (uint32_t k = 0; k < (1u << D); ++k)
                elem[g][k] = (base + corner_row<D>(geo, cell, k)) * C;
        } else {
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) frac[g][d] = 0.0f;
#pragma unroll
            for (uint32_t k = 0; k < (1u << D); ++k) elem[g][k] = 0;
        }
    }

#pragma unroll 1
    for (uint32_t c0 = 0; c0 < C; c0 += CH) {
        // phase 1: every gather of the group in flight before the first FMA
        float val[G][1u << D][CH];
#pragma unroll
        for (uint32_t g = 0; g < G; ++g) {
            if (live[g]) {
#pragma unroll
                for (uint32_t k = 0; k < (1u << D); ++k)
                    RowIO<T, CH>::load(table + elem[g][k] + c0, val[g][k]);
            } else {
#pragma unroll
                for (uint32_t k = 0; k < (1u << D); ++k)
#pragma unroll
                    for (uint32_t c = 0; c < CH; ++c) val[g][k][c] = 0.0f;
            }
        }
        // phase 2: interpolate — corner order 0..2^D-1, weight = ((1*a0)*a1)*a2, one FMA per
        // channel: the reference's exact operation order (gridencoder.cu:171-192)
#pragma unroll
        for (uint32_t g = 0; g < G; ++g) {
            const uint32_t level = level0 + g;
            float acc[CH];
#pragma unroll
            for (uint32_t c = 0; c < CH; ++c) acc[c] = 0.0f;
#pragma unroll
            for (uint32_t k = 0; k < (1u << D); ++k) {
                float w = 1.0f;
#pragma unroll
                for (uint32_t d = 0; d < D; ++d) w *= (k & (1u << d)) ? frac[g][d] : (1.0f - frac[g][d]);
#pragma unroll
                for (uint32_t c = 0; c < CH; ++c) acc[c] = __fmaf_rn(w, val[g][k][c], acc[c]);
            }
            // phase 3: store (levels >= max_level only when asked to zero the tail)
            if (level < p.L && (level < p.max_level || p.zero_tail)) {
                T* dst = p.blc ? out + ((size_t)b * p.L + level) * C + c0
                               : out + ((size_t)level * p.B + b) * C + c0;
                RowIO<T, CH>::store(dst, acc);
            }
        }
    }

}

// Cold path: d out / d x (gridencoder.cu:205-248), layout [B, L, D, C].  xyz never requires
// grad in the shipped networks (grid.py:163 passes inputs.requires_grad), so this is a plain
// one-thread-per-(sample, level) kernel kept out of the hot kernel's register budget.
template <typename T, uint32_t D>
__global__ void __launch_bounds__(256) grid_dydx_kernel(const GridFwdParams p, const uint32_t C) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const uint32_t level = blockIdx.y;
    if (level >= p.max_level && !p.zero_tail) return;
    const T* __restrict__ table = static_cast<const T*>(p.table);
    T* __restrict__ dst = static_cast<T*>(p.dy_dx) + (((size_t)b * p.L + level) * D) * C;

    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) x[d] = __ldg(p.inputs + (size_t)b * D + d);
    if (out_of_range<D>(x) || level >= p.max_level) {
        for (uint32_t i = 0; i < D * C; ++i) dst[i] = from_float<T>(0.0f);
        return;
    }
    const LevelGeom<D> geo = level_geometry<D>(p.offsets, level, p.S, p.H, p.gridtype);
    const Cell<D> cell = locate<D>(geo, x, p.align_corners != 0, p.interp);
    const T* __restrict__ slice = table + (size_t)(uint32_t)__ldg(p.offsets + level) * C;
    const float scale = (float)(p.align_corners ? geo.res - 1u : geo.res);
#pragma unroll
    for (uint32_t gd = 0; gd < D; ++gd) {
        for (uint32_t c = 0; c < C; ++c) {
            float r = 0.0f;
#pragma unroll
            for (uint32_t k = 0; k < (1u << (D - 1)); ++k) {
                // spread the D-1 bits of k over the dims != gd
                float w = scale;
                uint32_t corner = 0;
#pragma unroll
                for (uint32_t nd = 0; nd < D - 1; ++nd) {
                    const uint32_t d = (nd >= gd) ? nd + 1 : nd;
                    const bool up = (k >> nd) & 1u;
                    w *= up ? cell.f[d] : (1.0f - cell.f[d]);
                    corner |= up ? (1u << d) : 0u;
                }
                const float lo = to_float<T>(slice[(size_t)corner_row<D>(geo, cell, corner) * C + c]);
                const float hi = to_float<T>(slice[(size_t)corner_row<D>(geo, cell, corner | (1u << gd)) * C + c]);
                r = __fmaf_rn(w * (hi - lo), cell.df[gd], r);
            }
            dst[gd * C + c] = from_float<T>(r);
        }
    }
}

// Debug / parity: dump table rows and device-side level geometry.
template <uint32_t D>
__global__ void grid_dump_kernel(const float* __restrict__ inputs, const int32_t* __restrict__ offsets,
                                 uint32_t* __restrict__ rows, uint32_t* __restrict__ geometry,
                                 uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                 int align_corners) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t level = blockIdx.y;
    const LevelGeom<D> geo = level_geometry<D>(offsets, level, S, H, gridtype);
    if (geometry != nullptr && b == 0) {
        geometry[level * 4 + 0] = geo.res;
        geometry[level * 4 + 1] = geo.rows;
        geometry[level * 4 + 2] = geo.hashed ? 1u : 0u;
        geometry[level * 4 + 3] = geo.covered;
    }
    if (b >= B) return;
    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) x[d] = inputs[(size_t)b * D + d];
    uint32_t* dst = rows + ((size_t)b * L + level) * (1u << D);
    if (out_of_range<D>(x)) {
        for (uint32_t k = 0; k < (1u << D); ++k) dst[k] = 0xffffffffu;
        return;
    }
    const Cell<D> cell = locate<D>(geo, x, align_corners != 0, 0u);
#pragma unroll
    for (uint32_t k = 0; k < (1u << D); ++k) dst[k] = corner_row<D>(geo, cell, k);
}

// ---- dispatch --------------------------------------------------------------------------
template <typename T, uint32_t D, uint32_t C>
static int launch_forward(const GridFwdParams& p, cudaStream_t stream) {
    // Channel chunk CH and level-group size G.  G aims at one 32-byte sector of the
    // [B, L*C] row per thread, capped by a register budget of 64 gathered floats per thread.
    constexpr uint32_t kCorners = 1u << D;
    constexpr uint32_t kBudget = 64;
    constexpr uint32_t CH0 = (C < 8u) ? C : 8u;
    constexpr uint32_t CH = (CH0 * kCorners <= kBudget) ? CH0 : ((kBudget / kCorners) ? kBudget / kCorners : 1u);
    constexpr uint32_t kBytes = C * sizeof(T);
    constexpr uint32_t kSector = (kBytes >= 32u) ? 1u : 32u / kBytes;
    constexpr uint32_t kRegG = (kBudget / (kCorners * CH)) ? kBudget / (kCorners * CH) : 1u;
    constexpr uint32_t G = (CH < C) ? 1u : (kSector < kRegG ? kSector : kRegG);
    constexpr uint32_t kThreads = 256;
    const uint32_t n_levels = p.zero_tail ? p.L : p.max_level;
    if (n_levels == 0 || p.B == 0) return SANERF_OK;
    dim3 grid(div_up(p.B, kThreads), div_up(n_levels, G), 1);
    grid_forward_kernel<T, D, C, G, CH><<<grid, kThreads, 0, stream>>>(p);
    if (p.dy_dx != nullptr) {
        dim3 grid2(div_up(p.B, kThreads), n_levels, 1);
        grid_dydx_kernel<T, D><<<grid2, kThreads, 0, stream>>>(p, C);
    }
    return check_launch("grid_forward_kernel");
}

template <typename T, uint32_t D>
static int dispatch_C(const GridFwdParams& p, uint32_t C, cudaStream_t stream) {
    switch (C) {
        case 1: return launch_forward<T, D, 1>(p, stream);
        case 2: return launch_forward<T, D, 2>(p, stream);
        case 4: return launch_forward<T, D, 4>(p, stream);
        case 8: return launch_forward<T, D, 8>(p, stream);
        case 16: return launch_forward<T, D, 16>(p, stream);
        case 32: return launch_forward<T, D, 32>(p, stream);
        default: return fail(SANERF_ERR_INVALID_ARG, "GridEncoding: C must be 1, 2, 4, 8, 16 or 32.");
    }
}

template <typename T>
static int dispatch_D(const GridFwdParams& p, uint32_t D, uint32_t C, cudaStream_t stream) {
    switch (D) {
        case 2: return dispatch_C<T, 2>(p, C, stream);
        case 3: return dispatch_C<T, 3>(p, C, stream);
        case 4: return dispatch_C<T, 4>(p, C, stream);
        case 5: return dispatch_C<T, 5>(p, C, stream);
        default: return fail(SANERF_ERR_INVALID_ARG, "GridEncoding: D must be 2, 3, 4 or 5.");
    }
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_grid_encode_forward(const float* inputs, const void* embeddings,
                                          const int32_t* offsets, void* outputs, uint32_t B,
                                          uint32_t D, uint32_t C, uint32_t L, uint32_t max_level,
                                          float S, uint32_t H, void* dy_dx, uint32_t gridtype,
                                          int align_corners, uint32_t interp, int dtype,
                                          int out_layout, int zero_tail, void* stream) {
    if (B == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(inputs);
    SANERF_REQUIRE_PTR(embeddings);
    SANERF_REQUIRE_PTR(offsets);
    SANERF_REQUIRE_PTR(outputs);
    if (out_layout != SANERF_LAYOUT_LBC && out_layout != SANERF_LAYOUT_BLC)
        return fail(SANERF_ERR_INVALID_ARG, "out_layout must be SANERF_LAYOUT_LBC or SANERF_LAYOUT_BLC");
    if (gridtype > 1u) return fail(SANERF_ERR_INVALID_ARG, "gridtype must be 0 (hash) or 1 (tiled)");
    if (interp > 1u) return fail(SANERF_ERR_INVALID_ARG, "interp must be 0 (linear) or 1 (smoothstep)");
    if (max_level > L) max_level = L;
    GridFwdParams p;
    p.inputs = inputs; p.table = embeddings; p.offsets = offsets; p.outputs = outputs; p.dy_dx = dy_dx;
    p.B = B; p.L = L; p.max_level = max_level; p.H = H; p.S = S;
    p.gridtype = gridtype; p.interp = interp; p.align_corners = align_corners;
    p.blc = (out_layout == SANERF_LAYOUT_BLC); p.zero_tail = zero_tail;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case SANERF_F32: return dispatch_D<float>(p, D, C, st);
        case SANERF_F16: return dispatch_D<__half>(p, D, C, st);
        default: return fail(SANERF_ERR_INVALID_ARG, "dtype must be SANERF_F32 or SANERF_F16");
    }
}

extern "C" int sanerf_grid_dump_indices(const float* inputs, const int32_t* offsets, uint32_t* rows,
                                        uint32_t* geometry, uint32_t B, uint32_t D, uint32_t L,
                                        float S, uint32_t H, uint32_t gridtype, int align_corners,
                                        void* stream) {
    SANERF_REQUIRE_PTR(offsets);
    if (B > 0) { SANERF_REQUIRE_PTR(inputs); SANERF_REQUIRE_PTR(rows); }
    if (L == 0) return SANERF_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(div_up(B > 0 ? B : 1u, 128u), L, 1);
    switch (D) {
        case 2: grid_dump_kernel<2><<<grid, 128, 0, st>>>(inputs, offsets, rows, geometry, B, L, S, H, gridtype, align_corners); break;
        case 3: grid_dump_kernel<3><<<grid, 128, 0, st>>>(inputs, offsets, rows, geometry, B, L, S, H, gridtype, align_corners); break;
        case 4: grid_dump_kernel<4><<<grid, 128, 0, st>>>(inputs, offsets, rows, geometry, B, L, S, H, gridtype, align_corners); break;
        case 5: grid_dump_kernel<5><<<grid, 128, 0, st>>>(inputs, offsets, rows, geometry, B, L, S, H, gridtype, align_corners); break;
        default: return fail(SANERF_ERR_INVALID_ARG, "GridEncoding: D must be 2, 3, 4 or 5.");
    }
    return check_launch("grid_dump_kernel");
}
