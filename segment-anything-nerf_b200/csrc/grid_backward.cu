// Multiresolution hash-grid encoder: parameter-gradient scatter, input gradient,
// total-variation and weight-decay gradient kernels (sm_100a).
//
// Replaces kernel_grid_backward / kernel_input_backward / kernel_grad_tv / kernel_grad_wd of
// the reference (gridencoder/src/gridencoder.cu:252-378, 525-713).
//
// Scatter design (different from the reference's "one thread per channel pair, two scalar
// atomics per corner"):
//  * one thread owns (sample, level-group) and ALL C channels: the cell is located once, not
//    C/2 times;
//  * every corner update is ONE vector reduction into L2 — red.global.add.v2.f32 (C=2),
//    .v4.f32 (C>=4), .v{1,2,4}.f16x2 for half tables — i.e. 2x..4x fewer L2 atomic operations
//    than scalar atomicAdd;
//  * the incoming gradient is read straight from the [B, L*C] layout autograd hands over
//    (the reference first materialises a [L,B,C] permuted copy, grid.py:80);
//  * levels whose gradient row is exactly zero are skipped.
#include "grid_common.cuh"

namespace sanerf {

struct GridBwdParams {
    const void* grad;
    const float* inputs;
    const int32_t* offsets;
    void* grad_table;
    uint32_t B, L, max_level, H;
    float S;
    uint32_t gridtype, interp;
    int align_corners, blc;
};

template <typename T, uint32_t D, uint32_t C, uint32_t G, uint32_t CH>
__global__ void __launch_bounds__(256) grid_backward_kernel(const GridBwdParams p) {
    pdl_begin();
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t level0 = blockIdx.y * G;
    // Tables with few values per cell: merge consecutive samples of a ray that share a cell before the reductions are
    // issued (grid_common.cuh: warp_run_reduce; sums are kept in fp32 and, for half tables, rounded ONCE per run instead
    // of once per sample).  Needs warp-uniform control flow: no early exits.  Without it the 4096 rows of level 0 take
    // every sample's eight updates one by one: 360 us of the 977 us cfg5 scatter (T = 2^22 fp16, 2^20 samples).
    constexpr bool kAggregate = ((1u << D) * CH <= 16u) && (CH == C);
    const uint32_t lane = threadIdx.x & 31u;
    // geometry of this CTA's G levels, once per CTA (evaluated per thread it was ~200 of a warp's ~1450 instructions in a kernel
    // whose warps wait for an issue slot a third of the time, and put two dependent offset loads in front of everything)
    __shared__ LevelGeom<D> s_geo[G];
    __shared__ uint32_t s_base[G];
    if (threadIdx.x < G && level0 + threadIdx.x < p.L) {
        s_geo[threadIdx.x] = level_geometry<D>(p.offsets, level0 + threadIdx.x, p.S, p.H, p.gridtype);
        s_base[threadIdx.x] = (uint32_t)__ldg(p.offsets + level0 + threadIdx.x);
    }
    __syncthreads();

    float x[D];
    bool ok = b < p.B;
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) x[d] = ok ? __ldg(p.inputs + (size_t)b * D + d) : 0.5f;
    ok = ok && !out_of_range<D>(x);  // gridencoder.cu:279-284: gradient of OOB samples is dropped
    if constexpr (!kAggregate) {
        if (!ok) return;
    }

    const T* __restrict__ grad = static_cast<const T*>(p.grad);
    T* __restrict__ gtab = static_cast<T*>(p.grad_table);

#pragma unroll
    for (uint32_t g = 0; g < G; ++g) {
        const uint32_t level = level0 + g;
        if (level >= p.max_level) break;
        const T* src = p.blc ? grad + ((size_t)b * p.L + level) * C
                             : grad + ((size_t)level * p.B + b) * C;
        const LevelGeom<D> geo = s_geo[g];
        const Cell<D> cell = locate<D>(geo, x, p.align_corners != 0, p.interp);
        T* __restrict__ slice = gtab + (size_t)s_base[g] * C;
        if constexpr (kAggregate) {
            float gv[CH];
#pragma unroll
            for (uint32_t c = 0; c < CH; ++c) gv[c] = 0.0f;
            if (ok) RowIO<T, CH>::load(src, gv);
            bool any = false;
#pragma unroll
            for (uint32_t c = 0; c < CH; ++c) any |= (gv[c] != 0.0f);
            const bool contributes = ok && any;
            float v[(1u << D) * CH];
#pragma unroll
            for (uint32_t k = 0; k < (1u << D); ++k) {
                const float w = contributes ? corner_weight<D>(cell, k) : 0.0f;
#pragma unroll
                for (uint32_t c = 0; c < CH; ++c) v[k * CH + c] = w * gv[c];
            }
            uint32_t key[D];
#pragma unroll
            for (uint32_t d = 0; d < D; ++d) key[d] = cell.lo[d];
            if (!contributes) key[0] = 0xffffffffu - lane;       // a key no cell has: its own (skipped) run
            const bool head = warp_run_reduce<(1u << D) * CH, D>(v, key, lane);
            if (head && contributes) {
                if constexpr (CH == 2) {
                    // The two corners along the first dimension are adjacent rows whenever the lower one is even
                    // (dense levels: stride 1; hashed levels: prime 1, so x ^ h and (x + 1) ^ h differ in bit 0 only):
                    // then both fit ONE 16-byte-aligned slot and go out as one red.global.add.v4.f32 - the scatter is
                    // bound by the number of reduction requests an SM can inject, not by their payload.
#pragma unroll
                    for (uint32_t k = 0; k < (1u << D); k += 2) {
                        const uint32_t r0 = corner_row<D>(geo, cell, k), r1 = corner_row<D>(geo, cell, k + 1);
                        if ((r0 ^ r1) == 1u) {
                            const bool lo_first = r0 < r1;
                            const float a0 = lo_first ? v[2 * k] : v[2 * k + 2], a1 = lo_first ? v[2 * k + 1] : v[2 * k + 3];
                            const float b0 = lo_first ? v[2 * k + 2] : v[2 * k], b1 = lo_first ? v[2 * k + 3] : v[2 * k + 1];
                            if constexpr (sizeof(T) == 4) {
                                red_add_v4_f32(reinterpret_cast<float*>(slice) + (size_t)(lo_first ? r0 : r1) * 2, a0, a1, b0, b1);
                            } else {                         // two 4-byte rows in one aligned 8-byte slot
                                red_add_v2_f16x2(reinterpret_cast<__half*>(slice) + (size_t)(lo_first ? r0 : r1) * 2,
                                                 __floats2half2_rn(a0, a1), __floats2half2_rn(b0, b1));
                            }
                        } else {
                            float u0[2] = {v[2 * k], v[2 * k + 1]}, u1[2] = {v[2 * k + 2], v[2 * k + 3]};
                            RowIO<T, 2>::red(slice + (size_t)r0 * 2, u0);
                            RowIO<T, 2>::red(slice + (size_t)r1 * 2, u1);
                        }
                    }
                } else {
#pragma unroll
                    for (uint32_t k = 0; k < (1u << D); ++k) {
                        float upd[CH];
#pragma unroll
                        for (uint32_t c = 0; c < CH; ++c) upd[c] = v[k * CH + c];
                        RowIO<T, CH>::red(slice + (size_t)corner_row<D>(geo, cell, k) * C, upd);
                    }
                }
            }
        } else {
#pragma unroll 1
            for (uint32_t c0 = 0; c0 < C; c0 += CH) {
                float gv[CH];
                RowIO<T, CH>::load(src + c0, gv);
                bool any = false;
#pragma unroll
                for (uint32_t c = 0; c < CH; ++c) any |= (gv[c] != 0.0f);
                if (!any) continue;
#pragma unroll
                for (uint32_t k = 0; k < (1u << D); ++k) {
                    const float w = corner_weight<D>(cell, k);
                    float upd[CH];
#pragma unroll
                    for (uint32_t c = 0; c < CH; ++c) upd[c] = w * gv[c];
                    RowIO<T, CH>::red(slice + (size_t)corner_row<D>(geo, cell, k) * C + c0, upd);
                }
            }
        }
    }
}

// grad_inputs[b,d] = sum_l sum_c grad[l,b,c] * dy_dx[b,l,d,c]   (gridencoder.cu:352-378)
template <typename T>
__global__ void grid_input_backward_kernel(const T* __restrict__ grad, const T* __restrict__ dy_dx,
                                           T* __restrict__ grad_inputs, uint32_t B, uint32_t D,
                                           uint32_t C, uint32_t L, int blc) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)B * D) return;
    const uint32_t b = (uint32_t)(t / D), d = (uint32_t)(t - (size_t)b * D);
    const T* jac = dy_dx + (size_t)b * L * D * C;
    float acc = 0.0f;
    for (uint32_t l = 0; l < L; ++l) {
        const T* g = blc ? grad + ((size_t)b * L + l) * C : grad + ((size_t)l * B + b) * C;
        for (uint32_t c = 0; c < C; ++c)
            acc = __fmaf_rn(to_float<T>(g[c]), to_float<T>(jac[((size_t)l * D + d) * C + c]), acc);
    }
    grad_inputs[t] = from_float<T>(acc);
}

// Total-variation gradient, in place on `grad` (gridencoder.cu:525-631).  The neighbour one
// step up is taken without clamping (cur < res always holds there) and is wrapped by the
// modulo — quirk kept.
template <typename T, uint32_t D>
__global__ void __launch_bounds__(256) grid_tv_kernel(const T* __restrict__ inputs, const T* __restrict__ table,
                                                      T* __restrict__ grad, const int32_t* __restrict__ offsets,
                                                      float weight, uint32_t B, uint32_t C, float S, uint32_t H,
                                                      uint32_t gridtype, int align_corners) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) x[d] = to_float<T>(inputs[(size_t)b * D + d]);
    if (out_of_range<D>(x)) return;

    const LevelGeom<D> geo = level_geometry<D>(offsets, level, S, H, gridtype);
    const size_t base = (size_t)(uint32_t)__ldg(offsets + level) * C;
    table += base;
    grad += base;

    uint32_t pg[D];
#pragma unroll
    for (uint32_t d = 0; d < D; ++d) {
        if (align_corners) {
            pg[d] = min((uint32_t)floorf(x[d] * (float)(geo.res - 1u)), geo.res - 2u);
        } else {
            const float pos = fminf(fmaxf(__fmaf_rn(x[d], (float)geo.res, -0.5f), 0.0f), (float)(geo.res - 1u));
            pg[d] = (uint32_t)floorf(pos);
        }
    }
    auto row_of = [&](const uint32_t (&q)[D]) -> size_t {
        uint32_t index = 0;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            const uint32_t term = q[d] * geo.mult[d];
            index = geo.hashed ? (index ^ term) : (index + term);
        }
        return (size_t)(index % geo.rows) * C;  // neighbours may leave the dense range: always wrap
    };
    const size_t centre = row_of(pg);
    const float w = weight / (float)(2u * D);

    for (uint32_t c = 0; c < C; ++c) {
        const float v0 = to_float<T>(table[centre + c]);
        float sum = 0.0f, sq = 0.0f;
#pragma unroll
        for (uint32_t d = 0; d < D; ++d) {
            const uint32_t cur = pg[d];
            if (cur < geo.res) {
                pg[d] = cur + 1u;
                const float gvd = v0 - to_float<T>(table[row_of(pg) + c]);
                sum += gvd;
                sq = __fmaf_rn(gvd, gvd, sq);
            }
            if (cur > 0u) {
                pg[d] = cur - 1u;
                const float gvd = v0 - to_float<T>(table[row_of(pg) + c]);
                sum += gvd;
                sq = __fmaf_rn(gvd, gvd, sq);
            }
            pg[d] = cur;
        }
        const float upd = (w * sum) * rsqrtf(sq + 1e-9f);
        if constexpr (sizeof(T) == 4) {
            red_add_f32(reinterpret_cast<float*>(grad + centre + c), upd);
        } else {
            atomicAdd(reinterpret_cast<__half*>(grad + centre + c), __float2half_rn(upd));
        }
    }
}

// Level-mean weight decay, in place: grad += 2*w*param/level_rows (gridencoder.cu:670-703).
template <typename T>
__global__ void __launch_bounds__(256) grid_wd_kernel(const T* __restrict__ table, T* __restrict__ grad,
                                                      const int32_t* __restrict__ offsets, float weight,
                                                      uint32_t rows, uint32_t C, uint32_t L) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * C) return;
    const uint32_t n = (uint32_t)(i / C);
    uint32_t level = 0;  // largest m < L with offsets[m] <= n (the reference binary-searches)
    for (uint32_t m = 1; m < L; ++m) level = ((uint32_t)__ldg(offsets + m) <= n) ? m : level;
    const float level_rows = (float)(uint32_t)(__ldg(offsets + level + 1) - __ldg(offsets + level));
    const float upd = ((2.0f * weight) * to_float<T>(table[i])) / level_rows;
    grad[i] = from_float<T>(to_float<T>(grad[i]) + upd);
}

// ---- dispatch --------------------------------------------------------------------------
template <typename T, uint32_t D, uint32_t C>
static int launch_backward(const GridBwdParams& p, cudaStream_t stream) {
    constexpr uint32_t CH = (C < 8u) ? C : 8u;
    constexpr uint32_t kBytes = C * sizeof(T);
    constexpr uint32_t kSector = (kBytes >= 32u) ? 1u : 32u / kBytes;
    constexpr uint32_t G = (kSector > 4u) ? 4u : kSector;
    constexpr uint32_t kThreads = 256;
    if (p.max_level == 0 || p.B == 0) return SANERF_OK;
    dim3 grid(div_up(p.B, kThreads), div_up(p.max_level, G), 1);
    SANERF_LAUNCH((grid_backward_kernel<T, D, C, G, CH>), grid, kThreads, 0, stream, p);
    return check_launch("grid_backward_kernel");
}

template <typename T, uint32_t D>
static int bwd_dispatch_C(const GridBwdParams& p, uint32_t C, cudaStream_t stream) {
    switch (C) {
        case 1: return launch_backward<T, D, 1>(p, stream);
        case 2: return launch_backward<T, D, 2>(p, stream);
        case 4: return launch_backward<T, D, 4>(p, stream);
        case 8: return launch_backward<T, D, 8>(p, stream);
        case 16: return launch_backward<T, D, 16>(p, stream);
        case 32: return launch_backward<T, D, 32>(p, stream);
        default: return fail(SANERF_ERR_INVALID_ARG, "GridEncoding: C must be 1, 2, 4, 8, 16 or 32.");
    }
}

template <typename T>
static int bwd_dispatch_D(const GridBwdParams& p, uint32_t D, uint32_t C, cudaStream_t stream) {
    switch (D) {
        case 2: return bwd_dispatch_C<T, 2>(p, C, stream);
        case 3: return bwd_dispatch_C<T, 3>(p, C, stream);
        case 4: return bwd_dispatch_C<T, 4>(p, C, stream);
        case 5: return bwd_dispatch_C<T, 5>(p, C, stream);
        default: return fail(SANERF_ERR_INVALID_ARG, "GridEncoding: D must be 2, 3, 4 or 5.");
    }
}

template <typename T>
static int launch_tv(const void* inputs, const void* table, void* grad, const int32_t* offsets,
                     float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                     uint32_t gridtype, int align_corners, cudaStream_t st) {
    dim3 grid(div_up(B, 256u), L, 1);
    const T* in = static_cast<const T*>(inputs);
    const T* tb = static_cast<const T*>(table);
    T* gr = static_cast<T*>(grad);
    switch (D) {
        case 2: grid_tv_kernel<T, 2><<<grid, 256, 0, st>>>(in, tb, gr, offsets, weight, B, C, S, H, gridtype, align_corners); break;
        case 3: grid_tv_kernel<T, 3><<<grid, 256, 0, st>>>(in, tb, gr, offsets, weight, B, C, S, H, gridtype, align_corners); break;
        case 4: grid_tv_kernel<T, 4><<<grid, 256, 0, st>>>(in, tb, gr, offsets, weight, B, C, S, H, gridtype, align_corners); break;
        case 5: grid_tv_kernel<T, 5><<<grid, 256, 0, st>>>(in, tb, gr, offsets, weight, B, C, S, H, gridtype, align_corners); break;
        default: return fail(SANERF_ERR_INVALID_ARG, "GridEncoding: D must be 2, 3, 4, or 5.");
    }
    return check_launch("grid_tv_kernel");
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                                           const int32_t* offsets, void* grad_embeddings, uint32_t B,
                                           uint32_t D, uint32_t C, uint32_t L, uint32_t max_level,
                                           float S, uint32_t H, const void* dy_dx, void* grad_inputs,
                                           uint32_t gridtype, int align_corners, uint32_t interp,
                                           int dtype, int grad_layout, void* stream) {
    (void)embeddings;  // the scatter does not read the table (the reference passes it but never uses it)
    if (B == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(grad);
    SANERF_REQUIRE_PTR(inputs);
    SANERF_REQUIRE_PTR(offsets);
    SANERF_REQUIRE_PTR(grad_embeddings);
    if (grad_layout != SANERF_LAYOUT_LBC && grad_layout != SANERF_LAYOUT_BLC)
        return fail(SANERF_ERR_INVALID_ARG, "grad_layout must be SANERF_LAYOUT_LBC or SANERF_LAYOUT_BLC");
    if (gridtype > 1u) return fail(SANERF_ERR_INVALID_ARG, "gridtype must be 0 (hash) or 1 (tiled)");
    if (interp > 1u) return fail(SANERF_ERR_INVALID_ARG, "interp must be 0 (linear) or 1 (smoothstep)");
    if ((dy_dx == nullptr) != (grad_inputs == nullptr))
        return fail(SANERF_ERR_INVALID_ARG, "dy_dx and grad_inputs must be given together");
    if (max_level > L) max_level = L;
    GridBwdParams p;
    p.grad = grad; p.inputs = inputs; p.offsets = offsets; p.grad_table = grad_embeddings;
    p.B = B; p.L = L; p.max_level = max_level; p.H = H; p.S = S;
    p.gridtype = gridtype; p.interp = interp; p.align_corners = align_corners;
    p.blc = (grad_layout == SANERF_LAYOUT_BLC);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    switch (dtype) {
        case SANERF_F32: rc = bwd_dispatch_D<float>(p, D, C, st); break;
        case SANERF_F16: rc = bwd_dispatch_D<__half>(p, D, C, st); break;
        default: return fail(SANERF_ERR_INVALID_ARG, "dtype must be SANERF_F32 or SANERF_F16");
    }
    if (rc != SANERF_OK) return rc;
    if (dy_dx != nullptr) {
        const uint32_t blocks = (uint32_t)div_up((size_t)B * D, (size_t)256);
        if (dtype == SANERF_F32)
            grid_input_backward_kernel<float><<<blocks, 256, 0, st>>>(
                static_cast<const float*>(grad), static_cast<const float*>(dy_dx),
                static_cast<float*>(grad_inputs), B, D, C, L, p.blc);
        else
            grid_input_backward_kernel<__half><<<blocks, 256, 0, st>>>(
                static_cast<const __half*>(grad), static_cast<const __half*>(dy_dx),
                static_cast<__half*>(grad_inputs), B, D, C, L, p.blc);
        return check_launch("grid_input_backward_kernel");
    }
    return SANERF_OK;
}

extern "C" int sanerf_grad_total_variation(const void* inputs, const void* embeddings, void* grad,
                                           const int32_t* offsets, float weight, uint32_t B, uint32_t D,
                                           uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                           int align_corners, int dtype, void* stream) {
    if (B == 0 || L == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(inputs);
    SANERF_REQUIRE_PTR(embeddings);
    SANERF_REQUIRE_PTR(grad);
    SANERF_REQUIRE_PTR(offsets);
    if (C != 1 && C != 2 && C != 4 && C != 8 && C != 16 && C != 32)
        return fail(SANERF_ERR_INVALID_ARG, "GridEncoding: C must be 1, 2, 4, 8, 16 or 32.");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case SANERF_F32: return launch_tv<float>(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners, st);
        case SANERF_F16: return launch_tv<__half>(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners, st);
        default: return fail(SANERF_ERR_INVALID_ARG, "dtype must be SANERF_F32 or SANERF_F16");
    }
}

extern "C" int sanerf_grad_weight_decay(const void* embeddings, void* grad, const int32_t* offsets,
                                        float weight, uint32_t B, uint32_t C, uint32_t L, int dtype,
                                        void* stream) {
    if (B == 0 || C == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(embeddings);
    SANERF_REQUIRE_PTR(grad);
    SANERF_REQUIRE_PTR(offsets);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t blocks = (uint32_t)div_up((size_t)B * C, (size_t)256);
    switch (dtype) {
        case SANERF_F32:
            grid_wd_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(embeddings), static_cast<float*>(grad), offsets, weight, B, C, L);
            break;
        case SANERF_F16:
            grid_wd_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<const __half*>(embeddings), static_cast<__half*>(grad), offsets, weight, B, C, L);
            break;
        default: return fail(SANERF_ERR_INVALID_ARG, "dtype must be SANERF_F32 or SANERF_F16");
    }
    return check_launch("grid_wd_kernel");
}
