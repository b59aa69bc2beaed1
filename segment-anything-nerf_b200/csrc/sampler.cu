// Proposal sampling chain, fused (sm_100a).
//
// Replaces the ~25 small torch kernels per sampling level of nerf/renderer.py:232-286 + 84-119
// (near/far slab test, spacing functions, linspace + jitter or inverse-CDF resampling, bin edges ->
// mid points / interval lengths, ray points, scene contraction, mapping to the grid's unit cube)
// with ONE kernel per level.  Outputs are exactly the tensors the rest of the path consumes:
//
//   bins    [N, T+1]  normalised bin edges in [0,1]        (next level's PDF support, losses)
//   t_mid   [N, T]    metric distance of the sample          (depth compositing)
//   deltas  [N, T]    metric interval length                 (sigma -> alpha)
//   x01     [N, T, 3] contracted position mapped to [0,1]^3  (grid encoder input, grid.py:156)
//
// Random jitter comes from a caller-provided uniform tensor (torch's generator), so parity tests
// can share the draw with the oracle; NULL = no perturbation.
#include "composite_common.cuh"

namespace sanerf {

struct RayFrame {
    float ox, oy, oz, dx, dy, dz;
    float s_near, s_far;
};

// near_far_from_aabb (renderer.py:122-139) + spacing_fn (renderer.py:250)
__device__ __forceinline__ RayFrame ray_frame(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                              const float* __restrict__ aabb, float min_near,
                                              const float* __restrict__ cam_near_far, uint32_t cnf_stride,
                                              uint32_t r) {
    RayFrame f;
    f.ox = __ldg(rays_o + 3 * (size_t)r); f.oy = __ldg(rays_o + 3 * (size_t)r + 1); f.oz = __ldg(rays_o + 3 * (size_t)r + 2);
    f.dx = __ldg(rays_d + 3 * (size_t)r); f.dy = __ldg(rays_d + 3 * (size_t)r + 1); f.dz = __ldg(rays_d + 3 * (size_t)r + 2);
    const float o[3] = {f.ox, f.oy, f.oz}, d[3] = {f.dx, f.dy, f.dz};
    float near = -INFINITY, far = INFINITY;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float inv = d[k] + 1e-15f;
        const float t0 = __fdiv_rn(__ldg(aabb + k) - o[k], inv), t1 = __fdiv_rn(__ldg(aabb + 3 + k) - o[k], inv);
        near = fmaxf(near, (t0 < t1) ? t0 : t1);
        far = fminf(far, (t0 > t1) ? t0 : t1);
    }
    if (far < near) { near = 1e9f; far = 1e9f; }
    near = fmaxf(near, min_near);
    if (cam_near_far != nullptr) {
        near = fmaxf(near, __ldg(cam_near_far + (size_t)r * cnf_stride));
        far = fminf(far, __ldg(cam_near_far + (size_t)r * cnf_stride + 1));
    }
    f.s_near = (near < 1.0f) ? near * 0.5f : 1.0f - __fdiv_rn(1.0f, 2.0f * near);
    f.s_far = (far < 1.0f) ? far * 0.5f : 1.0f - __fdiv_rn(1.0f, 2.0f * far);
    return f;
}

// normalised edge in [0,1] -> metric distance (renderer.py:252-253, :278)
__device__ __forceinline__ float edge_distance(const RayFrame& f, float b) {
    const float s = __fadd_rn(__fmul_rn(f.s_near, 1.0f - b), __fmul_rn(f.s_far, b));
    return (s < 0.5f) ? 2.0f * s : __fdiv_rn(1.0f, 2.0f - 2.0f * s);
}

// mid point, interval, position, contraction (renderer.py:60-69), unit-cube mapping (grid.py:156)
__device__ __forceinline__ void emit_sample(const RayFrame& f, float e0, float e1, int contract, float bound,
                                            float* __restrict__ t_mid, float* __restrict__ deltas,
                                            float* __restrict__ x01, size_t idx) {
    const float t = __fmul_rn(__fadd_rn(e1, e0), 0.5f);
    t_mid[idx] = t;
    deltas[idx] = e1 - e0;
    float p[3] = {__fadd_rn(f.ox, __fmul_rn(f.dx, t)), __fadd_rn(f.oy, __fmul_rn(f.dy, t)),
                  __fadd_rn(f.oz, __fmul_rn(f.dz, t))};
    if (contract) {
        const float ax = fabsf(p[0]), ay = fabsf(p[1]), az = fabsf(p[2]);
        float mag = ax; int dom = 0;              // first maximal axis, like torch.max(dim)
        if (ay > mag) { mag = ay; dom = 1; }
        if (az > mag) { mag = az; dom = 2; }
        if (!(mag < 1.0f)) {
            const float inv = __fdiv_rn(1.0f, mag);
            const float sd = __fdiv_rn(2.0f - inv, mag);
#pragma unroll
            for (int k = 0; k < 3; ++k) p[k] = __fmul_rn(p[k], (k == dom) ? sd : inv);
        }
    }
    const float two_b = 2.0f * bound;
#pragma unroll
    for (int k = 0; k < 3; ++k) x01[3 * idx + k] = __fdiv_rn(__fadd_rn(p[k], bound), two_b);
}

// torch.linspace(start, end, steps)[j] as ATen evaluates it on CUDA (symmetric about the middle)
__device__ __forceinline__ float linspace_at(float start, float end, uint32_t steps, uint32_t j) {
    const float step = (end - start) / (float)(steps - 1u);
    return (j < steps / 2u) ? start + step * (float)j : end - step * (float)(steps - 1u - j);
}

// Level 0: uniform (optionally jittered) edges, renderer.py:263-271.
__global__ void __launch_bounds__(256) sample_uniform_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ aabb,
    float min_near, const float* __restrict__ cam_near_far, uint32_t cnf_stride, const float* __restrict__ noise,
    uint32_t N, uint32_t T, int contract, float bound, float* __restrict__ bins, float* __restrict__ t_mid,
    float* __restrict__ deltas, float* __restrict__ x01) {
    pdl_begin();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * T) return;
    const uint32_t r = (uint32_t)(i / T), j = (uint32_t)(i - (size_t)r * T);
    const RayFrame f = ray_frame(rays_o, rays_d, aabb, min_near, cam_near_far, cnf_stride, r);
    float b0 = linspace_at(0.0f, 1.0f, T + 1u, j), b1 = linspace_at(0.0f, 1.0f, T + 1u, j + 1u);
    if (noise != nullptr) {
        const float* u = noise + (size_t)r * (T + 1u);
        b0 = fminf(fmaxf(b0 + __fdiv_rn(__ldg(u + j) - 0.5f, (float)T), 0.0f), 1.0f);
        b1 = fminf(fmaxf(b1 + __fdiv_rn(__ldg(u + j + 1u) - 0.5f, (float)T), 0.0f), 1.0f);
    }
    float* brow = bins + (size_t)r * (T + 1u);
    brow[j] = b0;
    if (j == T - 1u) brow[T] = b1;
    emit_sample(f, edge_distance(f, b0), edge_distance(f, b1), contract, bound, t_mid, deltas, x01, i);
}

// Levels >= 1: inverse-CDF resampling of T+1 edges from the previous level's weights
// (sample_pdf, renderer.py:84-119), one warp per ray.  Dynamic shared memory per warp:
// cdf[T0+1] | prev_bins[T0+1] | new_bins[T+1].
constexpr int kSamplerWarps = 4;

__global__ void __launch_bounds__(32 * kSamplerWarps) sample_pdf_kernel(
    const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ aabb,
    float min_near, const float* __restrict__ cam_near_far, uint32_t cnf_stride,
    const float* __restrict__ prev_bins, const float* __restrict__ prev_weights, uint32_t T0,
    const float* __restrict__ noise, uint32_t N, uint32_t T, int contract, float bound,
    float* __restrict__ bins, float* __restrict__ t_mid, float* __restrict__ deltas, float* __restrict__ x01,
    const float* __restrict__ prev_sigmas, const float* __restrict__ prev_deltas, int last_opaque,
    float* __restrict__ prev_weights_out) {
    pdl_begin();
    extern __shared__ float smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * kSamplerWarps + warp;
    if (r >= N) return;
    const uint32_t per_warp = (T0 + 1u) * 2u + (T + 1u) + T0;
    float* cdf = smem + (size_t)warp * per_warp;
    float* pbin = cdf + (T0 + 1u);
    float* nbin = pbin + (T0 + 1u);
    float* wbuf = nbin + (T + 1u);                   // the previous level's weights of this ray
    const uint32_t E = T + 1u;                       // edges to draw

    if (prev_sigmas != nullptr) {
        // fused compositing of the previous level (renderer.py:309-326, the C = 0 case of csrc/composite.cu): weights
        // from sigma and interval lengths, written out for the losses / the backward and kept here for the resampling
        CompositeArgs a{prev_sigmas, prev_deltas, nullptr, nullptr, nullptr, N, T0, 0u, last_opaque, 0.0f, 0u, 0u};
        float carry_x = 0.0f;
        for (uint32_t base = 0; base < T0; base += 32) {
            const SampleTerms st = chunk_terms(a, (size_t)r * T0, T0, base, lane, carry_x);
            if (st.valid) {
                wbuf[base + lane] = st.w;
                prev_weights_out[(size_t)r * T0 + base + lane] = st.w;
            }
        }
    } else {
        for (uint32_t i = lane; i < T0; i += 32) wbuf[i] = __ldg(prev_weights + (size_t)r * T0 + i);
    }
    __syncwarp();
    // pdf = (w + 0.01) / sum ; cdf = min(cumsum, 1) with a leading 0
    const float* w = wbuf;
    float total = 0.0f;
    for (uint32_t i = lane; i < T0; i += 32) total += w[i] + 0.01f;
    total = warp_sum(total);
    float carry = 0.0f;
    for (uint32_t base = 0; base < T0; base += 32) {
        const uint32_t i = base + lane;
        float p = (i < T0) ? __fdiv_rn(w[i] + 0.01f, total) : 0.0f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float v = __shfl_up_sync(0xffffffffu, p, o);
            if (lane >= (uint32_t)o) p += v;
        }
        if (i < T0) cdf[i + 1u] = fminf(carry + p, 1.0f);
        carry += __shfl_sync(0xffffffffu, p, 31);
    }
    if (lane == 0) cdf[0] = 0.0f;
    for (uint32_t i = lane; i <= T0; i += 32) pbin[i] = __ldg(prev_bins + (size_t)r * (T0 + 1u) + i);
    __syncwarp();

    for (uint32_t j = lane; j < E; j += 32) {
        float u = linspace_at(0.5f / (float)E, 1.0f - 0.5f / (float)E, E, j);
        if (noise != nullptr) u = u + __fdiv_rn(__ldg(noise + (size_t)r * E + j) - 0.5f, (float)E);
        // searchsorted(cdf, u, right=True): first index with cdf[idx] > u
        uint32_t lo = 0, hi = T0 + 1u;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (cdf[mid] > u) hi = mid; else lo = mid + 1u;
        }
        const uint32_t below = (lo == 0u) ? 0u : min(lo - 1u, T0);
        const uint32_t above = min(lo, T0);
        const float c0 = cdf[below], c1 = cdf[above];
        float t = __fdiv_rn(u - c0, c1 - c0);
        if (isnan(t)) t = 0.0f;                       // torch.nan_to_num
        else if (isinf(t)) t = (t > 0.0f) ? 3.4028234663852886e38f : -3.4028234663852886e38f;
        t = fminf(fmaxf(t, 0.0f), 1.0f);
        const float b0 = pbin[below], b1 = pbin[above];
        const float b = __fadd_rn(b0, __fmul_rn(t, b1 - b0));
        nbin[j] = b;
        bins[(size_t)r * E + j] = b;
    }
    __syncwarp();

    const RayFrame f = ray_frame(rays_o, rays_d, aabb, min_near, cam_near_far, cnf_stride, r);
    for (uint32_t j = lane; j < T; j += 32)
        emit_sample(f, edge_distance(f, nbin[j]), edge_distance(f, nbin[j + 1u]), contract, bound, t_mid, deltas,
                    x01, (size_t)r * T + j);
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_sample_uniform(const float* rays_o, const float* rays_d, const float* aabb, float min_near,
                                     const float* cam_near_far, uint32_t cnf_stride, const float* noise, uint32_t N,
                                     uint32_t T, int contract, float bound, float* bins, float* t_mid, float* deltas,
                                     float* x01, void* stream) {
    if (N == 0 || T == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(rays_o); SANERF_REQUIRE_PTR(rays_d); SANERF_REQUIRE_PTR(aabb);
    SANERF_REQUIRE_PTR(bins); SANERF_REQUIRE_PTR(t_mid); SANERF_REQUIRE_PTR(deltas); SANERF_REQUIRE_PTR(x01);
    if (!(bound > 0.0f)) return fail(SANERF_ERR_INVALID_ARG, "sample_uniform: bound must be > 0");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t blocks = (uint32_t)div_up((size_t)N * T, (size_t)256);
    SANERF_LAUNCH(sample_uniform_kernel, blocks, 256, 0, st, rays_o, rays_d, aabb, min_near, cam_near_far, cnf_stride, noise, N, T,
                                                  contract, bound, bins, t_mid, deltas, x01);
    return check_launch("sample_uniform_kernel");
}

extern "C" int sanerf_sample_pdf(const float* rays_o, const float* rays_d, const float* aabb, float min_near,
                                 const float* cam_near_far, uint32_t cnf_stride, const float* prev_bins,
                                 const float* prev_weights, uint32_t T0, const float* noise, uint32_t N, uint32_t T,
                                 int contract, float bound, float* bins, float* t_mid, float* deltas, float* x01,
                                 const float* prev_sigmas, const float* prev_deltas, int last_sample_opaque,
                                 float* prev_weights_out, void* stream) {
    if (N == 0 || T == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(rays_o); SANERF_REQUIRE_PTR(rays_d); SANERF_REQUIRE_PTR(aabb);
    SANERF_REQUIRE_PTR(prev_bins);
    if (prev_sigmas != nullptr) {
        SANERF_REQUIRE_PTR(prev_deltas); SANERF_REQUIRE_PTR(prev_weights_out);
    } else {
        SANERF_REQUIRE_PTR(prev_weights);
    }
    SANERF_REQUIRE_PTR(bins); SANERF_REQUIRE_PTR(t_mid); SANERF_REQUIRE_PTR(deltas); SANERF_REQUIRE_PTR(x01);
    if (T0 == 0) return fail(SANERF_ERR_INVALID_ARG, "sample_pdf: previous level has no samples");
    if (!(bound > 0.0f)) return fail(SANERF_ERR_INVALID_ARG, "sample_pdf: bound must be > 0");
    const size_t smem = (size_t)kSamplerWarps * ((T0 + 1u) * 2u + (T + 1u) + T0) * sizeof(float);
    if (smem > 200 * 1024) return fail(SANERF_ERR_INVALID_ARG, "sample_pdf: T0/T too large for shared memory");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SANERF_LAUNCH(sample_pdf_kernel, div_up(N, (uint32_t)kSamplerWarps), 32 * kSamplerWarps, smem, st, 
        rays_o, rays_d, aabb, min_near, cam_near_far, cnf_stride, prev_bins, prev_weights, T0, noise, N, T, contract,
        bound, bins, t_mid, deltas, x01, prev_sigmas, prev_deltas, last_sample_opaque, prev_weights_out);
    return check_launch("sample_pdf_kernel");
}

// ---- ray generation ---------------------------------------------------------------------------------------------
// get_rays of the reference (nerf/utils.py:145-279) for the pixels `inds` (flat row * W + col; NULL = the whole image):
// pixel centres (+0.5), pinhole directions ((i - cx) / fx, -(j - cy) / fy, -1), NOT normalised, rotated by the
// camera-to-world pose; origin = the pose's translation.  One pose / intrinsics for all rays (stride 0) or one per ray.
namespace sanerf {
__global__ void __launch_bounds__(256) generate_rays_kernel(const float* __restrict__ poses, uint32_t pose_stride,
                                                            const float* __restrict__ intrinsics, uint32_t intr_stride,
                                                            const int64_t* __restrict__ inds, uint32_t W, uint32_t N,
                                                            float* __restrict__ rays_o, float* __restrict__ rays_d) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int64_t pix = inds ? inds[n] : (int64_t)n;
    const float i = (float)(pix % W) + 0.5f, j = (float)(pix / W) + 0.5f;
    const float* P = poses + (size_t)n * pose_stride;           // row-major 4x4
    const float* K = intrinsics + (size_t)n * intr_stride;      // fx, fy, cx, cy
    const float xs = __fdiv_rn(i - __ldg(K + 2), __ldg(K + 0));
    const float ys = -__fdiv_rn(j - __ldg(K + 3), __ldg(K + 1));
    const float zs = -1.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float d = __fmaf_rn(zs, __ldg(P + 4 * k + 2), __fmaf_rn(ys, __ldg(P + 4 * k + 1), xs * __ldg(P + 4 * k)));
        rays_d[(size_t)n * 3 + k] = d;
        rays_o[(size_t)n * 3 + k] = __ldg(P + 4 * k + 3);
    }
}
}  // namespace sanerf

extern "C" int sanerf_generate_rays(const float* poses, uint32_t pose_stride, const float* intrinsics, uint32_t intr_stride,
                                    const int64_t* inds, uint32_t W, uint32_t N, float* rays_o, float* rays_d, void* stream) {
    if (N == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(poses); SANERF_REQUIRE_PTR(intrinsics); SANERF_REQUIRE_PTR(rays_o); SANERF_REQUIRE_PTR(rays_d);
    if (W == 0) return sanerf::fail(SANERF_ERR_INVALID_ARG, "generate_rays: W must be > 0");
    if ((pose_stride != 0 && pose_stride != 16) || (intr_stride != 0 && intr_stride != 4))
        return sanerf::fail(SANERF_ERR_INVALID_ARG, "generate_rays: pose stride 0 or 16 floats, intrinsics stride 0 or 4");
    sanerf::generate_rays_kernel<<<sanerf::div_up(N, 256u), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        poses, pose_stride, intrinsics, intr_stride, inds, W, N, rays_o, rays_d);
    return sanerf::check_launch("generate_rays_kernel");
}
