// Per-ray transmittance / weight terms shared by the compositing kernels and the resampling kernel (sm_100a).
// Arithmetic of nerf/renderer.py:309-326; see composite.cu for the full description.
#pragma once

#include <cfloat>

#include "common.cuh"

namespace sanerf {

constexpr uint32_t kFull = 0xffffffffu;
constexpr int kRaysPerBlock = 4;   // one warp per ray
constexpr int kMaxChunks = 8;      // backward keeps per-chunk terms in registers: rays up to 256 samples

struct CompositeArgs {
    const float* sigmas;
    const float* deltas;
    const float* ts;
    const float* feats;
    const int32_t* ray_offsets;
    uint32_t N, T, C;
    int last_opaque;
    float t_thresh;
    uint32_t fs;    // feats row stride (floats)
    uint32_t gfs;   // grad_feats row stride (floats)
};

struct SampleTerms {
    float x, T, w;      // delta*sigma, incoming transmittance, final weight
    bool valid, alive, finite;
};

// weights of one 32-sample chunk; `carry` is sum of x over previous chunks (updated)
__device__ __forceinline__ SampleTerms chunk_terms(const CompositeArgs& a, size_t start, uint32_t n,
                                                   uint32_t base, uint32_t lane, float& carry) {
    SampleTerms s;
    const uint32_t i = base + lane;
    s.valid = i < n;
    float x = 0.0f;
    if (s.valid) {
        x = __ldg(a.deltas + start + i) * __ldg(a.sigmas + start + i);
        if (a.last_opaque && i == n - 1u) x = INFINITY;
    }
    float incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(kFull, incl, o);
        if (lane >= (uint32_t)o) incl += v;
    }
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 0.0f;
    const float S = carry + excl;
    carry += __shfl_sync(kFull, incl, 31);
    s.x = x;
    s.T = expf(-S);
    float w = (1.0f - expf(-x)) * s.T;
    s.alive = s.valid && !(s.T < a.t_thresh);
    s.finite = isfinite(w);
    if (isnan(w)) w = 0.0f;                          // weights.nan_to_num_(0)  (renderer.py:326)
    else if (isinf(w)) w = copysignf(FLT_MAX, w);
    s.w = s.alive ? w : 0.0f;
    return s;
}

}  // namespace sanerf
