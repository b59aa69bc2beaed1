// Final-level density activation + compositing of the 16-wide field-head output, ONE kernel per direction (sm_100a).
//
// The field head returns [N, T, 16]: column 0 is the density logit, columns 1..15 the geometry feature
// (nerf/network.py:226-227).  The reference then runs trunc_exp (activation.py:5-18) and the sigma -> alpha ->
// transmittance -> weights -> weighted-sum chain of nerf/renderer.py:309-338.  Here a warp owns a ray and a LANE owns a
// SAMPLE: the lane reads its whole 64-byte row with four 16-byte loads (a warp instruction covers 32 consecutive rows),
// takes sigma = exp(row[0]), joins the warp scan of delta*sigma, and the 15 weighted channel sums are reduced across
// lanes with a 31-shuffle butterfly.  The backward needs no cross-lane reduction for the per-sample dot product
// g_out . feats_i at all, and writes the 16-wide gradient row (logit gradient through trunc_exp's clamped derivative
// + w_i * g_out) with four 16-byte stores.  Same arithmetic as csrc/composite.cu (+ encoders_misc.cu trunc_exp), which
// remain the general path (any C, packed rays).
#include <cfloat>

#include "common.cuh"

namespace sanerf {

namespace hc {
constexpr uint32_t kFull = 0xffffffffu;
constexpr int kRaysPerBlock = 4;
constexpr uint32_t kW = 16;         // row width: 1 logit + 15 features

struct Args {
    const float* head;      // [N*T, 16]
    const float* deltas;    // [N*T]
    const float* ts;        // [N*T]
    uint32_t N, T;
    int last_opaque;
    float t_thresh;
};

struct Terms {
    float x, T, w, sigma;
    bool valid, alive, finite;
};

__device__ __forceinline__ void load_row(const float* head, size_t i, float (&f)[16]) {
    const float4* p = reinterpret_cast<const float4*>(head + i * kW);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = __ldg(p + q);
        f[4 * q] = v.x; f[4 * q + 1] = v.y; f[4 * q + 2] = v.z; f[4 * q + 3] = v.w;
    }
}

// weights of one 32-sample chunk (same arithmetic as composite.cu: chunk_terms); `carry` = sum of x over earlier chunks
__device__ __forceinline__ Terms chunk_terms(const Args& a, size_t start, uint32_t n, uint32_t base, uint32_t lane, float logit,
                                             float& carry) {
    Terms s;
    const uint32_t i = base + lane;
    s.valid = i < n;
    s.sigma = s.valid ? expf(logit) : 0.0f;                 // trunc_exp forward (activation.py:10)
    float x = 0.0f;
    if (s.valid) {
        x = __ldg(a.deltas + start + i) * s.sigma;
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
    if (isnan(w)) w = 0.0f;                                 // weights.nan_to_num_(0)  (renderer.py:326)
    else if (isinf(w)) w = copysignf(FLT_MAX, w);
    s.w = s.alive ? w : 0.0f;
    return s;
}

// 16 values per lane -> lane c (< 16) ends with the sum over all 32 lanes of v[c]
__device__ __forceinline__ float transpose_sum16(float (&v)[16], uint32_t lane) {
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] += __shfl_xor_sync(kFull, v[k], 16);
#pragma unroll
    for (int s = 8; s >= 1; s >>= 1) {
        const bool upper = (lane & (uint32_t)s) != 0u;
#pragma unroll
        for (int k = 0; k < s; ++k) {
            const float send = upper ? v[k] : v[k + s];
            const float keep = upper ? v[k + s] : v[k];
            v[k] = keep + __shfl_xor_sync(kFull, send, s);
        }
    }
    return v[0];
}
}  // namespace hc

// ---- chunked front-to-back compositing with early ray termination ---------------------------------------------------
// The frame renderer evaluates the final level in chunks of CL = 8 samples per ray, front to back (mlp_tc.cu:
// sanerf_field_head_forward_chunk), for the rays still alive only.  This kernel composites ONE chunk of those rays, carrying
// the optical depth, weights_sum, depth, the 15 channel sums and the alive-sample count in per-ray state arrays, and
// appends the rays whose transmittance after the chunk is still >= t_thresh to the work list of the next chunk
// (warp-aggregated append; the order of the list does not matter).  Eight lanes own a ray, a lane owns a sample.
// Same per-sample arithmetic and the same termination rule as head_composite_forward_kernel: sample k contributes iff
// T_k = exp(-sum_{j<k} delta_j sigma_j) >= t_thresh, so the result equals the un-chunked kernel's up to the
// association order of the prefix sum; what chunking adds is that the FIELD is not evaluated behind the termination point.
struct ChunkArgs {
    const float* head;          // [N*T, 16]
    const float* deltas;        // [N*T]
    const float* ts;            // [N*T]
    const uint32_t* ray_list;   // rays of this chunk
    const uint32_t* list_count;
    uint32_t* next_list;        // or NULL for the last chunk
    uint32_t* next_count;
    float* optical;             // [N] carried sum of delta * sigma
    float* weights_sum;         // [N]
    float* depth;               // [N]
    float* out;                 // [N, 15]
    int32_t* n_alive;           // [N]
    uint32_t T, chunk, chunk_len;
    int last_opaque;
    float t_thresh;
};

__global__ void __launch_bounds__(128) head_composite_chunk_kernel(const ChunkArgs a) {
    pdl_begin();
    constexpr uint32_t kFull = 0xffffffffu, CL = 8;
    const uint32_t lane = threadIdx.x & 31u, sub = lane >> 3, j = lane & 7u;
    const uint32_t count = __ldg(a.list_count);
    const uint32_t q = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4u + sub;     // index into the work list
    const bool have = q < count;
    const uint32_t r = have ? __ldg(a.ray_list + q) : 0u;
    const uint32_t k = a.chunk * CL + j;                           // sample index within the ray
    const size_t i = (size_t)r * a.T + k;
    float f[16];
    if (have) hc::load_row(a.head, i, f);
    else {
#pragma unroll
        for (int c = 0; c < 16; ++c) f[c] = 0.0f;
    }
    const float sigma = have ? expf(f[0]) : 0.0f;
    float x = have ? __ldg(a.deltas + i) * sigma : 0.0f;
    if (have && a.last_opaque && k == a.T - 1u) x = INFINITY;
    float incl = x;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const float v = __shfl_up_sync(kFull, incl, o, 8);
        if (j >= (uint32_t)o) incl += v;
    }
    float excl = __shfl_up_sync(kFull, incl, 1, 8);
    if (j == 0) excl = 0.0f;
    const float carry = have ? a.optical[r] : 0.0f;
    const float S = carry + excl;
    const float total = carry + __shfl_sync(kFull, incl, 7, 8);   // optical depth after this chunk
    const float Tk = expf(-S);
    float w = (1.0f - expf(-x)) * Tk;
    const bool alive = have && !(Tk < a.t_thresh);
    if (isnan(w)) w = 0.0f;
    else if (isinf(w)) w = copysignf(FLT_MAX, w);
    w = alive ? w : 0.0f;
    float ws = w, dep = have ? w * __ldg(a.ts + i) : 0.0f;
    float acc[15];
#pragma unroll
    for (int c = 0; c < 15; ++c) acc[c] = w * f[c + 1];
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) {
        ws += __shfl_xor_sync(kFull, ws, o, 8);
        dep += __shfl_xor_sync(kFull, dep, o, 8);
#pragma unroll
        for (int c = 0; c < 15; ++c) acc[c] += __shfl_xor_sync(kFull, acc[c], o, 8);
    }
    const uint32_t alive_mask = __ballot_sync(kFull, alive);
    const int n_alive_chunk = __popc((alive_mask >> (sub * 8u)) & 0xffu);
    // ray survives into the next chunk iff its transmittance there is still above the threshold
    const bool more = have && (a.next_list != nullptr) && !(expf(-total) < a.t_thresh);
    if (have) {
        if (j == 0) {
            a.optical[r] = total;
            a.weights_sum[r] += ws;
            a.depth[r] += dep;
            a.n_alive[r] += n_alive_chunk;
        }
        // 15 channel sums: lanes j = 0..7 write two channels each (all lanes of the segment hold the totals)
#pragma unroll
        for (int c = 0; c < 15; ++c)
            if ((uint32_t)(c >> 1) == j) a.out[(size_t)r * 15 + c] += acc[c];
    }
    const uint32_t more_mask = __ballot_sync(kFull, more && j == 0);
    if (more_mask != 0u) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.next_count, (uint32_t)__popc(more_mask));
        base = __shfl_sync(kFull, base, 0);
        if (more && j == 0) a.next_list[base + __popc(more_mask & ((1u << lane) - 1u))] = r;
    }
}

__global__ void __launch_bounds__(32 * hc::kRaysPerBlock) head_composite_forward_kernel(
    const hc::Args a, float* __restrict__ sigma, float* __restrict__ weights, float* __restrict__ weights_sum,
    float* __restrict__ depth, float* __restrict__ out, int32_t* __restrict__ n_alive) {
    pdl_begin();
    using namespace hc;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t r = blockIdx.x * kRaysPerBlock + (threadIdx.x >> 5);
    if (r >= a.N) return;
    const size_t start = (size_t)r * a.T;
    const uint32_t n = a.T;
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.0f;
    float ws = 0.0f, dep = 0.0f, carry = 0.0f;
    int alive = 0;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t i = base + lane;
        float f[16];
        if (i < n) load_row(a.head, start + i, f);
        else {
#pragma unroll
            for (int c = 0; c < 16; ++c) f[c] = 0.0f;
        }
        const Terms s = chunk_terms(a, start, n, base, lane, f[0], carry);
        alive += __popc(__ballot_sync(kFull, s.alive));
        if (s.valid) {
            if (sigma) sigma[start + i] = s.sigma;
            weights[start + i] = s.w;
            ws += s.w;
            dep = __fmaf_rn(s.w, __ldg(a.ts + start + i), dep);
#pragma unroll
            for (int c = 1; c < 16; ++c) acc[c] = __fmaf_rn(s.w, f[c], acc[c]);
        }
    }
    ws = warp_sum(ws);
    dep = warp_sum(dep);
    if (lane == 0) {
        weights_sum[r] = ws;
        depth[r] = dep;
        if (n_alive) n_alive[r] = alive;
    }
    const float total = transpose_sum16(acc, lane);          // lane c holds channel c (c = 1..15)
    if (lane >= 1 && lane < 16) out[(size_t)r * 15 + (lane - 1)] = total;
}

// Backward.  g_i = g_out . feats_i + g_depth t_i + g_ws + g_w_i ;  dL/dx_i = g_i T_{i+1} - sum_{j>i} g_j w_j ;
// dL/dsigma_i = delta_i dL/dx_i (+ a direct gradient on sigma, if any), dL/dlogit_i = dL/dsigma_i * exp(clamp(logit_i, -15, 15)),
// dL/dfeats_i = w_i g_out.  Two sweeps over the ray's chunks (total of g_j w_j first), rows re-read from L1/L2.
__global__ void __launch_bounds__(32 * hc::kRaysPerBlock) head_composite_backward_kernel(
    const hc::Args a, const float* __restrict__ g_weights, const float* __restrict__ g_weights_sum,
    const float* __restrict__ g_depth, const float* __restrict__ g_out, const float* __restrict__ g_sigma_direct,
    float* __restrict__ g_head) {
    pdl_begin();
    using namespace hc;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t r = blockIdx.x * kRaysPerBlock + (threadIdx.x >> 5);
    if (r >= a.N) return;
    const size_t start = (size_t)r * a.T;
    const uint32_t n = a.T;
    const float g_ws = g_weights_sum ? __ldg(g_weights_sum + r) : 0.0f;
    const float g_dp = g_depth ? __ldg(g_depth + r) : 0.0f;
    float go[16];
    go[0] = 0.0f;
#pragma unroll
    for (int c = 1; c < 16; ++c) go[c] = g_out ? __ldg(g_out + (size_t)r * 15 + (c - 1)) : 0.0f;

    auto sample_grad = [&](const Terms& s, const float (&f)[16], uint32_t i) {
        float gi = 0.0f;
#pragma unroll
        for (int c = 1; c < 16; ++c) gi = __fmaf_rn(go[c], f[c], gi);
        gi += g_ws;
        gi = __fmaf_rn(g_dp, __ldg(a.ts + start + i), gi);
        if (g_weights) gi += __ldg(g_weights + start + i);
        if (!s.alive || !s.finite) gi = 0.0f;               // weight was forced to 0 / clamped: no gradient path
        return gi;
    };
    // sweep 1: total of g_j w_j over the ray
    float total = 0.0f, carry = 0.0f;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t i = base + lane;
        float f[16];
        if (i < n) load_row(a.head, start + i, f);
        else {
#pragma unroll
            for (int c = 0; c < 16; ++c) f[c] = 0.0f;
        }
        const Terms s = chunk_terms(a, start, n, base, lane, f[0], carry);
        if (s.valid) total += sample_grad(s, f, i) * s.w;
    }
    total = warp_sum(total);
    // sweep 2: suffix = total - inclusive prefix
    float prefix_carry = 0.0f;
    carry = 0.0f;
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t i = base + lane;
        float f[16];
        if (i < n) load_row(a.head, start + i, f);
        else {
#pragma unroll
            for (int c = 0; c < 16; ++c) f[c] = 0.0f;
        }
        const Terms s = chunk_terms(a, start, n, base, lane, f[0], carry);
        const float gi = s.valid ? sample_grad(s, f, i) : 0.0f;
        float incl = gi * s.w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float v = __shfl_up_sync(kFull, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        const float sfx = total - (prefix_carry + incl);
        prefix_carry += __shfl_sync(kFull, incl, 31);
        if (s.valid) {
            const float Tnext = s.T * expf(-s.x);
            float ds = __ldg(a.deltas + start + i) * (gi * Tnext - sfx);
            if (a.last_opaque && i == n - 1u) ds = 0.0f;    // x := inf is a constant (renderer.py:315-316)
            if (g_sigma_direct) ds += __ldg(g_sigma_direct + start + i);
            const float dlogit = ds * expf(fminf(fmaxf(f[0], -15.0f), 15.0f));   // trunc_exp backward (activation.py:16)
            float4* dst = reinterpret_cast<float4*>(g_head + (start + i) * kW);
            dst[0] = make_float4(dlogit, s.w * go[1], s.w * go[2], s.w * go[3]);
#pragma unroll
            for (int q = 1; q < 4; ++q)
                dst[q] = make_float4(s.w * go[4 * q], s.w * go[4 * q + 1], s.w * go[4 * q + 2], s.w * go[4 * q + 3]);
        }
    }
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_head_composite_forward(const float* head, const float* deltas, const float* ts, uint32_t N, uint32_t T,
                                             int last_sample_opaque, float t_thresh, float* sigma, float* weights,
                                             float* weights_sum, float* depth, float* out, int32_t* n_alive, void* stream) {
    if (N == 0 || T == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(head); SANERF_REQUIRE_PTR(deltas); SANERF_REQUIRE_PTR(ts);
    SANERF_REQUIRE_PTR(weights); SANERF_REQUIRE_PTR(weights_sum); SANERF_REQUIRE_PTR(depth); SANERF_REQUIRE_PTR(out);
    if ((uintptr_t)head & 15u) return fail(SANERF_ERR_MISALIGNED, "head_composite: head must be 16-byte aligned");
    hc::Args a{head, deltas, ts, N, T, last_sample_opaque, t_thresh};
    SANERF_LAUNCH(head_composite_forward_kernel, div_up(N, (uint32_t)hc::kRaysPerBlock), 32 * hc::kRaysPerBlock, 0, static_cast<cudaStream_t>(stream), a, sigma, weights, weights_sum, depth, out, n_alive);
    return check_launch("head_composite_forward_kernel");
}

extern "C" int sanerf_head_composite_backward(const float* head, const float* deltas, const float* ts, uint32_t N, uint32_t T,
                                              int last_sample_opaque, float t_thresh, const float* g_weights,
                                              const float* g_weights_sum, const float* g_depth, const float* g_out,
                                              const float* g_sigma_direct, float* g_head, void* stream) {
    if (N == 0 || T == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(head); SANERF_REQUIRE_PTR(deltas); SANERF_REQUIRE_PTR(ts); SANERF_REQUIRE_PTR(g_head);
    if (((uintptr_t)head | (uintptr_t)g_head) & 15u) return fail(SANERF_ERR_MISALIGNED, "head_composite: 16-byte alignment");
    hc::Args a{head, deltas, ts, N, T, last_sample_opaque, t_thresh};
    SANERF_LAUNCH(head_composite_backward_kernel, div_up(N, (uint32_t)hc::kRaysPerBlock), 32 * hc::kRaysPerBlock, 0, static_cast<cudaStream_t>(stream), a, g_weights, g_weights_sum, g_depth, g_out,
                                                                          g_sigma_direct, g_head);
    return check_launch("head_composite_backward_kernel");
}

extern "C" int sanerf_head_composite_chunk(const float* head, const float* deltas, const float* ts, const uint32_t* ray_list,
                                           const uint32_t* list_count, uint32_t max_rays, uint32_t T, uint32_t chunk,
                                           uint32_t chunk_len, int last_sample_opaque, float t_thresh, uint32_t* next_list,
                                           uint32_t* next_count, float* optical, float* weights_sum, float* depth, float* out,
                                           int32_t* n_alive, void* stream) {
    if (max_rays == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(head); SANERF_REQUIRE_PTR(deltas); SANERF_REQUIRE_PTR(ts); SANERF_REQUIRE_PTR(ray_list);
    SANERF_REQUIRE_PTR(list_count); SANERF_REQUIRE_PTR(optical); SANERF_REQUIRE_PTR(weights_sum); SANERF_REQUIRE_PTR(depth);
    SANERF_REQUIRE_PTR(out); SANERF_REQUIRE_PTR(n_alive);
    if (chunk_len != 8 || T % 8 != 0 || (chunk + 1) * chunk_len > T)
        return fail(SANERF_ERR_INVALID_ARG, "head_composite_chunk: chunk_len must be 8, T a multiple of 8, the chunk inside the ray");
    if ((next_list == nullptr) != (next_count == nullptr)) return fail(SANERF_ERR_INVALID_ARG, "next_list and next_count go together");
    ChunkArgs a{head, deltas, ts, ray_list, list_count, next_list, next_count, optical, weights_sum, depth, out, n_alive,
                T, chunk, chunk_len, last_sample_opaque, t_thresh};
    SANERF_LAUNCH(head_composite_chunk_kernel, div_up(max_rays, 16u), 128, 0, static_cast<cudaStream_t>(stream), a);
    return check_launch("head_composite_chunk_kernel");
}
