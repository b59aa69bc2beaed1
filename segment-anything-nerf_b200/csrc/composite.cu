// Per-ray front-to-back compositing, forward and backward (sm_100a).
//
// Replaces the ~10 elementwise / cumsum / broadcast-multiply / reduction torch kernels of
// nerf/renderer.py:309-338, :377-383 (and the [N,T,C] temporaries autograd keeps for them)
// with ONE warp-per-ray pass in each direction:
//
//   x_i = delta_i * sigma_i  (last sample := +inf when the background is the opaque last sample)
//   T_i = exp(-sum_{j<i} x_j)           -- warp exclusive scan, carried across 32-sample chunks
//   w_i = (1 - exp(-x_i)) * T_i ,  NaN -> 0
//   weights_sum = sum w_i ; depth = sum w_i t_i ; out[c] = sum w_i feats[i,c]
//
// Samples are packed: ray r owns [ray_offsets[r], ray_offsets[r+1]) (dense [N,T] when
// ray_offsets == NULL).  Early ray termination: samples whose incoming transmittance is below
// t_thresh get weight 0 and their feature rows are never read.
//
// Channel reduction: C > 8 maps lanes to channels (each feature row is one coalesced read,
// accumulators stay in registers); C <= 8 maps lanes to samples and warp-reduces at the end.
//
// Backward:  g_i = g_out . feats_i + g_depth t_i + g_ws + g_w_i
//   dL/dx_i = g_i T_{i+1} - sum_{j>i} g_j w_j ,  dL/dsigma_i = delta_i dL/dx_i (0 for the opaque sample)
//   dL/dfeats_i = w_i g_out
// The per-sample dot products of a 32-sample chunk are reduced with a 31-shuffle butterfly
// (lane j ends with sample j's dot product); the suffix sums are a reverse warp scan over
// per-chunk registers.
#include "composite_common.cuh"

namespace sanerf {

template <int KC>  // KC = ceil(C/32) for the lane-per-channel path; 0 = lane-per-sample path (C <= 8)
__global__ void __launch_bounds__(32 * kRaysPerBlock) composite_forward_kernel(
    const CompositeArgs a, float* __restrict__ weights, float* __restrict__ weights_sum,
    float* __restrict__ depth, float* __restrict__ out, int32_t* __restrict__ n_alive) {
    pdl_begin();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t r = blockIdx.x * kRaysPerBlock + (threadIdx.x >> 5);
    if (r >= a.N) return;
    const size_t start = a.ray_offsets ? (size_t)a.ray_offsets[r] : (size_t)r * a.T;
    const uint32_t n = a.ray_offsets ? (uint32_t)(a.ray_offsets[r + 1] - a.ray_offsets[r]) : a.T;
    const uint32_t C = a.C;

    constexpr int NACC = (KC > 0) ? KC : 8;
    float acc[NACC];
#pragma unroll
    for (int q = 0; q < NACC; ++q) acc[q] = 0.0f;
    float ws = 0.0f, dep = 0.0f, carry = 0.0f;
    int alive = 0;

    for (uint32_t base = 0; base < n; base += 32) {
        const SampleTerms s = chunk_terms(a, start, n, base, lane, carry);
        const uint32_t alive_mask = __ballot_sync(kFull, s.alive);
        alive += __popc(alive_mask);
        if (s.valid) {
            weights[start + base + lane] = s.w;
            ws += s.w;
            dep = __fmaf_rn(s.w, __ldg(a.ts + start + base + lane), dep);
        }
        if (C > 0) {
            if constexpr (KC > 0) {
                const uint32_t cnt = min(32u, n - base);
                const float* rows = a.feats + (start + base) * a.fs;
                // rows are fetched in batches of U independent loads per lane before the first FMA (memory-level
                // parallelism: a ray is a strictly sequential sum); the summation order is unchanged
                constexpr int U = (KC <= 2) ? 8 : (KC <= 4) ? 4 : 2;
                for (uint32_t j0 = 0; j0 < cnt; j0 += U) {
                    float wj[U], v[U][KC];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        wj[u] = __shfl_sync(kFull, s.w, (j0 + u) & 31u);
                        if (j0 + u >= cnt) wj[u] = 0.0f;
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const float* row = rows + (size_t)(j0 + u) * a.fs;
#pragma unroll
                        for (int q = 0; q < KC; ++q) {
                            const uint32_t c = lane + 32u * q;
                            // terminated / transparent sample (w = 0): row never read
                            v[u][q] = (wj[u] != 0.0f && c < C) ? __ldg(row + c) : 0.0f;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (wj[u] == 0.0f) continue;
#pragma unroll
                        for (int q = 0; q < KC; ++q) acc[q] = __fmaf_rn(wj[u], v[u][q], acc[q]);
                    }
                }
            } else {
                if (s.valid && s.w != 0.0f) {
                    const float* row = a.feats + (start + base + lane) * a.fs;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if ((uint32_t)q < C) acc[q] = __fmaf_rn(s.w, __ldg(row + q), acc[q]);
                }
            }
        }
        if (alive_mask == 0u && a.t_thresh > 0.0f) {
            // every later sample has T < t_thresh as well: zero the tail and stop
            for (uint32_t i = base + 32 + lane; i < n; i += 32) weights[start + i] = 0.0f;
            break;
        }
    }
    ws = warp_sum(ws);
    dep = warp_sum(dep);
    if (lane == 0) {
        weights_sum[r] = ws;
        depth[r] = dep;
        if (n_alive) n_alive[r] = alive;
    }
    if (C > 0) {
        if constexpr (KC > 0) {
#pragma unroll
            for (int q = 0; q < KC; ++q) {
                const uint32_t c = lane + 32u * q;
                if (c < C) out[(size_t)r * C + c] = acc[q];
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float v = warp_sum(acc[q]);
                if (lane == 0 && (uint32_t)q < C) out[(size_t)r * C + q] = v;
            }
        }
    }
}

// 32 values per lane -> lane j holds sum over lanes of v[j]   (31 shuffles)
__device__ __forceinline__ float butterfly_transpose_sum(float (&v)[32], uint32_t lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
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

struct CompositeGrads {
    const float* g_weights;
    const float* g_weights_sum;
    const float* g_depth;
    const float* g_out;
};

template <int KC, int NCHUNK>
__global__ void __launch_bounds__(32 * kRaysPerBlock) composite_backward_kernel(
    const CompositeArgs a, const CompositeGrads g, float* __restrict__ grad_sigmas,
    float* __restrict__ grad_feats) {
    pdl_begin();
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t r = blockIdx.x * kRaysPerBlock + (threadIdx.x >> 5);
    if (r >= a.N) return;
    const size_t start = a.ray_offsets ? (size_t)a.ray_offsets[r] : (size_t)r * a.T;
    const uint32_t n = a.ray_offsets ? (uint32_t)(a.ray_offsets[r + 1] - a.ray_offsets[r]) : a.T;
    const uint32_t C = a.C;

    const float g_ws = g.g_weights_sum ? __ldg(g.g_weights_sum + r) : 0.0f;
    const float g_dp = g.g_depth ? __ldg(g.g_depth + r) : 0.0f;
    constexpr int NG = (KC > 0) ? KC : 8;
    float go[NG];
#pragma unroll
    for (int q = 0; q < NG; ++q) {
        const uint32_t c = (KC > 0) ? lane + 32u * q : (uint32_t)q;
        go[q] = (g.g_out && c < C) ? __ldg(g.g_out + (size_t)r * C + c) : 0.0f;
    }

    float gw_term[NCHUNK];   // g_i * w_i
    float gT_term[NCHUNK];   // g_i * T_{i+1}
    float carry = 0.0f;
#pragma unroll
    for (int k = 0; k < NCHUNK; ++k) {
        gw_term[k] = 0.0f;
        gT_term[k] = 0.0f;
        const uint32_t base = 32u * k;
        if (base >= n) continue;   // warp-uniform
        const SampleTerms s = chunk_terms(a, start, n, base, lane, carry);
        float dot = 0.0f;
        if (C > 0 && g.g_out) {
            if constexpr (KC > 0) {
                const uint32_t cnt = min(32u, n - base);
                const float* rows = a.feats + (start + base) * a.fs;
                float* grows = grad_feats ? grad_feats + (start + base) * a.gfs : nullptr;
                float part[32];
                // batches of U rows: all loads of a batch are issued before its FMAs and stores (the stores may alias
                // the loads as far as the compiler knows, so without batching every row is a full memory round trip)
                constexpr uint32_t U = (KC <= 2) ? 8 : (KC <= 4) ? 4 : 1;      // KC = 8 is register-bound already
#pragma unroll
                for (uint32_t j0 = 0; j0 < 32; j0 += U) {
                    float v[U][KC];
#pragma unroll
                    for (uint32_t u = 0; u < U; ++u) {
                        const float* row = rows + (size_t)(j0 + u) * a.fs;
#pragma unroll
                        for (int q = 0; q < KC; ++q) {
                            const uint32_t c = lane + 32u * q;
                            v[u][q] = (j0 + u < cnt && c < C) ? __ldg(row + c) : 0.0f;
                        }
                    }
#pragma unroll
                    for (uint32_t u = 0; u < U; ++u) {
                        const uint32_t j = j0 + u;
                        const float wj = __shfl_sync(kFull, s.w, j);
                        float p = 0.0f;
#pragma unroll
                        for (int q = 0; q < KC; ++q) {
                            const uint32_t c = lane + 32u * q;
                            p = __fmaf_rn(go[q], v[u][q], p);
                            if (grows && j < cnt && c < C) grows[(size_t)j * a.gfs + c] = wj * go[q];
                        }
                        part[j] = p;
                    }
                }
                dot = butterfly_transpose_sum(part, lane);
            } else {
                if (s.valid) {
                    const float* row = a.feats + (start + base + lane) * a.fs;
                    float* grow = grad_feats ? grad_feats + (start + base + lane) * a.gfs : nullptr;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        if ((uint32_t)q < C) {
                            dot = __fmaf_rn(go[q], __ldg(row + q), dot);
                            if (grow) grow[q] = s.w * go[q];
                        }
                    }
                }
            }
        } else if (C > 0 && grad_feats) {
            // no gradient reaches `out`: grad_feats is all zero
            const uint32_t cnt = min(32u, n - base);
            float* grows = grad_feats + (start + base) * a.gfs;
            for (uint32_t e = lane; e < cnt * C; e += 32) grows[(size_t)(e / C) * a.gfs + (e % C)] = 0.0f;
        }
        if (s.valid) {
            float gi = dot + g_ws;
            gi = __fmaf_rn(g_dp, __ldg(a.ts + start + base + lane), gi);
            if (g.g_weights) gi += __ldg(g.g_weights + start + base + lane);
            if (!s.alive || !s.finite) gi = 0.0f;   // weight was forced to 0 / clamped: no gradient path
            const float Tnext = s.T * expf(-s.x);
            gw_term[k] = gi * s.w;
            gT_term[k] = gi * Tnext;
        }
    }
    // reverse pass: suffix sums of g_j w_j
    float suffix_carry = 0.0f;
#pragma unroll
    for (int k = NCHUNK - 1; k >= 0; --k) {
        const uint32_t base = 32u * k;
        if (base >= n) continue;
        float incl = gw_term[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float v = __shfl_down_sync(kFull, incl, o);
            if (lane + (uint32_t)o < 32u) incl += v;
        }
        float excl = __shfl_down_sync(kFull, incl, 1);
        if (lane == 31) excl = 0.0f;
        const float sfx = excl + suffix_carry;
        suffix_carry += __shfl_sync(kFull, incl, 0);
        const uint32_t i = base + lane;
        if (i < n) {
            float ds = __ldg(a.deltas + start + i) * (gT_term[k] - sfx);
            if (a.last_opaque && i == n - 1u) ds = 0.0f;   // x := inf is a constant (renderer.py:315-316)
            grad_sigmas[start + i] = ds;
        }
    }
}

// Fallback for rays longer than 32*kMaxChunks samples: one thread per ray, sequential.
__global__ void composite_backward_long_kernel(const CompositeArgs a, const CompositeGrads g,
                                               float* __restrict__ grad_sigmas,
                                               float* __restrict__ grad_feats) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.N) return;
    const size_t start = a.ray_offsets ? (size_t)a.ray_offsets[r] : (size_t)r * a.T;
    const uint32_t n = a.ray_offsets ? (uint32_t)(a.ray_offsets[r + 1] - a.ray_offsets[r]) : a.T;
    const uint32_t C = a.C;
    const float g_ws = g.g_weights_sum ? g.g_weights_sum[r] : 0.0f;
    const float g_dp = g.g_depth ? g.g_depth[r] : 0.0f;
    // pass 1 (forward): grad_sigmas[i] <- g_i*T_{i+1} ; accumulate total of g_i*w_i
    float S = 0.0f, total = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        float x = a.deltas[start + i] * a.sigmas[start + i];
        if (a.last_opaque && i == n - 1u) x = INFINITY;
        const float T = expf(-S);
        float w = (1.0f - expf(-x)) * T;
        const bool alive = !(T < a.t_thresh);
        const bool fin = isfinite(w);
        if (isnan(w)) w = 0.0f; else if (isinf(w)) w = copysignf(FLT_MAX, w);
        if (!alive) w = 0.0f;
        float gi = g_ws + g_dp * a.ts[start + i] + (g.g_weights ? g.g_weights[start + i] : 0.0f);
        for (uint32_t c = 0; c < C; ++c) {
            const float go = g.g_out ? g.g_out[(size_t)r * C + c] : 0.0f;
            gi += go * a.feats[(start + i) * a.fs + c];
            if (grad_feats) grad_feats[(start + i) * a.gfs + c] = w * go;
        }
        if (!alive || !fin) gi = 0.0f;
        grad_sigmas[start + i] = gi * (T * expf(-x));
        total += gi * w;
        S += x;
    }
    // pass 2 (forward again): subtract the suffix = total - inclusive prefix
    S = 0.0f;
    float prefix = 0.0f;
    for (uint32_t i = 0; i < n; ++i) {
        float x = a.deltas[start + i] * a.sigmas[start + i];
        if (a.last_opaque && i == n - 1u) x = INFINITY;
        const float T = expf(-S);
        float w = (1.0f - expf(-x)) * T;
        const bool alive = !(T < a.t_thresh);
        const bool fin = isfinite(w);
        if (isnan(w)) w = 0.0f; else if (isinf(w)) w = copysignf(FLT_MAX, w);
        if (!alive) w = 0.0f;
        float gi = g_ws + g_dp * a.ts[start + i] + (g.g_weights ? g.g_weights[start + i] : 0.0f);
        for (uint32_t c = 0; c < C; ++c)
            gi += (g.g_out ? g.g_out[(size_t)r * C + c] : 0.0f) * a.feats[(start + i) * a.fs + c];
        if (!alive || !fin) gi = 0.0f;
        prefix += gi * w;
        float ds = a.deltas[start + i] * (grad_sigmas[start + i] - (total - prefix));
        if (a.last_opaque && i == n - 1u) ds = 0.0f;
        grad_sigmas[start + i] = ds;
        S += x;
    }
}

template <int KC>
static int launch_bwd_chunks(const CompositeArgs& a, const CompositeGrads& g, float* gs, float* gf,
                             cudaStream_t st) {
    const uint32_t blocks = div_up(a.N, (uint32_t)kRaysPerBlock);
    const uint32_t chunks = div_up(a.T, 32u);
    if (chunks <= 1) SANERF_LAUNCH((composite_backward_kernel<KC, 1>), blocks, 32 * kRaysPerBlock, 0, st, a, g, gs, gf);
    else if (chunks <= 2) SANERF_LAUNCH((composite_backward_kernel<KC, 2>), blocks, 32 * kRaysPerBlock, 0, st, a, g, gs, gf);
    else if (chunks <= 4) SANERF_LAUNCH((composite_backward_kernel<KC, 4>), blocks, 32 * kRaysPerBlock, 0, st, a, g, gs, gf);
    else SANERF_LAUNCH((composite_backward_kernel<KC, kMaxChunks>), blocks, 32 * kRaysPerBlock, 0, st, a, g, gs, gf);
    return check_launch("composite_backward_kernel");
}

}  // namespace sanerf

using namespace sanerf;

static int check_composite_args(const float* sigmas, const float* deltas, const float* ts, const float* feats,
                                uint32_t C) {
    SANERF_REQUIRE_PTR(sigmas);
    SANERF_REQUIRE_PTR(deltas);
    SANERF_REQUIRE_PTR(ts);
    if (C > 0) SANERF_REQUIRE_PTR(feats);
    if (C > 256) return fail(SANERF_ERR_INVALID_ARG, "composite: at most 256 channels per call");
    return SANERF_OK;
}

extern "C" int sanerf_composite_forward(const float* sigmas, const float* deltas, const float* ts,
                                        const float* feats, uint32_t feat_stride, const int32_t* ray_offsets, uint32_t N, uint32_t T,
                                        uint32_t C, int last_sample_opaque, float t_thresh, float* weights,
                                        float* weights_sum, float* depth, float* out, int32_t* n_alive,
                                        void* stream) {
    if (N == 0) return SANERF_OK;
    int rc = check_composite_args(sigmas, deltas, ts, feats, C);
    if (rc != SANERF_OK) return rc;
    SANERF_REQUIRE_PTR(weights);
    SANERF_REQUIRE_PTR(weights_sum);
    SANERF_REQUIRE_PTR(depth);
    if (C > 0) SANERF_REQUIRE_PTR(out);
    if (feat_stride != 0 && feat_stride < C) return fail(SANERF_ERR_INVALID_ARG, "composite: feat_stride < C");
    CompositeArgs a{sigmas, deltas, ts, feats, ray_offsets, N, T, C, last_sample_opaque, t_thresh,
                    feat_stride ? feat_stride : C, C};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t blocks = div_up(N, (uint32_t)kRaysPerBlock);
    const int threads = 32 * kRaysPerBlock;
    if (C <= 8) SANERF_LAUNCH((composite_forward_kernel<0>), blocks, threads, 0, st, a, weights, weights_sum, depth, out, n_alive);
    else if (C <= 32) SANERF_LAUNCH((composite_forward_kernel<1>), blocks, threads, 0, st, a, weights, weights_sum, depth, out, n_alive);
    else if (C <= 64) SANERF_LAUNCH((composite_forward_kernel<2>), blocks, threads, 0, st, a, weights, weights_sum, depth, out, n_alive);
    else if (C <= 128) SANERF_LAUNCH((composite_forward_kernel<4>), blocks, threads, 0, st, a, weights, weights_sum, depth, out, n_alive);
    else SANERF_LAUNCH((composite_forward_kernel<8>), blocks, threads, 0, st, a, weights, weights_sum, depth, out, n_alive);
    return check_launch("composite_forward_kernel");
}

extern "C" int sanerf_composite_backward(const float* sigmas, const float* deltas, const float* ts,
                                         const float* feats, uint32_t feat_stride, const int32_t* ray_offsets, uint32_t N, uint32_t T,
                                         uint32_t C, int last_sample_opaque, float t_thresh, const float* weights,
                                         const float* g_weights, const float* g_weights_sum, const float* g_depth,
                                         const float* g_out, float* grad_sigmas, float* grad_feats,
                                         uint32_t grad_feat_stride, void* stream) {
    (void)weights;  // recomputed from sigmas/deltas in registers (cheaper than re-reading)
    if (N == 0) return SANERF_OK;
    int rc = check_composite_args(sigmas, deltas, ts, feats, C);
    if (rc != SANERF_OK) return rc;
    SANERF_REQUIRE_PTR(grad_sigmas);
    if ((feat_stride != 0 && feat_stride < C) || (grad_feat_stride != 0 && grad_feat_stride < C))
        return fail(SANERF_ERR_INVALID_ARG, "composite: feature stride < C");
    CompositeArgs a{sigmas, deltas, ts, feats, ray_offsets, N, T, C, last_sample_opaque, t_thresh,
                    feat_stride ? feat_stride : C, grad_feat_stride ? grad_feat_stride : C};
    CompositeGrads g{g_weights, g_weights_sum, g_depth, g_out};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (T > 32u * kMaxChunks) {
        composite_backward_long_kernel<<<div_up(N, 128u), 128, 0, st>>>(a, g, grad_sigmas, grad_feats);
        return check_launch("composite_backward_long_kernel");
    }
    if (C <= 8) return launch_bwd_chunks<0>(a, g, grad_sigmas, grad_feats, st);
    if (C <= 32) return launch_bwd_chunks<1>(a, g, grad_sigmas, grad_feats, st);
    if (C <= 64) return launch_bwd_chunks<2>(a, g, grad_sigmas, grad_feats, st);
    if (C <= 128) return launch_bwd_chunks<4>(a, g, grad_sigmas, grad_feats, st);
    return launch_bwd_chunks<8>(a, g, grad_sigmas, grad_feats, st);
}
