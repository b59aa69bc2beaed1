// Sampling regularisers of the training step, fused (sm_100a): inter-level proposal loss and
// distortion loss, each ONE warp-per-ray kernel that returns the loss AND d loss / d weights.
//
// Replaces the cumsum / searchsorted / take_along_dim / clamp chains of nerf/renderer.py:30-57 (proposal_loss)
// and renderer.py:17-27 + torch_efficient_distloss (distort_loss), forward and backward (~40 torch kernels).
//
//  proposal:  for every reference interval k of a ray:  bound_k = sum of proposal weights over the proposal
//             intervals overlapping it (outer measure);  L = mean_k max(w_ref_k - bound_k, 0)^2 / (w_ref_k + 1e-8)
//             dL/dw_p[i] = - sum_{k : lo_k <= i <= hi_k} 2 max(w_ref_k - bound_k, 0) / (w_ref_k + 1e-8) / (N T_ref)
//  distortion (Mip-NeRF 360 eq. 15, O(T) form of Sun et al. 2022):
//             L = 1/N sum_rays [ 1/3 sum_i d_i w_i^2 + 2 sum_i w_i (m_i W_<i - WM_<i) ]
//             dL/dw_i = 1/N [ 2/3 d_i w_i + 2 ( m_i (W_<i - W_>i) + WM_>i - WM_<i ) ]
#include "common.cuh"

namespace sanerf {

constexpr int kLossWarps = 4;
constexpr uint32_t kFullMask = 0xffffffffu;

// inclusive scan of per-lane values laid out as element i = base + lane, carried across chunks
__device__ __forceinline__ float warp_incl_scan(float v, uint32_t lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(kFullMask, v, o);
        if (lane >= (uint32_t)o) v += t;
    }
    return v;
}

// searchsorted(a[0..n), x, right=True): first index with a[idx] > x
__device__ __forceinline__ uint32_t upper_bound(const float* a, uint32_t n, float x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] > x) hi = mid; else lo = mid + 1u;
    }
    return lo;
}

// shared per warp: t_p[Tp+1] | cum[Tp+1] (exclusive prefix, cum[0]=0 .. cum[Tp]=total) | diff[Tp+1]
__global__ void __launch_bounds__(32 * kLossWarps) proposal_loss_kernel(
    const float* __restrict__ t_ref, const float* __restrict__ w_ref, uint32_t Tr, const float* __restrict__ t_p,
    const float* __restrict__ w_p, uint32_t Tp, uint32_t N, float weight, float* __restrict__ loss_out,
    float* __restrict__ g_wp) {
    pdl_begin();
    extern __shared__ float smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * kLossWarps + warp;
    if (r >= N) return;
    float* tp = smem + (size_t)warp * 3u * (Tp + 1u);
    float* cum = tp + (Tp + 1u);
    float* diff = cum + (Tp + 1u);
    const float scale = weight / ((float)N * (float)Tr);   // .mean() over [N, Tr], times lambda_proposal

    for (uint32_t i = lane; i <= Tp; i += 32) { tp[i] = __ldg(t_p + (size_t)r * (Tp + 1u) + i); diff[i] = 0.0f; }
    float carry = 0.0f;
    for (uint32_t base = 0; base < Tp; base += 32) {
        const uint32_t i = base + lane;
        const float v = (i < Tp) ? __ldg(w_p + (size_t)r * Tp + i) : 0.0f;
        const float inc = warp_incl_scan(v, lane);
        if (i < Tp) cum[i + 1u] = carry + inc;
        carry += __shfl_sync(kFullMask, inc, 31);
    }
    if (lane == 0) cum[0] = 0.0f;
    __syncwarp();

    float loss = 0.0f;
    for (uint32_t k = lane; k < Tr; k += 32) {
        const float a = __ldg(t_ref + (size_t)r * (Tr + 1u) + k), b = __ldg(t_ref + (size_t)r * (Tr + 1u) + k + 1u);
        const float wr = __ldg(w_ref + (size_t)r * Tr + k);
        // inds_lo = clamp(searchsorted(t_p[:-1], a, right) - 1, 0, Tp-1); inds_hi = clamp(searchsorted(t_p[1:], b, right), 0, Tp-1)
        const uint32_t ub_lo = upper_bound(tp, Tp, a);
        const uint32_t lo = (ub_lo == 0u) ? 0u : min(ub_lo - 1u, Tp - 1u);
        const uint32_t hi = min(upper_bound(tp + 1, Tp, b), Tp - 1u);
        const float bound = cum[hi + 1u] - cum[lo];          // cw1[1:][hi] - cw1[:-1][lo]
        const float excess = fmaxf(wr - bound, 0.0f);
        const float denom = wr + 1e-8f;
        loss += excess * excess / denom;
        if (excess > 0.0f && g_wp != nullptr) {
            const float c = -2.0f * excess / denom * scale;
            // d bound / d w_p[i] = 1 for lo <= i <= hi when hi >= lo; if hi < lo the difference of prefix sums
            // is minus the sum over (hi, lo): keep the exact derivative of the expression
            if (hi >= lo) { atomicAdd(diff + lo, c); atomicAdd(diff + hi + 1u, -c); }
            else { atomicAdd(diff + hi + 1u, -c); atomicAdd(diff + lo, c); }
        }
    }
    loss = warp_sum(loss);
    if (lane == 0) atomicAdd(loss_out, loss * scale);
    if (g_wp == nullptr) return;
    __syncwarp();
    carry = 0.0f;
    for (uint32_t base = 0; base < Tp; base += 32) {
        const uint32_t i = base + lane;
        const float v = (i < Tp) ? diff[i] : 0.0f;
        const float inc = warp_incl_scan(v, lane);
        if (i < Tp) g_wp[(size_t)r * Tp + i] = carry + inc;
        carry += __shfl_sync(kFullMask, inc, 31);
    }
}

__global__ void __launch_bounds__(32 * kLossWarps) distortion_loss_kernel(
    const float* __restrict__ bins, const float* __restrict__ w, uint32_t T, uint32_t N, float weight,
    float* __restrict__ loss_out, float* __restrict__ g_w) {
    pdl_begin();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t r = blockIdx.x * kLossWarps + warp;
    if (r >= N) return;
    const float inv_n = weight / (float)N;
    const float* b = bins + (size_t)r * (T + 1u);
    const float* wr = w + (size_t)r * T;
    // totals first (needed for the suffix terms of the gradient)
    float w_tot = 0.0f, wm_tot = 0.0f;
    for (uint32_t i = lane; i < T; i += 32) {
        const float d = __ldg(b + i + 1u) - __ldg(b + i);
        const float m = __ldg(b + i) + d * 0.5f;
        const float wi = __ldg(wr + i);
        w_tot += wi;
        wm_tot = __fmaf_rn(wi, m, wm_tot);
    }
    w_tot = warp_sum(w_tot);
    wm_tot = warp_sum(wm_tot);
    float loss = 0.0f, cw = 0.0f, cwm = 0.0f;
    for (uint32_t base = 0; base < T; base += 32) {
        const uint32_t i = base + lane;
        float d = 0.0f, m = 0.0f, wi = 0.0f;
        if (i < T) {
            d = __ldg(b + i + 1u) - __ldg(b + i);
            m = __ldg(b + i) + d * 0.5f;
            wi = __ldg(wr + i);
        }
        const float iw = warp_incl_scan(wi, lane), iwm = warp_incl_scan(wi * m, lane);
        const float w_pre = cw + iw - wi, wm_pre = cwm + iwm - wi * m;     // exclusive prefixes
        if (i < T) {
            loss += d * wi * wi * (1.0f / 3.0f) + 2.0f * wi * (m * w_pre - wm_pre);
            if (g_w != nullptr) {
                const float w_suf = w_tot - (w_pre + wi), wm_suf = wm_tot - (wm_pre + wi * m);
                g_w[(size_t)r * T + i] = inv_n * ((2.0f / 3.0f) * d * wi + 2.0f * (m * (w_pre - w_suf) + (wm_suf - wm_pre)));
            }
        }
        cw += __shfl_sync(kFullMask, iw, 31);
        cwm += __shfl_sync(kFullMask, iwm, 31);
    }
    loss = warp_sum(loss);
    if (lane == 0) atomicAdd(loss_out, loss * inv_n);
}

// ---- fused Adam over a flat parameter buffer -------------------------------------------------------
// torch.optim.Adam semantics (no weight decay, no amsgrad):  m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ;
// p -= lr / (1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).  `dyn` = {lr_t, 1-b1^t, 1-b2^t} lives on the device so
// that a captured CUDA graph can be replayed with a moving step count / learning-rate schedule.
__global__ void adam_schedule_kernel(int32_t* step, float* dyn, float lr0, float beta1, float beta2, float decay_iters,
                                     int32_t* gate, float ema_decay) {
    pdl_begin();
    const int32_t t = *step + 1;
    *step = t;
    // torch_ema.ExponentialMovingAverage.update (nerf/utils.py:616, 1862): decay_t = min(decay, (1 + t) / (10 + t))
    dyn[3] = 1.0f - fminf(ema_decay, (1.0f + (float)t) / (10.0f + (float)t));
    if (gate) *gate = 1;      // the step that starts here leaves a gradient behind for every deferred range
    // LambdaLR(0.1 ** min(iter / iters, 1)) evaluated at iter = t-1 (main.py:312-313; scheduler steps after the optimizer)
    const float frac = (decay_iters > 0.0f) ? fminf((float)(t - 1) / decay_iters, 1.0f) : 0.0f;
    dyn[0] = lr0 * powf(0.1f, frac);
    dyn[1] = 1.0f - powf(beta1, (float)t);
    dyn[2] = 1.0f - powf(beta2, (float)t);
}

template <int kThreadsMax>
__global__ void __launch_bounds__(kThreadsMax) adam_step_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, size_t n4, size_t n,
                                                        const float* __restrict__ dyn, float beta1, float beta2,
                                                        float eps, float grad_scale, int zero_grad,
                                                        const int32_t* __restrict__ gate, float* __restrict__ ema) {
    pdl_begin();
    // A deferred range is updated at the START of the next step; after a flush() (which applied the update early and
    // cleared the gate) there is nothing pending and a pass over zero gradients must not move the parameters by momentum.
    if (gate != nullptr && *gate == 0) return;
    const float lr = __ldg(dyn), bc1 = __ldg(dyn + 1), bc2 = __ldg(dyn + 2), ema_w = __ldg(dyn + 3);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    auto update = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= grad_scale;
        mm = beta1 * mm + (1.0f - beta1) * gg;
        vv = beta2 * vv + (1.0f - beta2) * gg * gg;
        pp -= step_size * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
    };
    // shadow -= (1 - decay_t) * (shadow - param), with the parameter just updated (torch_ema update after optimizer.step)
    auto ema4 = [&](size_t k, const float4& q) {
        float4 e = __ldcs(reinterpret_cast<float4*>(ema) + k);
        e.x -= ema_w * (e.x - q.x); e.y -= ema_w * (e.y - q.y); e.z -= ema_w * (e.z - q.z); e.w -= ema_w * (e.w - q.w);
        __stcs(reinterpret_cast<float4*>(ema) + k, e);
    };
    // Every element is read and written exactly once per step: streaming (evict-first) accesses keep the pass from
    // flushing the tables and activations of the kernels it overlaps with out of the L2.
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n4; i += 2 * stride) {                // two elements per trip: 8 independent 16-byte loads in flight
        const size_t j = i + stride;
        float4 pp = __ldcs(reinterpret_cast<float4*>(p) + i), gg = __ldcs(reinterpret_cast<float4*>(g) + i);
        float4 mm = __ldcs(reinterpret_cast<float4*>(m) + i), vv = __ldcs(reinterpret_cast<float4*>(v) + i);
        float4 pq = __ldcs(reinterpret_cast<float4*>(p) + j), gq = __ldcs(reinterpret_cast<float4*>(g) + j);
        float4 mq = __ldcs(reinterpret_cast<float4*>(m) + j), vq = __ldcs(reinterpret_cast<float4*>(v) + j);
        update(pp.x, gg.x, mm.x, vv.x); update(pp.y, gg.y, mm.y, vv.y);
        update(pp.z, gg.z, mm.z, vv.z); update(pp.w, gg.w, mm.w, vv.w);
        update(pq.x, gq.x, mq.x, vq.x); update(pq.y, gq.y, mq.y, vq.y);
        update(pq.z, gq.z, mq.z, vq.z); update(pq.w, gq.w, mq.w, vq.w);
        __stcs(reinterpret_cast<float4*>(p) + i, pp); __stcs(reinterpret_cast<float4*>(m) + i, mm); __stcs(reinterpret_cast<float4*>(v) + i, vv);
        __stcs(reinterpret_cast<float4*>(p) + j, pq); __stcs(reinterpret_cast<float4*>(m) + j, mq); __stcs(reinterpret_cast<float4*>(v) + j, vq);
        if (ema) { ema4(i, pp); ema4(j, pq); }
        if (zero_grad) {
            __stcs(reinterpret_cast<float4*>(g) + i, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
            __stcs(reinterpret_cast<float4*>(g) + j, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
        }
    }
    for (; i < n4; i += stride) {
        float4 pp = __ldcs(reinterpret_cast<float4*>(p) + i), gg = __ldcs(reinterpret_cast<float4*>(g) + i);
        float4 mm = __ldcs(reinterpret_cast<float4*>(m) + i), vv = __ldcs(reinterpret_cast<float4*>(v) + i);
        update(pp.x, gg.x, mm.x, vv.x); update(pp.y, gg.y, mm.y, vv.y);
        update(pp.z, gg.z, mm.z, vv.z); update(pp.w, gg.w, mm.w, vv.w);
        __stcs(reinterpret_cast<float4*>(p) + i, pp); __stcs(reinterpret_cast<float4*>(m) + i, mm); __stcs(reinterpret_cast<float4*>(v) + i, vv);
        if (ema) ema4(i, pp);
        if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        update(p[i], g[i], m[i], v[i]);
        if (ema) ema[i] -= ema_w * (ema[i] - p[i]);
        if (zero_grad) g[i] = 0.0f;
    }
}

// Adam for a HALF-precision table with an fp32 master copy (BASELINE configs[4]: T = 2^22 fp16 parameters): the gradient
// arrives in half (the scatter's red.global.add.f16x2 target, or the reduce-scattered sum of them), the moments and the
// master parameters stay fp32, and the half table the gather kernels read is rewritten in the same pass.  8 elements per
// thread and trip: one 16-byte access per half array, two per fp32 array.
__global__ void __launch_bounds__(256) adam_step_half_kernel(float* __restrict__ p, __half* __restrict__ p16, __half* __restrict__ g16,
                                                             float* __restrict__ m, float* __restrict__ v, size_t n8,
                                                             const float* __restrict__ dyn, float beta1, float beta2, float eps,
                                                             float grad_scale, int zero_grad) {
    pdl_begin();
    const float lr = __ldg(dyn), bc1 = __ldg(dyn + 1), bc2 = __ldg(dyn + 2);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        uint4 graw = __ldcs(reinterpret_cast<const uint4*>(g16) + i);
        const __half2* gh = reinterpret_cast<const __half2*>(&graw);
        float4 pp[2], mm[2], vv[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            pp[h] = __ldcs(reinterpret_cast<float4*>(p) + 2 * i + h);
            mm[h] = __ldcs(reinterpret_cast<float4*>(m) + 2 * i + h);
            vv[h] = __ldcs(reinterpret_cast<float4*>(v) + 2 * i + h);
        }
        float* pf = reinterpret_cast<float*>(pp);
        float* mf = reinterpret_cast<float*>(mm);
        float* vf = reinterpret_cast<float*>(vv);
        uint4 praw;
        __half2* ph = reinterpret_cast<__half2*>(&praw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 g2 = __half22float2(gh[j]);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 2 * j + e;
                const float gg = (e ? g2.y : g2.x) * grad_scale;
                mf[k] = beta1 * mf[k] + (1.0f - beta1) * gg;
                vf[k] = beta2 * vf[k] + (1.0f - beta2) * gg * gg;
                pf[k] -= step_size * mf[k] / (sqrtf(vf[k]) * inv_sqrt_bc2 + eps);
            }
            ph[j] = __floats2half2_rn(pf[2 * j], pf[2 * j + 1]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            __stcs(reinterpret_cast<float4*>(p) + 2 * i + h, pp[h]);
            __stcs(reinterpret_cast<float4*>(m) + 2 * i + h, mm[h]);
            __stcs(reinterpret_cast<float4*>(v) + 2 * i + h, vv[h]);
        }
        reinterpret_cast<uint4*>(p16)[i] = praw;               // the table the next forward gathers from: keep it cacheable
        if (zero_grad) __stcs(reinterpret_cast<uint4*>(g16) + i, make_uint4(0u, 0u, 0u, 0u));
    }
}

// ---- uniform jitter for the samplers (renderer.py:269 and :101: torch.rand_like, [N, T+1] per level) -----------------
// Philox4x32-10 (Salmon et al. 2011), counter = (element index / 4, call number), key = seed: 4 uniforms in [0,1) per
// counter.  The call number lives on the device and advances by one per launch, so a captured CUDA graph draws fresh
// numbers at every replay (the stream torch's generator would have produced is not reproduced: the reference has no seed
// contract for the jitter, and parity tests feed the very buffer this kernel fills to the oracle).
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__global__ void __launch_bounds__(256) uniform_fill_kernel(float* __restrict__ out, size_t n, uint64_t seed,
                                                           uint32_t* __restrict__ call_counter, float* __restrict__ zero_a,
                                                           uint32_t zero_n) {
    pdl_begin();
    const uint32_t call = *call_counter;
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 < zero_n) zero_a[i4] = 0.0f;                      // optional: clear a small accumulator (the step's loss) as well
    if (i4 * 4 < n) {
        uint32_t c[4] = {(uint32_t)i4, (uint32_t)(i4 >> 32), call, 0x5a4e5246u};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            philox_round(c, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        float u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = (float)(c[j] >> 8) * (1.0f / 16777216.0f);    // 24 bits: [0, 1)
        if (i4 * 4 + 3 < n) {
            reinterpret_cast<float4*>(out)[i4] = make_float4(u[0], u[1], u[2], u[3]);
        } else {
            for (size_t j = 0; i4 * 4 + j < n; ++j) out[i4 * 4 + j] = u[j];
        }
    }
    __syncthreads();
    // the last block to finish advances the call counter: every thread of every block has read it by then
    __shared__ bool last;
    if (threadIdx.x == 0) {
        __threadfence();
        last = (atomicAdd(call_counter + 1, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        call_counter[1] = 0u;
        call_counter[0] = call + 1u;
    }
}

// torch_ema.ExponentialMovingAverage.update(): shadow -= (1 - decay_t) * (shadow - param).  The reference calls it once per
// EPOCH (nerf/utils.py:1862, after the loop of train_one_epoch) / once per 16 GUI steps (:1627), not per optimizer step.
__global__ void __launch_bounds__(256) ema_update_kernel(float* __restrict__ shadow, const float* __restrict__ param, size_t n4,
                                                         size_t n, float one_minus_decay) {
    pdl_begin();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 e = __ldcs(reinterpret_cast<const float4*>(shadow) + i);
        const float4 q = __ldcs(reinterpret_cast<const float4*>(param) + i);
        e.x -= one_minus_decay * (e.x - q.x); e.y -= one_minus_decay * (e.y - q.y);
        e.z -= one_minus_decay * (e.z - q.z); e.w -= one_minus_decay * (e.w - q.w);
        __stcs(reinterpret_cast<float4*>(shadow) + i, e);
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        shadow[i] -= one_minus_decay * (shadow[i] - param[i]);
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_proposal_loss(const float* t_ref, const float* w_ref, uint32_t Tr, const float* t_p,
                                    const float* w_p, uint32_t Tp, uint32_t N, float weight, float* loss_out,
                                    float* g_wp, void* stream) {
    if (N == 0 || Tr == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(t_ref); SANERF_REQUIRE_PTR(w_ref); SANERF_REQUIRE_PTR(t_p); SANERF_REQUIRE_PTR(w_p);
    SANERF_REQUIRE_PTR(loss_out);
    if (Tp == 0) return fail(SANERF_ERR_INVALID_ARG, "proposal_loss: proposal level has no samples");
    const size_t smem = (size_t)kLossWarps * 3u * (Tp + 1u) * sizeof(float);
    if (smem > 48 * 1024) return fail(SANERF_ERR_INVALID_ARG, "proposal_loss: Tp too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SANERF_LAUNCH(proposal_loss_kernel, div_up(N, (uint32_t)kLossWarps), 32 * kLossWarps, smem, st, t_ref, w_ref, Tr, t_p, w_p, Tp,
                                                                                         N, weight, loss_out, g_wp);
    return check_launch("proposal_loss_kernel");
}

extern "C" int sanerf_distortion_loss(const float* bins, const float* w, uint32_t T, uint32_t N, float weight,
                                      float* loss_out, float* g_w, void* stream) {
    if (N == 0 || T == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(bins); SANERF_REQUIRE_PTR(w); SANERF_REQUIRE_PTR(loss_out);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SANERF_LAUNCH(distortion_loss_kernel, div_up(N, (uint32_t)kLossWarps), 32 * kLossWarps, 0, st, bins, w, T, N, weight, loss_out, g_w);
    return check_launch("distortion_loss_kernel");
}

extern "C" int sanerf_adam_schedule(int32_t* step, float* dyn, float lr0, float beta1, float beta2, float decay_iters,
                                    int32_t* gate, float ema_decay, void* stream) {
    SANERF_REQUIRE_PTR(step); SANERF_REQUIRE_PTR(dyn);
    SANERF_LAUNCH(adam_schedule_kernel, 1, 1, 0, static_cast<cudaStream_t>(stream), step, dyn, lr0, beta1, beta2, decay_iters,
                  gate, ema_decay);
    return check_launch("adam_schedule_kernel");
}

extern "C" int sanerf_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, uint64_t n,
                                const float* dyn, float beta1, float beta2, float eps, float grad_scale,
                                int zero_grad, const int32_t* gate, float* ema, void* stream) {
    if (n == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(params); SANERF_REQUIRE_PTR(grads); SANERF_REQUIRE_PTR(exp_avg); SANERF_REQUIRE_PTR(exp_avg_sq);
    SANERF_REQUIRE_PTR(dyn);
    const uintptr_t align = (uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)ema;
    if (align & 15u) return fail(SANERF_ERR_MISALIGNED, "adam_step: buffers must be 16-byte aligned");
    const size_t n4 = (size_t)n / 4;
    size_t blocks = div_up(n4 > 0 ? n4 : (size_t)1, (size_t)256);
    // Short-lived CTAs (4 grid-stride trips each) instead of a persistent grid that owns every thread slot of the machine
    // for the whole pass: the deferred table update runs on a default-priority stream beside the step's high-priority
    // critical chain, whose CTAs can only be placed when CTAs of this kernel retire.  (Measured alternative: two
    // persistent CTAs per SM - no better; the interference is in the memory system, see the streaming accesses above.)
    blocks = div_up(blocks, (size_t)4);
    // Large passes that run BESIDE other kernels (the deferred table update beside the next step's front) can instead be
    // launched as a small persistent grid of 1024-thread CTAs (SANERF_ADAM_PERSISTENT=<CTAs>): each owns one SM's register
    // file for the whole pass, so the pass keeps a fixed slice of the machine instead of waiting for slots the front's
    // register-hungry CTAs leave free ("register-file exclusivity").
    static const int persistent = [] { const char* e = getenv("SANERF_ADAM_PERSISTENT"); return e ? atoi(e) : 0; }();
    if (persistent > 0 && n >= (1u << 22)) {
        SANERF_LAUNCH(adam_step_kernel<1024>, (uint32_t)persistent, 1024, 0, static_cast<cudaStream_t>(stream),
            params, grads, exp_avg, exp_avg_sq, n4, (size_t)n, dyn, beta1, beta2, eps, grad_scale, zero_grad, gate, ema);
        return check_launch("adam_step_kernel");
    }
    SANERF_LAUNCH(adam_step_kernel<256>, (uint32_t)blocks, 256, 0, static_cast<cudaStream_t>(stream), 
        params, grads, exp_avg, exp_avg_sq, n4, (size_t)n, dyn, beta1, beta2, eps, grad_scale, zero_grad, gate, ema);
    return check_launch("adam_step_kernel");
}

extern "C" int sanerf_adam_step_half(float* master, void* params16, void* grads16, float* exp_avg, float* exp_avg_sq,
                                     uint64_t n, const float* dyn, float beta1, float beta2, float eps, float grad_scale,
                                     int zero_grad, void* stream) {
    if (n == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(master); SANERF_REQUIRE_PTR(params16); SANERF_REQUIRE_PTR(grads16);
    SANERF_REQUIRE_PTR(exp_avg); SANERF_REQUIRE_PTR(exp_avg_sq); SANERF_REQUIRE_PTR(dyn);
    const uintptr_t align = (uintptr_t)master | (uintptr_t)params16 | (uintptr_t)grads16 | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq;
    if ((align & 15u) || (n & 7u)) return fail(SANERF_ERR_MISALIGNED, "adam_step_half: 16-byte aligned buffers and n % 8 == 0");
    const size_t n8 = (size_t)n / 8;
    const size_t blocks = div_up(div_up(n8, (size_t)256), (size_t)4);
    SANERF_LAUNCH(adam_step_half_kernel, (uint32_t)blocks, 256, 0, static_cast<cudaStream_t>(stream), master,
                  static_cast<__half*>(params16), static_cast<__half*>(grads16), exp_avg, exp_avg_sq, n8, dyn, beta1, beta2,
                  eps, grad_scale, zero_grad);
    return check_launch("adam_step_half_kernel");
}

extern "C" int sanerf_uniform_fill(float* out, uint64_t n, uint64_t seed, uint32_t* state, float* zero, uint32_t zero_n,
                                   void* stream) {
    if (n == 0 && zero_n == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(out); SANERF_REQUIRE_PTR(state);
    if ((uintptr_t)out & 15u) return fail(SANERF_ERR_MISALIGNED, "uniform_fill: out must be 16-byte aligned");
    const size_t n4 = div_up((size_t)n, (size_t)4);
    const size_t work = n4 > zero_n ? n4 : (size_t)zero_n;
    SANERF_LAUNCH(uniform_fill_kernel, (uint32_t)div_up(work, (size_t)256), 256, 0, static_cast<cudaStream_t>(stream), out,
                  (size_t)n, seed, state, zero, zero_n);
    return check_launch("uniform_fill_kernel");
}

extern "C" int sanerf_ema_update(float* shadow, const float* params, uint64_t n, float one_minus_decay, void* stream) {
    if (n == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(shadow); SANERF_REQUIRE_PTR(params);
    if (((uintptr_t)shadow | (uintptr_t)params) & 15u) return fail(SANERF_ERR_MISALIGNED, "ema_update: buffers must be 16-byte aligned");
    const size_t n4 = (size_t)n / 4;
    const size_t blocks = div_up(div_up(n4 > 0 ? n4 : (size_t)1, (size_t)256), (size_t)4);
    SANERF_LAUNCH(ema_update_kernel, (uint32_t)blocks, 256, 0, static_cast<cudaStream_t>(stream), shadow, params, n4, (size_t)n,
                  one_minus_decay);
    return check_launch("ema_update_kernel");
}

extern "C" int sanerf_copy_rows(float* dst, uint32_t ld_dst, const float* src, uint32_t ld_src, uint32_t rows, uint32_t cols,
                                void* stream) {
    if (rows == 0 || cols == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(dst); SANERF_REQUIRE_PTR(src);
    if (ld_dst < cols || ld_src < cols) return fail(SANERF_ERR_INVALID_ARG, "copy_rows: leading dimensions must cover the columns");
    cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)ld_dst * 4u, src, (size_t)ld_src * 4u, (size_t)cols * 4u, rows,
                                      cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return fail(SANERF_ERR_CUDA, "copy_rows: %s", cudaGetErrorString(e));
    return SANERF_OK;
}
