// Deferred-shading view head of the training / rendering step, ONE kernel (sm_100a):
//   f = [ sum_i w_i geo_i (15) , weights_sum * SH4(d) (16) ]            renderer.py:333-338, network.py:237
//   rgb = sigmoid(W3 relu(W2 relu(W1 f)))                               network.py:107 (view_mlp 31->32->32->3, no bias)
//   image = rgb + (1 - weights_sum) * bg                                renderer.py:345, 358
// and, when a target is given, the photometric loss of Trainer.train_step (nerf/utils.py:897-930)
//   loss += weight * mean((image - gt)^2)
// together with its complete backward: d loss / d geo_sum, d loss / d weights_sum and the three weight gradients.
// Replaces ~50 torch launches per step (SH kernel, cat, 3 cuBLAS SGEMMs + ReLU/sigmoid, MSE, and their autograd
// mirrors).  One thread = one ray; the weights (8.4 KB) are broadcast from shared memory; weight gradients are
// reduced over the block's rays in shared memory and added to HBM once per block.
#include "common.cuh"
#include "sh_common.cuh"

namespace sanerf {

namespace vh {
constexpr int kGeo = 15, kSh = 16, kIn = kGeo + kSh, kHid = 32, kOut = 3;
constexpr int kRays = 64;                 // rays (threads) per block
constexpr int kPitch = kRays + 1;         // stash pitch: conflict-free reads along features and along rays
constexpr int kW1 = kHid * kIn, kW2 = kHid * kHid, kW3 = kOut * kHid;
constexpr int kStashRows = kIn + kHid + kHid + kHid + kHid + kOut;     // f, h1, h2, g1, g2, g3
constexpr size_t kSmemFwd = (size_t)(kW1 + kW2 + kW3) * sizeof(float);
constexpr size_t kSmemBwd = kSmemFwd + (size_t)kStashRows * kPitch * sizeof(float);
}  // namespace vh

struct ViewHeadParams {
    const float* geo_sum;      // [N,15]
    const float* weights_sum;  // [N]
    const float* rays_d;       // [N,3] (normalised here, sphere_harmonics.py:79-82)
    const float* gt;           // [N,3] or NULL (forward only)
    const float* w1;           // [32,31]
    const float* w2;           // [32,32]
    const float* w3;           // [3,32]
    float bg;
    float loss_weight;
    uint32_t N;
    float* image;              // [N,3]
    float* loss;               // scalar, accumulated
    float* g_geo_sum;          // [N,15]
    float* g_weights_sum;      // [N]
    float* g_w1;               // accumulated
    float* g_w2;
    float* g_w3;
};

template <bool TRAIN>
__global__ void __launch_bounds__(vh::kRays) view_head_kernel(const ViewHeadParams p) {
    pdl_begin();
    using namespace vh;
    extern __shared__ float smem[];
    float* sW1 = smem;
    float* sW2 = sW1 + kW1;
    float* sW3 = sW2 + kW2;
    float* stash = sW3 + kW3;
    const int t = threadIdx.x;
    for (int i = t; i < kW1; i += kRays) sW1[i] = __ldg(p.w1 + i);
    for (int i = t; i < kW2; i += kRays) sW2[i] = __ldg(p.w2 + i);
    for (int i = t; i < kW3; i += kRays) sW3[i] = __ldg(p.w3 + i);
    __syncthreads();

    const uint32_t r = blockIdx.x * kRays + t;
    const bool live = r < p.N;
    float f[kIn], sh[kSh], h1[kHid], h2[kHid];
    float ws = 0.0f;
    if (live) {
        ws = __ldg(p.weights_sum + r);
#pragma unroll
        for (int i = 0; i < kGeo; ++i) f[i] = __ldg(p.geo_sum + (size_t)r * kGeo + i);
        float x = __ldg(p.rays_d + (size_t)r * 3), y = __ldg(p.rays_d + (size_t)r * 3 + 1), z = __ldg(p.rays_d + (size_t)r * 3 + 2);
        const float n = sqrtf(x * x + y * y + z * z);
        x /= n; y /= n; z /= n;
        sh_eval<4, false>(x, y, z, sh, nullptr);
    } else {
#pragma unroll
        for (int i = 0; i < kGeo; ++i) f[i] = 0.0f;
#pragma unroll
        for (int i = 0; i < kSh; ++i) sh[i] = 0.0f;
    }
#pragma unroll
    for (int i = 0; i < kSh; ++i) f[kGeo + i] = ws * sh[i];
#pragma unroll
    for (int j = 0; j < kHid; ++j) {
        float a = 0.0f;
#pragma unroll
        for (int i = 0; i < kIn; ++i) a = __fmaf_rn(sW1[j * kIn + i], f[i], a);
        h1[j] = fmaxf(a, 0.0f);
    }
#pragma unroll
    for (int k = 0; k < kHid; ++k) {
        float a = 0.0f;
#pragma unroll
        for (int j = 0; j < kHid; ++j) a = __fmaf_rn(sW2[k * kHid + j], h1[j], a);
        h2[k] = fmaxf(a, 0.0f);
    }
    float rgb[kOut], img[kOut];
#pragma unroll
    for (int c = 0; c < kOut; ++c) {
        float a = 0.0f;
#pragma unroll
        for (int k = 0; k < kHid; ++k) a = __fmaf_rn(sW3[c * kHid + k], h2[k], a);
        rgb[c] = 1.0f / (1.0f + expf(-a));
        img[c] = rgb[c] + (1.0f - ws) * p.bg;
        if (live) p.image[(size_t)r * kOut + c] = img[c];
    }
    if constexpr (!TRAIN) return;

    // ---- loss and backward of this ray
    float g3[kOut], g2[kHid], g1[kHid];
    float loss = 0.0f, g_ws = 0.0f;
    const float inv = p.loss_weight / (3.0f * (float)p.N);       // mean over [N,3]
#pragma unroll
    for (int c = 0; c < kOut; ++c) {
        const float diff = live ? img[c] - __ldg(p.gt + (size_t)r * kOut + c) : 0.0f;
        loss = __fmaf_rn(diff, diff, loss);
        const float gi = 2.0f * diff * inv;
        g_ws -= gi * p.bg;
        g3[c] = gi * rgb[c] * (1.0f - rgb[c]);
    }
#pragma unroll
    for (int k = 0; k < kHid; ++k) {
        float a = 0.0f;
#pragma unroll
        for (int c = 0; c < kOut; ++c) a = __fmaf_rn(sW3[c * kHid + k], g3[c], a);
        g2[k] = (h2[k] > 0.0f) ? a : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < kHid; ++j) {
        float a = 0.0f;
#pragma unroll
        for (int k = 0; k < kHid; ++k) a = __fmaf_rn(sW2[k * kHid + j], g2[k], a);
        g1[j] = (h1[j] > 0.0f) ? a : 0.0f;
    }
    float gf[kIn];
#pragma unroll
    for (int i = 0; i < kIn; ++i) {
        float a = 0.0f;
#pragma unroll
        for (int j = 0; j < kHid; ++j) a = __fmaf_rn(sW1[j * kIn + i], g1[j], a);
        gf[i] = a;
    }
#pragma unroll
    for (int i = 0; i < kSh; ++i) g_ws = __fmaf_rn(gf[kGeo + i], sh[i], g_ws);
    if (live) {
#pragma unroll
        for (int i = 0; i < kGeo; ++i) p.g_geo_sum[(size_t)r * kGeo + i] = gf[i];
        p.g_weights_sum[r] = g_ws;
    }
    // ---- stash [feature][ray] for the block-level weight-gradient reduction
    float* sF = stash;
    float* sH1 = sF + kIn * kPitch;
    float* sH2 = sH1 + kHid * kPitch;
    float* sG1 = sH2 + kHid * kPitch;
    float* sG2 = sG1 + kHid * kPitch;
    float* sG3 = sG2 + kHid * kPitch;
#pragma unroll
    for (int i = 0; i < kIn; ++i) sF[i * kPitch + t] = f[i];
#pragma unroll
    for (int j = 0; j < kHid; ++j) {
        sH1[j * kPitch + t] = h1[j]; sH2[j * kPitch + t] = h2[j];
        sG1[j * kPitch + t] = g1[j]; sG2[j * kPitch + t] = g2[j];
    }
#pragma unroll
    for (int c = 0; c < kOut; ++c) sG3[c * kPitch + t] = g3[c];
    loss = warp_sum(loss);
    if ((t & 31) == 0) atomicAdd(p.loss, loss * inv);
    __syncthreads();
    auto dot = [&](const float* a, const float* b) {
        float acc = 0.0f;
#pragma unroll 16
        for (int q = 0; q < kRays; ++q) acc = __fmaf_rn(a[q], b[q], acc);
        return acc;
    };
    for (int e = t; e < kW1; e += kRays) {           // dW1[j][i] = sum_r g1[r][j] f[r][i]
        const int j = e / kIn, i = e - j * kIn;
        red_add_f32(p.g_w1 + e, dot(sG1 + j * kPitch, sF + i * kPitch));
    }
    for (int e = t; e < kW2; e += kRays) {           // dW2[k][j] = sum_r g2[r][k] h1[r][j]
        const int k = e / kHid, j = e - k * kHid;
        red_add_f32(p.g_w2 + e, dot(sG2 + k * kPitch, sH1 + j * kPitch));
    }
    for (int e = t; e < kW3; e += kRays) {           // dW3[c][k] = sum_r g3[r][c] h2[r][k]
        const int c = e / kHid, k = e - c * kHid;
        red_add_f32(p.g_w3 + e, dot(sG3 + c * kPitch, sH2 + k * kPitch));
    }
}


// ---- training variant: forward + loss + full backward, 32 rays per 256-thread block ---------------------------------
// The per-ray chains above are ~6 k dependent instructions; with 8192 rays a thread-per-ray kernel leaves the GPU
// almost empty.  Here every layer is a [32 rays x 32 units] tile product computed by the whole block from shared
// memory: thread (ray = lane, unit group = warp) produces 4 outputs per layer, activations are exchanged through
// padded shared tiles, and the weight gradients are dot products over the block's 32 rays (one atomic per entry).
namespace vt {
using namespace vh;
constexpr int kR = 32, kThreads = 256, kP = 33;           // rays per block, threads, tile pitch (conflict-free both ways)
constexpr int oW1 = 0, oW2 = oW1 + kW1, oW3 = oW2 + kW2, oF = oW3 + kW3, oH1 = oF + kR * kP, oH2 = oH1 + kR * kP,
              oG1 = oH2 + kR * kP, oG2 = oG1 + kR * kP, oGF = oG2 + kR * kP, oG3 = oGF + kR * kP, oSH = oG3 + kR * 4,
              oMisc = oSH + kR * 17, kFloats = oMisc + kR * 2 + 8;
constexpr size_t kSmem = (size_t)kFloats * sizeof(float);
}  // namespace vt

__global__ void __launch_bounds__(vt::kThreads) view_head_train_kernel(const ViewHeadParams p) {
    pdl_begin();
    using namespace vt;
    extern __shared__ float smem[];
    float *sW1 = smem + oW1, *sW2 = smem + oW2, *sW3 = smem + oW3, *sF = smem + oF, *sH1 = smem + oH1, *sH2 = smem + oH2;
    float *sG1 = smem + oG1, *sG2 = smem + oG2, *sGF = smem + oGF, *sG3 = smem + oG3, *sSH = smem + oSH;
    float *sWs = smem + oMisc, *sGws = sWs + kR, *sLoss = sGws + kR;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < kW1; i += kThreads) sW1[i] = __ldg(p.w1 + i);
    for (int i = t; i < kW2; i += kThreads) sW2[i] = __ldg(p.w2 + i);
    for (int i = t; i < kW3; i += kThreads) sW3[i] = __ldg(p.w3 + i);
    const uint32_t r0 = blockIdx.x * kR;
    const uint32_t r = r0 + lane;
    const bool live = r < p.N;
    if (t == 0) sLoss[0] = 0.0f;
    // ---- inputs: geometry sums spread over the block, SH by warp 0
    for (int e = t; e < kR * kGeo; e += kThreads) {
        const int rr = e / kGeo, i = e - rr * kGeo;
        sF[rr * kP + i] = (r0 + rr < p.N) ? __ldg(p.geo_sum + (size_t)(r0 + rr) * kGeo + i) : 0.0f;
    }
    if (warp == 0) {
        float sh[kSh], ws = 0.0f;
#pragma unroll
        for (int i = 0; i < kSh; ++i) sh[i] = 0.0f;
        if (live) {
            ws = __ldg(p.weights_sum + r);
            float x = __ldg(p.rays_d + (size_t)r * 3), y = __ldg(p.rays_d + (size_t)r * 3 + 1), z = __ldg(p.rays_d + (size_t)r * 3 + 2);
            const float n = sqrtf(x * x + y * y + z * z);
            x /= n; y /= n; z /= n;
            sh_eval<4, false>(x, y, z, sh, nullptr);
        }
        sWs[lane] = ws;
        sGws[lane] = 0.0f;
#pragma unroll
        for (int i = 0; i < kSh; ++i) { sSH[lane * 17 + i] = sh[i]; sF[lane * kP + kGeo + i] = ws * sh[i]; }
    }
    __syncthreads();
    // ---- layer 1 / 2: thread (ray = lane, units 4 warp .. 4 warp + 3)
    {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < kIn; ++i) {
            const float f = sF[lane * kP + i];
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] = __fmaf_rn(sW1[(warp * 4 + q) * kIn + i], f, a[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) sH1[lane * kP + warp * 4 + q] = fmaxf(a[q], 0.0f);
    }
    __syncthreads();
    {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < kHid; ++j) {
            const float h = sH1[lane * kP + j];
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] = __fmaf_rn(sW2[(warp * 4 + q) * kHid + j], h, a[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) sH2[lane * kP + warp * 4 + q] = fmaxf(a[q], 0.0f);
    }
    __syncthreads();
    // ---- output, loss, d loss / d pre-sigmoid: warps 0..2 <-> channels
    const float inv = p.loss_weight / (3.0f * (float)p.N);       // mean over [N,3]
    if (warp < kOut) {
        float a = 0.0f;
        for (int k = 0; k < kHid; ++k) a = __fmaf_rn(sW3[warp * kHid + k], sH2[lane * kP + k], a);
        const float rgb = 1.0f / (1.0f + expf(-a));
        const float img = rgb + (1.0f - sWs[lane]) * p.bg;
        float diff = 0.0f;
        if (live) {
            p.image[(size_t)r * kOut + warp] = img;
            diff = img - __ldg(p.gt + (size_t)r * kOut + warp);
        }
        const float gi = 2.0f * diff * inv;
        sG3[lane * 4 + warp] = gi * rgb * (1.0f - rgb);
        atomicAdd(sGws + lane, -gi * p.bg);
        const float l = warp_sum(diff * diff);
        if (lane == 0) atomicAdd(sLoss, l);
    }
    __syncthreads();
    if (t == 0) atomicAdd(p.loss, sLoss[0] * inv);
    // ---- backward through the three layers
    {
        const float g0 = sG3[lane * 4], g1 = sG3[lane * 4 + 1], g2 = sG3[lane * 4 + 2];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = warp * 4 + q;
            const float a = __fmaf_rn(sW3[2 * kHid + k], g2, __fmaf_rn(sW3[kHid + k], g1, sW3[k] * g0));
            sG2[lane * kP + k] = (sH2[lane * kP + k] > 0.0f) ? a : 0.0f;
        }
    }
    __syncthreads();
    {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < kHid; ++k) {
            const float g = sG2[lane * kP + k];
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q] = __fmaf_rn(sW2[k * kHid + warp * 4 + q], g, a[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) sG1[lane * kP + warp * 4 + q] = (sH1[lane * kP + warp * 4 + q] > 0.0f) ? a[q] : 0.0f;
    }
    __syncthreads();
    {
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < kHid; ++j) {
            const float g = sG1[lane * kP + j];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = warp * 4 + q;
                if (i < kIn) a[q] = __fmaf_rn(sW1[j * kIn + i], g, a[q]);
            }
        }
        float gws = 0.0f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int i = warp * 4 + q;
            if (i < kGeo) {
                if (live) p.g_geo_sum[(size_t)r * kGeo + i] = a[q];
            } else if (i < kIn) {
                gws = __fmaf_rn(a[q], sSH[lane * 17 + (i - kGeo)], gws);
            }
        }
        if (warp * 4 + 3 >= kGeo) atomicAdd(sGws + lane, gws);
    }
    __syncthreads();
    if (warp == 0 && live) p.g_weights_sum[r] = sGws[lane];
    // ---- weight gradients: dot products over the block's rays
    auto dot = [&](const float* a, const float* b) {
        float acc = 0.0f;
#pragma unroll 8
        for (int q = 0; q < kR; ++q) acc = __fmaf_rn(a[q * kP], b[q * kP], acc);
        return acc;
    };
    for (int e = t; e < kW1; e += kThreads) {           // dW1[j][i] = sum_r g1[r][j] f[r][i]
        const int j = e / kIn, i = e - j * kIn;
        red_add_f32(p.g_w1 + e, dot(sG1 + j, sF + i));
    }
    for (int e = t; e < kW2; e += kThreads) {           // dW2[k][j] = sum_r g2[r][k] h1[r][j]
        const int k = e / kHid, j = e - k * kHid;
        red_add_f32(p.g_w2 + e, dot(sG2 + k, sH1 + j));
    }
    if (t < kW3) {                                      // dW3[c][k] = sum_r g3[r][c] h2[r][k]
        const int c = t / kHid, k = t - c * kHid;
        float acc = 0.0f;
#pragma unroll 8
        for (int q = 0; q < kR; ++q) acc = __fmaf_rn(sG3[q * 4 + c], sH2[q * kP + k], acc);
        red_add_f32(p.g_w3 + t, acc);
    }
}

// Tail of the samvit head's input row (renderer.py:380 / :383): f[r] = [f_sam (written by ray_features_forward), f_image or
// geo_sum, image, depth].  f_image = [geo_sum (15), weights_sum * SH4(d) (16)] with the same normalisation + SH evaluation as
// view_head_kernel.  One thread per ray; replaces torch.cat + the broadcast multiply + a separate SH launch.
__global__ void __launch_bounds__(128) sam_pack_kernel(const float* __restrict__ geo_sum, const float* __restrict__ weights_sum,
                                                       const float* __restrict__ rays_d, const float* __restrict__ image,
                                                       const float* __restrict__ depth, uint32_t N, int use_view_direction,
                                                       float* __restrict__ out, uint32_t out_stride) {
    pdl_begin();
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    float* dst = out + (size_t)r * out_stride;
    uint32_t c = 0;
    for (uint32_t i = 0; i < 15; ++i) dst[c++] = __ldg(geo_sum + (size_t)r * 15 + i);
    if (use_view_direction) {
        float x = __ldg(rays_d + (size_t)r * 3), y = __ldg(rays_d + (size_t)r * 3 + 1), z = __ldg(rays_d + (size_t)r * 3 + 2);
        const float n = sqrtf(x * x + y * y + z * z);
        x /= n; y /= n; z /= n;
        float sh[16];
        sh_eval<4, false>(x, y, z, sh, nullptr);
        const float ws = __ldg(weights_sum + r);
#pragma unroll
        for (uint32_t i = 0; i < 16; ++i) dst[c++] = ws * sh[i];
    }
    for (uint32_t i = 0; i < 3; ++i) dst[c++] = __ldg(image + (size_t)r * 3 + i);
    dst[c] = __ldg(depth + r);
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_view_head(const float* geo_sum, const float* weights_sum, const float* rays_d, const float* gt,
                                const float* w1, const float* w2, const float* w3, float bg, float loss_weight, uint32_t N,
                                float* image, float* loss, float* g_geo_sum, float* g_weights_sum, float* g_w1, float* g_w2,
                                float* g_w3, void* stream) {
    if (N == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(geo_sum); SANERF_REQUIRE_PTR(weights_sum); SANERF_REQUIRE_PTR(rays_d);
    SANERF_REQUIRE_PTR(w1); SANERF_REQUIRE_PTR(w2); SANERF_REQUIRE_PTR(w3); SANERF_REQUIRE_PTR(image);
    const bool train = gt != nullptr;
    if (train) {
        SANERF_REQUIRE_PTR(loss); SANERF_REQUIRE_PTR(g_geo_sum); SANERF_REQUIRE_PTR(g_weights_sum);
        SANERF_REQUIRE_PTR(g_w1); SANERF_REQUIRE_PTR(g_w2); SANERF_REQUIRE_PTR(g_w3);
    }
    ViewHeadParams p{geo_sum, weights_sum, rays_d, gt, w1, w2, w3, bg, loss_weight, N, image, loss, g_geo_sum,
                     g_weights_sum, g_w1, g_w2, g_w3};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t blocks = div_up(N, (uint32_t)vh::kRays);
    if (train) {
        SANERF_LAUNCH(view_head_train_kernel, div_up(N, (uint32_t)vt::kR), vt::kThreads, vt::kSmem, st, p);
        return check_launch("view_head_train_kernel");
    }
    else SANERF_LAUNCH((view_head_kernel<false>), blocks, vh::kRays, vh::kSmemFwd, st, p);
    return check_launch("view_head_kernel");
}

extern "C" int sanerf_sam_pack(const float* geo_sum, const float* weights_sum, const float* rays_d, const float* image,
                               const float* depth, uint32_t N, int use_view_direction, float* out, uint32_t out_stride,
                               void* stream) {
    if (N == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(geo_sum); SANERF_REQUIRE_PTR(weights_sum); SANERF_REQUIRE_PTR(rays_d); SANERF_REQUIRE_PTR(image);
    SANERF_REQUIRE_PTR(depth); SANERF_REQUIRE_PTR(out);
    SANERF_LAUNCH(sam_pack_kernel, div_up(N, 128u), 128, 0, static_cast<cudaStream_t>(stream), geo_sum, weights_sum, rays_d, image,
                  depth, N, use_view_direction, out, out_stride);
    return check_launch("sam_pack_kernel");
}
