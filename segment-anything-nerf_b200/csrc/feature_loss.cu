// LayerNorm(256) + MSE against the target feature map, forward AND backward in one kernel (sm_100a).
//
// Replaces the tail of the stage-2 step (nerf/network.py:120-123 `nn.LayerNorm(256)` closing samvit_mlp;
// nerf/utils.py:1100-1106 `pred = samvit.permute(2,0,1)[None]; loss = mse_loss(pred, target)`) and its autograd:
// layer_norm, permute copy, mse, mean, mse_backward, copy, layer_norm_grad_input, GammaBetaBackward (40 us alone)
// = 10 launches / 100 us per 4096-ray step -> one launch.
//
//   y      = (x - mean) * rstd * gamma + beta                       (biased variance, eps inside the root: torch semantics)
//   loss  += sum (y - t)^2 / (M * 256)
//   g_x    = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma,  dy = 2 (y - t) / (M * 256)
//   g_gamma += sum_rows dy * xhat ;  g_beta += sum_rows dy
//
// One warp per row (ray); a lane owns columns 4*lane..+3 and 128+4*lane..+3 (two coalesced float4).  The target is read
// in place from the [1, 256, h, w] map the trainer holds (element (row, c) at c * t_col_stride + row * t_row_stride): a
// block covers 8 consecutive rows, so the 8 warps share each 32-byte sector of a channel plane through L1.
#include "common.cuh"

namespace sanerf {

constexpr uint32_t kFlWarps = 8, kFlN = 256;

__global__ void __launch_bounds__(kFlWarps * 32) layernorm_mse_kernel(
    const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
    const float* __restrict__ target, size_t t_row_stride, size_t t_col_stride, uint32_t M, uint32_t rows_per_block,
    float* __restrict__ y_out, float* __restrict__ loss, float* __restrict__ g_x, float* __restrict__ g_gamma,
    float* __restrict__ g_beta) {
    pdl_begin();
    __shared__ float s_g[kFlWarps][kFlN], s_b[kFlWarps][kFlN];
    __shared__ float s_loss[kFlWarps];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t c0 = lane * 4u, c1 = 128u + lane * 4u;
    const float4 ga = ldg_f4(gamma + c0), gb = ldg_f4(gamma + c1), ba = ldg_f4(beta + c0), bb = ldg_f4(beta + c1);
    const float gam[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    const float bet[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
    const float inv_n = 1.0f / (float)kFlN, inv_total = 1.0f / ((float)M * (float)kFlN);
    float acc_g[8], acc_b[8], acc_loss = 0.0f;
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j) acc_g[j] = acc_b[j] = 0.0f;

    const uint32_t r_begin = blockIdx.x * rows_per_block, r_end = min(r_begin + rows_per_block, M);
    for (uint32_t r = r_begin + warp; r < r_end; r += kFlWarps) {
        const float4 xa = ldg_f4(x + (size_t)r * kFlN + c0), xb = ldg_f4(x + (size_t)r * kFlN + c1);
        float v[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
        float t[8];
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {
            const uint32_t c = (j < 4u) ? c0 + j : c1 + j - 4u;
            t[j] = __ldg(target + (size_t)c * t_col_stride + (size_t)r * t_row_stride);
        }
        float sum = 0.0f;
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) sum += v[j];
        const float mean = warp_sum(sum) * inv_n;
        float sq = 0.0f;
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) { v[j] -= mean; sq = __fmaf_rn(v[j], v[j], sq); }
        const float rstd = rsqrtf(warp_sum(sq) * inv_n + eps);
        float dy[8], y[8], gsum = 0.0f, gx_sum = 0.0f;
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {
            v[j] *= rstd;                                       // xhat
            y[j] = __fmaf_rn(v[j], gam[j], bet[j]);
            const float d = y[j] - t[j];
            acc_loss = __fmaf_rn(d, d, acc_loss);
            dy[j] = 2.0f * d * inv_total;
            acc_g[j] = __fmaf_rn(dy[j], v[j], acc_g[j]);
            acc_b[j] += dy[j];
            const float g = dy[j] * gam[j];
            gsum += g;
            gx_sum = __fmaf_rn(g, v[j], gx_sum);
        }
        const float mg = warp_sum(gsum) * inv_n, mgx = warp_sum(gx_sum) * inv_n;
        float o[8];
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) o[j] = rstd * (dy[j] * gam[j] - mg - v[j] * mgx);
        *reinterpret_cast<float4*>(g_x + (size_t)r * kFlN + c0) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(g_x + (size_t)r * kFlN + c1) = make_float4(o[4], o[5], o[6], o[7]);
        if (y_out != nullptr) {
            *reinterpret_cast<float4*>(y_out + (size_t)r * kFlN + c0) = make_float4(y[0], y[1], y[2], y[3]);
            *reinterpret_cast<float4*>(y_out + (size_t)r * kFlN + c1) = make_float4(y[4], y[5], y[6], y[7]);
        }
    }
    // block reduction of the parameter gradients and the loss, then one reduction per column and block
#pragma unroll
    for (uint32_t j = 0; j < 8; ++j) {
        const uint32_t c = (j < 4u) ? c0 + j : c1 + j - 4u;
        s_g[warp][c] = acc_g[j];
        s_b[warp][c] = acc_b[j];
    }
    acc_loss = warp_sum(acc_loss);
    if (lane == 0) s_loss[warp] = acc_loss;
    __syncthreads();
    const uint32_t c = threadIdx.x;                              // 256 threads = 256 columns
    float tg = 0.0f, tb = 0.0f;
#pragma unroll
    for (uint32_t w = 0; w < kFlWarps; ++w) { tg += s_g[w][c]; tb += s_b[w][c]; }
    red_add_f32(g_gamma + c, tg);
    red_add_f32(g_beta + c, tb);
    if (threadIdx.x == 0) {
        float tl = 0.0f;
#pragma unroll
        for (uint32_t w = 0; w < kFlWarps; ++w) tl += s_loss[w];
        red_add_f32(loss, tl * inv_total);
    }
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_layernorm_mse(const float* x, const float* gamma, const float* beta, float eps, const float* target,
                                    uint64_t t_row_stride, uint64_t t_col_stride, uint32_t M, uint32_t N, float* y_out,
                                    float* loss, float* g_x, float* g_gamma, float* g_beta, void* stream) {
    if (M == 0) return SANERF_OK;
    SANERF_REQUIRE_PTR(x);
    SANERF_REQUIRE_PTR(gamma);
    SANERF_REQUIRE_PTR(beta);
    SANERF_REQUIRE_PTR(target);
    SANERF_REQUIRE_PTR(loss);
    SANERF_REQUIRE_PTR(g_x);
    SANERF_REQUIRE_PTR(g_gamma);
    SANERF_REQUIRE_PTR(g_beta);
    if (N != kFlN) return fail(SANERF_ERR_INVALID_ARG, "layernorm_mse: the feature width must be 256 (network.py:122)");
    const uint32_t rows_per_block = 32;                         // 4 rows per warp: 128 blocks at 4096 rays
    SANERF_LAUNCH(layernorm_mse_kernel, div_up(M, rows_per_block), kFlWarps * 32, 0, static_cast<cudaStream_t>(stream), 
        x, gamma, beta, eps, target, (size_t)t_row_stride, (size_t)t_col_stride, M, rows_per_block, y_out, loss, g_x, g_gamma,
        g_beta);
    return check_launch("layernorm_mse_kernel");
}
