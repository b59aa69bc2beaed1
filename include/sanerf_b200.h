/*
 * sanerf_b200.h — C ABI of libsanerf_b200.so, the sm_100a (B200) implementation of the
 * Segment-Anything-NeRF render hot path.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (passed as
 * `void*`, i.e. a `cudaStream_t`; NULL = legacy default stream) and returns an `int`
 * status (SANERF_OK == 0).  Nothing here allocates, frees or synchronises: the caller
 * owns every buffer (same ownership contract as the reference's pybind functions, which
 * write into caller-allocated tensors — SURVEY §8 b3).  All kernels are enqueued on the
 * given stream and are safe to capture into a CUDA graph.
 *
 * Each function cites the reference interface it replaces (paths relative to the
 * reference checkout).  The Python shims `_gridencoder`, `_shencoder`, `_freqencoder`
 * in `segment-anything-nerf_b200/` re-export the reference's 8 pybind names on top of
 * these symbols; INTEGRATION.md shows the binding.
 */
#ifndef SANERF_B200_H_
#define SANERF_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SANERF_ABI_VERSION 27

#if defined(__GNUC__)
#define SANERF_API __attribute__((visibility("default")))
#else
#define SANERF_API
#endif

/* status codes */
enum {
    SANERF_OK = 0,
    SANERF_ERR_INVALID_ARG = 1,  /* unsupported D / C / dtype / layout (reference: std::runtime_error, gridencoder.cu:392,409) */
    SANERF_ERR_NULL_POINTER = 2, /* a required pointer is NULL (reference: TORCH_CHECK, gridencoder.cu:468-484) */
    SANERF_ERR_CUDA = 3,         /* kernel launch failed; see sanerf_last_error() */
    SANERF_ERR_MISALIGNED = 4    /* a buffer violates the documented alignment */
};

/* element types of tables / encoder outputs */
enum { SANERF_F32 = 0, SANERF_F16 = 1 };

/* layouts of the per-level feature tensor */
enum {
    SANERF_LAYOUT_LBC = 0, /* [L, B, C]  — what the reference kernel writes (gridencoder.cu:103) */
    SANERF_LAYOUT_BLC = 1  /* [B, L*C]   — what GridEncoder.forward returns after its permute copy (grid.py:63) */
};

SANERF_API int sanerf_abi_version(void);
/* Thread-local description of the last non-OK status returned on this thread. */
SANERF_API const char* sanerf_last_error(void);
SANERF_API const char* sanerf_status_string(int status);

/* ------------------------------------------------------------------------------------------
 * Multiresolution hash / tiled grid encoder.
 * Replaces: grid_encode_forward / grid_encode_backward / grad_total_variation /
 * grad_weight_decay (gridencoder/src/gridencoder.h:12-16, gridencoder.cu:467-713).
 *
 *  inputs      f32 [B, D]           coordinates already mapped to [0,1]
 *  embeddings  T   [rows, C]        T = f32 or f16 (dtype)
 *  offsets     i32 [L+1]            level row offsets (grid.py:124-134)
 *  outputs     T   [L,B,C] or [B,L*C] (out_layout)
 *  dy_dx       T   [B, L*D*C] or NULL
 *  S = log2(per_level_scale), H = base resolution, gridtype 0 hash / 1 tiled,
 *  interp 0 linear / 1 smoothstep.  D in {2,3,4,5}, C in {1,2,4,8,16,32}.
 *  Levels >= max_level are left untouched unless zero_tail != 0 (then written as 0).
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets,
                               void* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                               uint32_t max_level, float S, uint32_t H, void* dy_dx,
                               uint32_t gridtype, int align_corners, uint32_t interp, int dtype,
                               int out_layout, int zero_tail, void* stream);

/*  grad             T [L,B,C] or [B,L*C] (grad_layout)
 *  grad_embeddings  T [rows, C]  accumulated into (caller zero-fills; grid.py:83)
 *  dy_dx / grad_inputs  both NULL, or T [B, L*D*C] / T [B, D] (grad_inputs is overwritten)
 */
SANERF_API int sanerf_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                                const int32_t* offsets, void* grad_embeddings, uint32_t B,
                                uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S,
                                uint32_t H, const void* dy_dx, void* grad_inputs, uint32_t gridtype,
                                int align_corners, uint32_t interp, int dtype, int grad_layout,
                                void* stream);

/* In-place total-variation gradient on `grad` (gridencoder.cu:525-668). inputs T [B, D] in [0,1]. */
SANERF_API int sanerf_grad_total_variation(const void* inputs, const void* embeddings, void* grad,
                                const int32_t* offsets, float weight, uint32_t B, uint32_t D,
                                uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                int align_corners, int dtype, void* stream);

/* In-place level-mean weight decay: grad += 2*w*param/level_rows (gridencoder.cu:670-713). B = rows. */
SANERF_API int sanerf_grad_weight_decay(const void* embeddings, void* grad, const int32_t* offsets,
                             float weight, uint32_t B, uint32_t C, uint32_t L, int dtype,
                             void* stream);

/* Ray-composited grid features (stage 2, SAM feature field): replaces `features = self.s_grid(xyzs)` +
 * `f_sam = torch.sum(weights.unsqueeze(-1) * features, dim=-2)` (nerf/renderer.py:302-303, 377) and the autograd chain
 * behind them (expand + multiply, permute copy, zeros_like, scatter: gridencoder/grid.py:74-95) by one kernel per
 * direction; no per-sample [N*T, L*C] matrix exists.  Hash grid, linear interpolation, align_corners = false, D = 3,
 * fp32, C in {2,4,8}.
 *  x01      f32 [N*T, 3]  samples of N dense rays (T each), already mapped to [0,1]^3
 *  weights  f32 [N*T]     compositing weights (constants here: the density field is frozen in stage 2, main.py:255-262)
 *  out      f32 [N, L*C]  sum_i weights[r,i] * encode(x01[r,i])
 *  g_out    f32 [N, L*C]  incoming gradient;  grad_embeddings f32 [rows, C] is accumulated into (caller zero-fills);
 *           the backward scatters levels [level_begin, level_end) only (0, L = all): a data-parallel trainer launches it
 *           per level group and starts the reduce-scatter of a finished slice of the table while the next group runs
 */
SANERF_API int sanerf_ray_features_forward(const float* x01, const float* weights, const float* embeddings,
                                const int32_t* offsets, uint32_t N, uint32_t T, uint32_t C, uint32_t L, float S,
                                uint32_t H, float* out, uint32_t out_stride, void* stream);
SANERF_API int sanerf_ray_features_backward(const float* x01, const float* weights, const float* g_out,
                                 const int32_t* offsets, uint32_t N, uint32_t T, uint32_t C, uint32_t L, float S,
                                 uint32_t H, float* grad_embeddings, uint32_t level_begin, uint32_t level_end,
                                 uint32_t g_stride, void* stream);
/* out_stride / g_stride: floats between consecutive rays of out / g_out (0 = L*C; otherwise >= L*C and a multiple of 4), so
 * that the features can be written straight into - and their gradient read straight from - the wider input row of the
 * samvit head (renderer.py:380) without a concatenation copy.
 * sanerf_sam_pack writes the rest of that row: out[r, 0..] = [geo_sum (15), weights_sum * SH4(normalised rays_d) (16, only
 * when use_view_direction), image (3), depth (1)], rows out_stride floats apart. */
SANERF_API int sanerf_sam_pack(const float* geo_sum, const float* weights_sum, const float* rays_d, const float* image,
                    const float* depth, uint32_t N, int use_view_direction, float* out, uint32_t out_stride, void* stream);

/* Debug / parity entry point (no reference equivalent — SURVEY §8 c7): for every sample,
 * level < L and corner < 2^D write the table row (relative to the level's offset) the
 * forward kernel would read, plus the level geometry computed on the device.
 *  rows      u32 [B, L, 2^D]
 *  geometry  u32 [L, 4] = {resolution, level_rows, uses_hash, covered_dims} (may be NULL)
 */
SANERF_API int sanerf_grid_dump_indices(const float* inputs, const int32_t* offsets, uint32_t* rows,
                             uint32_t* geometry, uint32_t B, uint32_t D, uint32_t L, float S,
                             uint32_t H, uint32_t gridtype, int align_corners, void* stream);

/* ------------------------------------------------------------------------------------------
 * Spherical-harmonics encoder, degree 1..8 (shencoder/src/shencoder.h:9-10, shencoder.cu:27-439).
 *  inputs f32 [B,3] (unit vectors), outputs f32 [B, deg*deg], dy_dx f32 [B, 3*deg*deg] or NULL.
 * `ray_stride` > 0 broadcasts: sample b uses inputs[b / ray_stride] (one direction per ray;
 * the reference re-evaluates the same direction T times, network.py:237). 0/1 = plain.
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D,
                             uint32_t degree, float* dy_dx, uint32_t ray_stride, void* stream);
/* grad_inputs[b,d] += sum_ch grad[b,ch]*dy_dx[b,d,ch]  (accumulates, like shencoder.cu:378) */
SANERF_API int sanerf_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D,
                              uint32_t degree, const float* dy_dx, float* grad_inputs,
                              void* stream);

/* ------------------------------------------------------------------------------------------
 * Frequency (positional) encoder (freqencoder/src/freqencoder.h:7,10, freqencoder.cu:30-129).
 *  inputs f32 [B,D], outputs f32 [B,C], C = D + 2*D*deg, column order [x, sin 2^0 x, cos 2^0 x, ...]
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg,
                               uint32_t C, float* outputs, void* stream);
SANERF_API int sanerf_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D,
                                uint32_t deg, uint32_t C, float* grad_inputs, void* stream);

/* ------------------------------------------------------------------------------------------
 * trunc_exp activation (activation.py:5-18): y = exp(x); dx = g * exp(clamp(x,-15,15)).
 * `stride`/`offset` let it read column `offset` of a [n, stride] matrix (network.py:226 takes
 * channel 0 of the 16-wide grid_mlp output) — pass stride=1, offset=0 for a flat vector.
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_trunc_exp_forward(const float* x, float* y, uint64_t n, uint32_t stride,
                             uint32_t offset, void* stream);
SANERF_API int sanerf_trunc_exp_backward(const float* g, const float* x, float* dx, uint64_t n,
                              uint32_t stride, uint32_t offset, void* stream);

/* ------------------------------------------------------------------------------------------
 * Per-ray front-to-back compositing (nerf/renderer.py:309-338, :377-383, :453).
 *
 * Samples are PACKED: ray r owns samples [ray_offsets[r], ray_offsets[r+1]).  The reference's
 * dense [N,T] block is the special case ray_offsets == NULL with uniform count T.
 *
 *  sigmas, deltas, ts   f32 [M]        density, interval length, mid-point distance
 *  feats                f32 [M, C]     per-sample channels (colour feature, SAM feature, ...); C may be 0.
 *                       Row stride `feat_stride` floats (0 = C): lets the kernel read columns of a wider
 *                       matrix in place, e.g. the 15 geometry features at columns 1..15 of the 16-wide MLP output
 *  last_sample_opaque   != 0: the last sample of each ray gets delta*sigma := +inf (renderer.py:314-316)
 *  t_thresh             early ray termination: samples whose incoming transmittance < t_thresh
 *                       get weight 0 (0 disables; the reference never terminates — SURVEY §8 c5)
 * outputs:
 *  weights      f32 [M]      alpha_i * T_i with NaN -> 0 (renderer.py:318-326)
 *  weights_sum  f32 [N]      sum_i w_i
 *  depth        f32 [N]      sum_i w_i t_i
 *  out          f32 [N, C]   sum_i w_i feats_i
 *  n_alive      i32 [N]      number of samples with T_i >= t_thresh (may be NULL)
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_composite_forward(const float* sigmas, const float* deltas, const float* ts,
                             const float* feats, uint32_t feat_stride, const int32_t* ray_offsets,
                             uint32_t N, uint32_t T, uint32_t C, int last_sample_opaque, float t_thresh,
                             float* weights, float* weights_sum, float* depth, float* out,
                             int32_t* n_alive, void* stream);

/* Backward of the above.  Incoming gradients (any may be NULL = zero):
 *  g_weights [M], g_weights_sum [N], g_depth [N], g_out [N,C]
 * Produces grad_sigmas [M] and grad_feats [M,C] with row stride `grad_feat_stride` (0 = C; grad_feats may be
 * NULL).  deltas / ts carry no gradient (bins are detached, renderer.py:275).
 */
SANERF_API int sanerf_composite_backward(const float* sigmas, const float* deltas, const float* ts,
                              const float* feats, uint32_t feat_stride, const int32_t* ray_offsets,
                              uint32_t N, uint32_t T, uint32_t C, int last_sample_opaque, float t_thresh,
                              const float* weights, const float* g_weights,
                              const float* g_weights_sum, const float* g_depth, const float* g_out,
                              float* grad_sigmas, float* grad_feats, uint32_t grad_feat_stride,
                              void* stream);

/* ------------------------------------------------------------------------------------------
 * Final-level density activation + compositing of the field head's [N*T, 16] output in one kernel per direction:
 * column 0 = density logit -> sigma = exp(.) (activation.py:5-18), columns 1..15 = geometry feature composited with the
 * weights of renderer.py:309-338.  Dense rays only (T samples each).  sigma (out, [N*T]) and n_alive may be NULL.
 * Backward: incoming g_weights [N*T], g_weights_sum [N], g_depth [N], g_out [N,15], g_sigma_direct [N*T] (any may be
 * NULL = zero); writes the whole 16-wide gradient row g_head [N*T,16] (column 0 through trunc_exp's clamped derivative).
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_head_composite_forward(const float* head, const float* deltas, const float* ts, uint32_t N,
                                  uint32_t T, int last_sample_opaque, float t_thresh, float* sigma, float* weights,
                                  float* weights_sum, float* depth, float* out, int32_t* n_alive, void* stream);
SANERF_API int sanerf_head_composite_backward(const float* head, const float* deltas, const float* ts, uint32_t N,
                                   uint32_t T, int last_sample_opaque, float t_thresh, const float* g_weights,
                                   const float* g_weights_sum, const float* g_depth, const float* g_out,
                                   const float* g_sigma_direct, float* g_head, void* stream);

/* ------------------------------------------------------------------------------------------
 * Proposal sampling chain, one kernel per level (nerf/renderer.py:122-139 near/far, :250-253 spacing,
 * :263-286 edges -> mid points / intervals / positions, :60-69 contraction, :84-119 sample_pdf,
 * gridencoder/grid.py:156 unit-cube mapping).
 *  rays_o, rays_d f32 [N,3]; aabb f32 [6]; cam_near_far f32 [.,2] with row stride cnf_stride (0 = one row
 *  for all rays) or NULL; noise f32 uniform [0,1) [N, T+1] or NULL (= no perturbation).
 * outputs: bins [N,T+1] in [0,1], t_mid [N,T], deltas [N,T], x01 [N,T,3] = (contract(o + d t) + bound)/(2 bound).
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_sample_uniform(const float* rays_o, const float* rays_d, const float* aabb, float min_near,
                          const float* cam_near_far, uint32_t cnf_stride, const float* noise, uint32_t N,
                          uint32_t T, int contract, float bound, float* bins, float* t_mid, float* deltas,
                          float* x01, void* stream);
/* prev_bins [N,T0+1], prev_weights [N,T0]: the previous level's edges and compositing weights.  With prev_sigmas /
 * prev_deltas [N,T0] given (prev_sigmas != NULL) the compositing of the previous level (renderer.py:309-326) is fused in:
 * its weights are computed here, written to prev_weights_out [N,T0] and used for the resampling (prev_weights unused). */
SANERF_API int sanerf_sample_pdf(const float* rays_o, const float* rays_d, const float* aabb, float min_near,
                      const float* cam_near_far, uint32_t cnf_stride, const float* prev_bins,
                      const float* prev_weights, uint32_t T0, const float* noise, uint32_t N, uint32_t T,
                      int contract, float bound, float* bins, float* t_mid, float* deltas, float* x01,
                      const float* prev_sigmas, const float* prev_deltas, int last_sample_opaque,
                      float* prev_weights_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Ray generation (get_rays, nerf/utils.py:145-279): pixel centres, un-normalised pinhole directions rotated by the
 * camera-to-world pose, origin = its translation.  poses f32 row-major 4x4 with stride pose_stride floats (0: one
 * pose for all rays, 16: one per ray); intrinsics f32 (fx, fy, cx, cy) with stride 0 or 4; inds i64 [N] flat pixel
 * indices (row * W + col) or NULL for pixels 0..N-1; outputs rays_o, rays_d f32 [N,3].
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_generate_rays(const float* poses, uint32_t pose_stride, const float* intrinsics,
                         uint32_t intr_stride, const int64_t* inds, uint32_t W, uint32_t N, float* rays_o,
                         float* rays_d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Proposal density, fused: hash-grid encode (D=3, F=2, L<=8, fp32) -> Linear(2L,16) -> ReLU -> Linear(16,1) ->
 * trunc_exp (nerf/network.py:211-219, 248-252).  w1 f32 [16, 2L], w2 f32 [1,16] (nn.Linear layout, no bias).
 * enc_out f32 [B,2L] or NULL: the encoding, kept for the backward (`enc`; NULL there = re-gather it).
 * Backward accumulates into grad_table [rows,2], grad_w1 [16,2L], grad_w2 [16] (caller zero-fills); consecutive
 * samples of a ray that fall into the same cell are merged in the warp before the reductions are issued.
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_prop_density_forward(const float* x01, const float* table, const int32_t* offsets,
                                const float* w1, const float* w2, uint32_t B, uint32_t L, float S,
                                uint32_t H, float* sigma, float* enc_out, void* stream);
SANERF_API int sanerf_prop_density_backward(const float* x01, const float* table, const int32_t* offsets,
                                 const float* w1, const float* w2, uint32_t B, uint32_t L, float S,
                                 uint32_t H, const float* enc, const float* g_sigma, float* grad_table,
                                 float* grad_w1, float* grad_w2, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sampling regularisers: loss value AND d loss / d weights in one kernel each.
 *  proposal loss (nerf/renderer.py:30-57) of ONE proposal level (t_p [N,Tp+1], w_p [N,Tp]) against the final
 *  level (t_ref [N,Tr+1], w_ref [N,Tr]); distortion loss (renderer.py:17-27 + torch_efficient_distloss).
 *  loss_out: one float, ACCUMULATED into (caller zero-fills); g_* may be NULL.  `weight` (lambda_proposal /
 *  lambda_distort, main.py) multiplies the loss contribution AND the gradient.
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_proposal_loss(const float* t_ref, const float* w_ref, uint32_t Tr, const float* t_p,
                         const float* w_p, uint32_t Tp, uint32_t N, float weight, float* loss_out, float* g_wp,
                         void* stream);
SANERF_API int sanerf_distortion_loss(const float* bins, const float* w, uint32_t T, uint32_t N, float weight,
                           float* loss_out, float* g_w, void* stream);

/* ------------------------------------------------------------------------------------------
 * Deferred-shading view head + photometric loss, forward AND backward in one kernel (renderer.py:333-345, 358;
 * network.py:107, 237; Trainer.train_step nerf/utils.py:897-930):
 *   f = [geo_sum (15), weights_sum * SH4(normalize(rays_d)) (16)];  rgb = sigmoid(W3 relu(W2 relu(W1 f)));
 *   image = rgb + (1 - weights_sum) * bg;  loss += loss_weight * mean((image - gt)^2).
 * gt == NULL: forward only (image).  Otherwise also writes g_geo_sum [N,15], g_weights_sum [N] and ACCUMULATES
 * loss (one float) and the weight gradients g_w1 [32,31], g_w2 [32,32], g_w3 [3,32] (caller zero-fills).
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_view_head(const float* geo_sum, const float* weights_sum, const float* rays_d, const float* gt,
                     const float* w1, const float* w2, const float* w3, float bg, float loss_weight, uint32_t N,
                     float* image, float* loss, float* g_geo_sum, float* g_weights_sum, float* g_w1, float* g_w2,
                     float* g_w3, void* stream);

/* ------------------------------------------------------------------------------------------
 * fp32-parity GEMM on the tensor cores (tcgen05.mma.kind::tf32, accumulators in tensor memory) with the elementwise
 * neighbours of a Linear layer folded in: the building block of the SAM feature head
 *   samvit_mlp = SkipConnMLP(163 -> 256 x 4 -> 256, bias, leaky ReLU, skip at layer 2)  (nerf/network.py:36-75, 120-123),
 * replacing its nn.Linear / cuBLAS SGEMM + bias + activation (+ autograd: data-gradient GEMM, weight-gradient GEMM,
 * activation backward) launches.
 *   C[M,N] (op)= A . B^T ;  A f32 [M,K] (lda) or, a_trans != 0, stored [K,M];  B f32 [N,K] (ldb) or, b_trans != 0, [K,N]
 *   epilogue 0: C  = act(acc + bias[n])          bias may be NULL; act != 0: leaky ReLU with `slope`
 *   epilogue 1: C  = acc * (mask[m,n] > 0 ? 1 : slope) for n < mask_cols (mask f32 [M, >= mask_cols], ldm);
 *               colsum f32 [mask_cols] or NULL: colsum[n] += sum_m C[m,n] (the bias gradient of the layer below)
 *   epilogue 2: C += acc  (reductions; K may be split over k_splits CTAs: weight gradients over the batch rows)
 *   precision 0: 3-term tf32 split (fp32 parity, ~2^-21 per product), 1: one tf32 pass.
 * Row strides need not be multiples of 4 (nn.Linear(163, .) / (419, .) weights): unaligned rows take scalar loads.
 * sanerf_colsum_add: out[n] += sum_m X[m,n] (bias gradient).
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_gemm_tc(const float* A, uint32_t lda, int a_trans, const float* B, uint32_t ldb, int b_trans, float* C,
                   uint32_t ldc, uint32_t M, uint32_t N, uint32_t K, uint32_t k_splits, int epilogue,
                   const float* bias, int act, float slope, const float* mask, uint32_t ldm, uint32_t mask_cols,
                   float* colsum, int precision, void* stream);
/* Strided device-to-device copy of a [rows, cols] fp32 matrix (a copy node, no kernel): used to keep copies of the two samvit
 * weights whose rows are not 16-byte multiples (163 and 419 columns) with a padded leading dimension, which makes them
 * addressable by the TMA (csrc/gemm_tma.cu). */
SANERF_API int sanerf_copy_rows(float* dst, uint32_t ld_dst, const float* src, uint32_t ld_src, uint32_t rows, uint32_t cols,
                     void* stream);
SANERF_API int sanerf_colsum_add(const float* X, uint32_t ld, uint32_t M, uint32_t N, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tail of the stage-2 step in one kernel: LayerNorm(256) closing samvit_mlp (nerf/network.py:120-123) + the MSE against
 * the target feature map (nerf/utils.py:1100-1106), forward and backward.
 *   x f32 [M,256] (MLP output), gamma / beta f32 [256], eps (torch default 1e-5);
 *   target element (row, c) at target[c * t_col_stride + row * t_row_stride] (the [1,256,h,w] map read in place:
 *   t_col_stride = h*w, t_row_stride = 1);  y_out f32 [M,256] or NULL (the normalised features);
 *   loss (1 float) += mean((y - t)^2);  g_x f32 [M,256] is overwritten;  g_gamma / g_beta f32 [256] are accumulated into.
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_layernorm_mse(const float* x, const float* gamma, const float* beta, float eps, const float* target,
                         uint64_t t_row_stride, uint64_t t_col_stride, uint32_t M, uint32_t N, float* y_out,
                         float* loss, float* g_x, float* g_gamma, float* g_beta, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused Adam over one flat fp32 buffer (main.py:296 Adam(eps=1e-15), :312-313 LambdaLR 0.1^min(it/iters,1)).
 * sanerf_adam_schedule advances the device-side step counter and writes dyn[4] = {lr_t, 1-b1^t, 1-b2^t, 1-ema_t}
 * with ema_t = min(ema_decay, (1+t)/(10+t)) (torch_ema's warm-up, nerf/utils.py:616 with main.py:316 ema_decay=0.95; only
 * used by the optional per-step EMA - the reference updates its EMA per epoch, see sanerf_ema_update);
 * when `gate` is given it is set to 1 ("the step starting now leaves a gradient for the deferred ranges").
 * sanerf_adam_step applies the update (gradient pre-multiplied by grad_scale, e.g. 1/world_size) and can zero
 * the gradient in the same pass; with `gate` != NULL it returns without touching anything when *gate == 0 (a deferred
 * range whose update was already applied by a flush); with `ema` != NULL (same length as params) it also performs
 * shadow -= (1-ema_t)*(shadow - param_new) in the same pass (per-step EMA; optional).  Both are CUDA-graph replayable.
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_adam_schedule(int32_t* step, float* dyn, float lr0, float beta1, float beta2,
                         float decay_iters, int32_t* gate, float ema_decay, void* stream);
SANERF_API int sanerf_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, uint64_t n,
                     const float* dyn, float beta1, float beta2, float eps, float grad_scale,
                     int zero_grad, const int32_t* gate, float* ema, void* stream);
/* torch_ema.ExponentialMovingAverage.update() over a flat buffer: shadow -= one_minus_decay * (shadow - params), with
 * one_minus_decay = 1 - min(decay, (1 + k) / (10 + k)) for the k-th update.  The reference updates once per EPOCH
 * (nerf/utils.py:1862) or per 16 GUI steps (:1627); passing `ema` to sanerf_adam_step instead updates per optimizer step. */
SANERF_API int sanerf_ema_update(float* shadow, const float* params, uint64_t n, float one_minus_decay, void* stream);
/* Half-precision table with an fp32 master copy (BASELINE configs[4], T = 2^22 fp16 parameters; the reference reaches
 * half tables through grid.py:43-46): grads16 [n] half (the scatter target / its reduce-scattered sum), master /
 * exp_avg / exp_avg_sq [n] fp32; the updated parameters are written to master AND, rounded, to params16 in one pass.
 * n % 8 == 0, 16-byte aligned buffers. */
SANERF_API int sanerf_adam_step_half(float* master, void* params16, void* grads16, float* exp_avg, float* exp_avg_sq,
                          uint64_t n, const float* dyn, float beta1, float beta2, float eps, float grad_scale,
                          int zero_grad, void* stream);

/* ------------------------------------------------------------------------------------------
 * Uniform [0,1) jitter for the samplers (renderer.py:269 `torch.rand_like(bins)`, :101 `torch.rand_like(u)`): out f32 [n]
 * filled with Philox4x32-10 numbers; state = uint32 [2] on the device, zero-initialised once, {call number, block
 * arrival count}: the call number advances by one per launch, so a replayed CUDA graph draws fresh numbers.  Optionally
 * clears zero[0..zero_n) (the step's loss accumulator) in the same launch.
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_uniform_fill(float* out, uint64_t n, uint64_t seed, uint32_t* state, float* zero, uint32_t zero_n,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel update fused with its exchange over NVLink / NVSwitch peer memory (csrc/symm_adam.cu): replaces the
 * gradient all-reduce of DistributedDataParallel + torch.optim.Adam on every rank (main.py:296).
 * All `world` ranks of one node call it with the same range; param / grad are this rank's flat buffers, which live in
 * symmetric memory: *_mc = their multicast (NVLS) addresses or NULL (independently: NULL selects the peer addresses), *_peers = HOST arrays [world] of every rank's
 * unicast address of the same buffers (index = rank), flag_peers = HOST array [world] of the symmetric flag arrays
 * (uint32 [256][8], zero-initialised once), epoch = this rank's uint32 [256] (zero-initialised once), error = uint32 [1]
 * set to 1 if a peer never arrived (bounded spin, the kernel then returns instead of hanging).
 * Per call: barrier; rank r sums the gradient of slice r of [start, stop) over all ranks (multimem.ld_reduce or peer
 * loads), applies Adam (grad_scale, typically 1/world; optional EMA as in sanerf_adam_step) to its slice of
 * exp_avg / exp_avg_sq / ema and writes the new parameters into EVERY rank's buffer; barrier; every rank clears its own
 * gradient range.  Afterwards all ranks hold bit-identical parameters and a zero gradient in [start, stop);
 * optimizer state and EMA are valid only for the rank's own slice.  gate as in sanerf_adam_step.  start, stop: multiples
 * of 4.  The blocks of one index wait for each other across ranks, so the grid must be co-resident: channel 0 (<= 160
 * blocks), 1 or 2 (<= 32 blocks each) select disjoint flag / epoch slots, so that calls running concurrently on different
 * streams do not share any.  threads: 32..1024 per block.  world: 2, 4 or 8.
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_symm_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, float* ema,
                          void* param_mc, void* grad_mc, const uint64_t* param_peers, const uint64_t* grad_peers,
                          const uint64_t* flag_peers, uint32_t* epoch, uint32_t* error, uint64_t start, uint64_t stop,
                          uint32_t world, uint32_t rank, const float* dyn, float beta1, float beta2, float eps,
                          float grad_scale, const int32_t* gate, uint32_t blocks, uint32_t threads, uint32_t channel, void* stream);

/* ------------------------------------------------------------------------------------------
 * Field head on the tensor cores (tcgen05 / TMEM), fused with the hash-grid gather:
 *   out[B,16] = W3 relu(W2 relu(W1 enc(x)))   with enc = GridEncoder(L=16, F=2, hash, linear) of nerf/network.py:102,
 *   W1 [64,32], W2 [64,64], W3 [16,64] = grid_mlp (network.py:103; nn.Linear layout [out,in], no bias, ReLU).
 * Replaces `h = self.grid(x); f = self.grid_mlp(h)` (network.py:223-224): one kernel, activations never leave the SM.
 *  x01 f32 [B,3] in [0,1]^3 (grid.py:156 mapping already applied), or NULL to read the encoding from enc_in [B,32];
 *  enc_out f32 [ceil(B/128)*128, 32] or NULL: the encoding, saved for the backward (values bit-identical to
 *  sanerf_grid_encode_forward).  The three saved activations are private to this pair of kernels and are stored
 *  TILE-CHUNK-MAJOR: element (b, k) of a W-wide matrix at float offset (((b/128)*(W/4) + k/4)*128 + b%128)*4 + k%4,
 *  so that a warp touches 512 contiguous bytes per instruction in both kernels;
 *  precision 0: fp32 parity (every product as a 3-term tf32 split, fp32 accumulate), 1: one tf32 pass.
 *  h1_out / h2_out f32 [ceil(B/128)*128, 64] or both NULL: the post-ReLU activations of layers 1 and 2 (same layout).
 * Backward: from enc [B,32], h1, h2 [B,64] and g_out [B,16] produces g_enc [B,32] (feed it to
 * sanerf_grid_encode_backward, layout [B,L*C]) and ACCUMULATES the weight gradients into g_w1 / g_w2 / g_w3
 * (caller zero-fills).  With x01 / offsets / g_table given (x01 != NULL) the hash-grid scatter of
 * sanerf_grid_encode_backward is FUSED into the kernel's last epilogue: the gradient of the encoding goes from tensor
 * memory straight into red.global.add.v2.f32 on g_table [rows,2] (warp-aggregated; g_enc may then be NULL and is not
 * written).  Data gradients: A operand in tensor memory x transposed weight; weight gradients: MN-major
 * operands in the 128-byte-swizzle / 32-byte-atom layout, accumulated in tensor memory over the CTA's tiles.
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_field_head_forward(const float* x01, const float* table, const int32_t* offsets, float S,
                              uint32_t H, const float* enc_in, const float* w1, const float* w2, const float* w3,
                              uint32_t B, float* enc_out, float* h1_out, float* h2_out, float* out, int precision,
                              void* stream);
SANERF_API int sanerf_field_head_backward(const float* enc, const float* h1, const float* h2, const float* g_out,
                               const float* w1, const float* w2, const float* w3, uint32_t B, float* g_enc,
                               const float* x01, const int32_t* offsets, float S, uint32_t H, float* g_table,
                               float* g_w1, float* g_w2, float* g_w3, int precision, void* stream);

/* ------------------------------------------------------------------------------------------
 * Early ray termination that skips WORK (inference frames): the final level is evaluated front to back in chunks of
 * chunk_len = 8 samples, for the rays still alive only.  The reference declares --T_thresh (main.py:71-72) and never reads
 * it; the rule used is SURVEY 8 c5: sample k of a ray contributes iff T_k = exp(-sum_{j<k} delta_j sigma_j) >= t_thresh.
 *  sanerf_field_head_forward_chunk: as sanerf_field_head_forward (inference form: nothing saved) for samples
 *    [chunk*chunk_len, (chunk+1)*chunk_len) of the rays ray_list[0 .. *list_count) (device-side count, <= max_rays);
 *    x01 [N,T,3] and out [N,T,16] keep their dense layout, rows of other samples are not touched.
 *  sanerf_head_composite_chunk: trunc_exp + compositing of that chunk (same arithmetic as sanerf_head_composite_forward)
 *    ACCUMULATED into per-ray state — optical [N] (carried sum of delta*sigma), weights_sum [N], depth [N], out [N,15],
 *    n_alive [N], all zero-filled by the caller before chunk 0 — and the rays whose transmittance after the chunk is still
 *    >= t_thresh are appended to next_list (next_count zero-filled by the caller; both NULL for the last chunk).
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_field_head_forward_chunk(const float* x01, const float* table, const int32_t* offsets, float S,
                              uint32_t H, const float* w1, const float* w2, const float* w3, float* out, int precision,
                              const uint32_t* ray_list, const uint32_t* list_count, uint32_t max_rays, uint32_t chunk,
                              uint32_t chunk_len, uint32_t T, void* stream);
SANERF_API int sanerf_head_composite_chunk(const float* head, const float* deltas, const float* ts, const uint32_t* ray_list,
                              const uint32_t* list_count, uint32_t max_rays, uint32_t T, uint32_t chunk, uint32_t chunk_len,
                              int last_sample_opaque, float t_thresh, uint32_t* next_list, uint32_t* next_count,
                              float* optical, float* weights_sum, float* depth, float* out, int32_t* n_alive, void* stream);

/* ------------------------------------------------------------------------------------------
 * Diagnostics (no reference equivalent): one tcgen05 tile product D[M,N] with operands staged the way the fused
 * MLP kernels stage them.  mode 0: A[M,K], B[N,K] both K-major; 1: At[K,M], Bt[K,N] both MN-major (M = 64 or 128;
 * 128-byte swizzle with 32-byte atoms, the only MN-major form 32-bit operands have);
 * 2: as 0 with A resident in TMEM; 3: as 0 with the 3xTF32 split.  M = 128 otherwise; N, K multiples of 16 / 8.
 * ---------------------------------------------------------------------------------------- */
SANERF_API int sanerf_umma_selftest(int mode, uint32_t M, uint32_t N, uint32_t K, const float* A, const float* B,
                         float* D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SANERF_B200_H_ */
