/*
 * plume_b200 -- C ABI of the B200-native (sm_100a) UNet smoke-plume segmenter hot path.
 *
 * Boundary note.  The reference project (gridl/kcl-ltss-bioatm) names a UNet in README.md:1-4 and
 * reserves src/models/train_model.py / predict_model.py for it (README.md:44-47), but src/models/ is
 * empty (src/models/__init__.py, 0 bytes).  There is therefore no reference FFI to replace: every
 * entry point below is the operator a `src/models` implementation needs, and the host side in
 * src/models/ binds them through ctypes.  Conventions:
 *   - extern "C", plain pointers and ints; the caller (PyTorch) owns every buffer;
 *   - all device work is enqueued on the caller's stream, no internal synchronisation;
 *   - return 0 on success, a negative code otherwise; plume_last_error() holds the message
 *     (thread-local).  Nothing is swallowed and there is no CPU fallback (the reference's scripts
 *     use `except: continue`, e.g. src/features/plume_identifier_gaussian_profile.py:120-121);
 *   - activations are NHWC bf16.  An activation argument is (pointer to channel 0 of pixel 0,
 *     ld = pixel stride in elements), so a channel slice of a wider buffer (the skip-concat buffer)
 *     is passed without a copy.  Channel offsets and strides must be multiples of 8 (16 bytes).
 *   - 3x3 weights are KRSC ([Cout][3][3][Cin]); transposed-conv weights are [2*2][Cout][Cin].
 */
#ifndef PLUME_B200_H_
#define PLUME_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* plume_stream_t; /* cudaStream_t */

const char* plume_version(void);
const char* plume_last_error(void);
/* Diagnostic word written by a kernel whose bounded barrier wait expired (0 = none). */
int plume_debug_word(void);
/* Deterministic mode (also PLUME_DETERMINISTIC=1 in the environment, read at the first call): every reduction
 * that normally ends in floating-point atomics -- BatchNorm statistics, BatchNorm-backward / bias / head sums, the
 * split-K weight gradients -- writes per-block partial sums to a library-owned scratch buffer and a second kernel
 * adds them in block order; the 3x3 weight-gradient kernel runs a single MMA issuer.  Repeated runs are then
 * bit-identical (tests/test_gpu_determinism.py).  The scratch grows with cudaMalloc on demand, which is not allowed
 * inside a stream capture: run one un-captured step first (Trainer.step_graphed does). */
void plume_set_deterministic(int on);
int plume_get_deterministic(void);
/* SMs left free by the persistent GEMM kernels (default 0; also PLUME_SM_MARGIN).  Data-parallel training sets it to
 * the number of CTAs NCCL may use: a one-CTA-per-SM persistent kernel that finds some SMs held by an all-reduce
 * runs its last CTAs as a second wave. */
void plume_set_sm_margin(int sms);
int plume_get_sm_margin(void);
int plume_num_sms(void);
/* Diagnostics: device buffer [num_sms][8] of int64 that the conv3x3 kernel fills with per-CTA cycle
 * counters of its producer / MMA / epilogue roles (NULL = off, the default). */
void plume_debug_set_prof(long long* buf);

/* ---- tensor-core implicit GEMM (tcgen05 / TMEM / TMA) -------------------------------------- */

/* y = act(conv3x3(x, w) * scale + shift); optional per-channel sum / sum-of-squares of the bf16
 * outputs accumulated (atomically) into stat_sum / stat_sq (fp64[Cout], caller zeroes them; per-tile partial
 * sums are fp32, the cross-tile accumulation is fp64 so that E[y^2] - E[y]^2 survives |mean| >> std).
 * Cin and Cout must be multiples of 64.  scale/shift may be NULL (1 / 0).
 * Fewer real input channels than Cin: pass ldx < Cin (a multiple of 8).  x is then taken as dense with
 * ldx channels per pixel and the weights' remaining input channels read as zero (TMA out-of-bounds fill)
 * -- how the 8-band input feeds the first layer without a padded copy.  Same rule in plume_conv3x3_wgrad,
 * where the gradient of those weight columns comes out zero. */
int plume_conv3x3_fwd(const void* x, int ldx, const void* w_krsc_bf16, const float* scale,
                      const float* shift, int relu, void* y, int ldy, double* stat_sum,
                      double* stat_sq, int N, int H, int W, int Cin, int Cout, plume_stream_t stream);

/* dx = conv3x3(dy, w_dgrad) with w_dgrad[ci][r][s][co] = w[co][2-r][2-s][ci] (plume_pack_conv3x3). */
int plume_conv3x3_dgrad(const void* dy, int lddy, const void* w_dgrad_bf16, void* dx, int lddx, int N,
                        int H, int W, int Cin, int Cout, plume_stream_t stream);

/* K-split count the weight-gradient kernel will use (taps = 9 or 4) -- informational.  The partial
 * sums of the splits are added into dw with fp32 atomic reductions, so no workspace is needed:
 * plume_wgrad_workspace_bytes() returns 0 and the workspace arguments below may be NULL / 0 (they are
 * kept so that a deterministic two-pass reduction can return without an ABI change). */
int plume_wgrad_splits(int N, int H, int W, int taps, int Cin, int Cout);
size_t plume_wgrad_workspace_bytes(int N, int H, int W, int taps, int Cin, int Cout);

/* dw[co][r][s][ci] (fp32) = sum_pixels dy[p][co] * x[p + (r-1, s-1)][ci].  Cin is 64 or a multiple
 * of 128, Cout a multiple of 64.  `accumulate` != 0 adds into dw; 0 zeroes dw first (on the stream).
 * The summation order over pixels is not fixed (atomics): results are reproducible to fp32 rounding. */
int plume_conv3x3_wgrad(const void* x, int ldx, const void* dy, int lddy, float* dw_krsc,
                        int accumulate, void* workspace, size_t workspace_bytes, int N, int H, int W,
                        int Cin, int Cout, plume_stream_t stream);

/* Transposed conv 2x2 stride 2 written into a channel slice of the concat buffer:
 * u[n, 2h+i, 2w+j, co] = sum_ci x[n,h,w,ci] * w[(i*2+j)][co][ci] + bias[co];  x is N x H x W. */
int plume_convT2x2_concat_fwd(const void* x, int ldx, const void* w_ijoc_bf16, const float* bias,
                              void* u, int ldu, int N, int H, int W, int Cin, int Cout,
                              plume_stream_t stream);
/* dx[n,h,w,ci] = sum_{ij,co} du[n,2h+i,2w+j,co] * w[ij][co][ci]; w_dgrad is [Cin][4][Cout]. */
int plume_convT2x2_dgrad(const void* du, int lddu, const void* w_dgrad_bf16, void* dx, int lddx, int N,
                         int H, int W, int Cin, int Cout, plume_stream_t stream);
/* dw[ij][co][ci] (fp32) = sum_pixels du[n,2h+i,2w+j,co] * x[n,h,w,ci].  Cin multiple of 128. */
int plume_convT2x2_wgrad(const void* x, int ldx, const void* du, int lddu, float* dw_ijoc,
                         int accumulate, void* workspace, size_t workspace_bytes, int N, int H, int W,
                         int Cin, int Cout, plume_stream_t stream);

/* ---- weight packing (fp32 master -> bf16 operand layouts) ---------------------------------- */
int plume_pack_conv3x3(const float* w_krsc, void* w_fwd_bf16, void* w_dgrad_bf16, int Cout, int Cin,
                       plume_stream_t stream);
int plume_pack_convT2x2(const float* w_ijoc, void* w_fwd_bf16, void* w_dgrad_bf16, int Cout, int Cin,
                        plume_stream_t stream);

/* All layers of a network in ONE launch.  `descs` is an array of n plume_pack_desc in DEVICE memory (the
 * pointers inside are device pointers); `first_block` is the running sum of plume_pack_blocks() over the
 * preceding entries and `total_blocks` the sum over all of them.  kind 0 = conv3x3 (same layouts as
 * plume_pack_conv3x3), kind 1 = convT2x2 (plume_pack_convT2x2); w_fwd / w_dgrad may be null.  kind | 2: bf16x3
 * mode, each output is the bf16 hi matrix followed by the lo matrix w - hi (2 x taps*Cout*Cin elements). */
typedef struct plume_pack_desc {
  const float* w;
  void* w_fwd_bf16;
  void* w_dgrad_bf16;
  int kind, Cout, Cin, first_block;
} plume_pack_desc;
int plume_pack_blocks(int kind, int Cout, int Cin);
int plume_pack_batch(const plume_pack_desc* descs, int n, int total_blocks, plume_stream_t stream);

/* ---- bandwidth kernels ----------------------------------------------------------------------- */

/* out[p][0:Cd] = concat(in[p][0:Cs], zeros) ; bf16, Cs and Cd multiples of 8. */
int plume_pad_channels(const void* in, int Cs, void* out, int Cd, long long pixels,
                       plume_stream_t stream);

/* BatchNorm (training) from accumulated sums (fp64): mean/var over `count` values per channel, updates the
 * running statistics (unbiased variance, PyTorch convention) and emits the fused scale = gamma*invstd,
 * shift = beta - mean*scale used by the apply kernels, plus mean / invstd for the backward pass. */
int plume_bn_finalize(const double* sum, const double* sq, long long count, const float* gamma,
                      const float* beta, float eps, float momentum, float* running_mean,
                      float* running_var, float* scale, float* shift, float* mean, float* invstd,
                      int C, plume_stream_t stream);
/* Eval-mode folding: scale = gamma/sqrt(var+eps), shift = (bias - mean)*scale + beta. */
int plume_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean,
                       const float* running_var, const float* conv_bias, float eps, float* scale,
                       float* shift, int C, plume_stream_t stream);

/* a = relu?(y*scale + shift) per channel. */
int plume_scale_shift_act(const void* y, int ldy, const float* scale, const float* shift, int relu,
                          void* a, int lda, long long pixels, int C, plume_stream_t stream);
/* Same, fused with the 2x2/stride-2 max pool: writes the full-resolution activation (the skip half
 * of the concat buffer), the pooled activation and the argmax position (uint8 in 0..3, = i*2+j). */
int plume_scale_shift_act_pool(const void* y, int ldy, const float* scale, const float* shift,
                               int relu, void* skip, int ldskip, void* pooled, int ldpooled,
                               uint8_t* argmax, int N, int H, int W, int C, plume_stream_t stream);
/* Plain max pool with argmax, and its backward. */
int plume_maxpool2x2_fwd(const void* x, int ldx, void* y, int ldy, uint8_t* argmax, int N, int H,
                         int W, int C, plume_stream_t stream);
/* dx[n,2h+i,2w+j,c] = (argmax==i*2+j ? dy[n,h,w,c] : 0) + (dskip ? dskip[...] : 0);  H, W are the
 * full-resolution extents. */
int plume_maxpool2x2_bwd(const void* dy, int lddy, const uint8_t* argmax, const void* dskip,
                         int lddskip, void* dx, int lddx, int N, int H, int W, int C,
                         plume_stream_t stream);

/* BatchNorm+ReLU backward.  Pass 1 reduces, per channel, sum(g) and sum(g*xhat) with
 * g = da * [scale*y+shift > 0] (or g = da when relu == 0), xhat = (y-mean)*invstd, into sum_g / sum_gx
 * (fp32[C] scratch the caller zeroes before every backward pass: they must hold THIS batch's sums only).
 * Pass 2 writes dy = scale * (g - sum_g/count - xhat*sum_gx/count), accumulates sum(dy) (the conv bias
 * gradient) and, when dgamma / dbeta are given, hands the parameter gradients over:
 * dgamma = sum_gx, dbeta = sum_g (added to the existing values when accumulate != 0: micro-batching). */
int plume_bn_bwd_reduce(const void* da, int ldda, const void* y, int ldy, const float* scale,
                        const float* shift, const float* mean, const float* invstd, int relu,
                        float* sum_g, float* sum_gx, long long pixels, int C, plume_stream_t stream);
int plume_bn_bwd_apply(const void* da, int ldda, const void* y, int ldy, const float* scale,
                       const float* shift, const float* mean, const float* invstd, int relu,
                       const float* sum_g, const float* sum_gx, void* dy, int lddy, float* sum_dy,
                       float* dgamma, float* dbeta, int accumulate, long long pixels, int C,
                       plume_stream_t stream);
/* ReLU(+bias) backward for norm == "none": dy = da * [a > 0]; accumulates sum(dy) per channel. */
int plume_relu_bwd(const void* da, int ldda, const void* a, int lda, void* dy, int lddy,
                   float* sum_dy, long long pixels, int C, plume_stream_t stream);
/* out[c] += sum_p x[p][c] (fp32; caller zeroes `out`). */
int plume_channel_sum(const void* x, int ldx, float* out, long long pixels, int C,
                      plume_stream_t stream);

/* 1x1 head + sigmoid + BCE/Dice.  logits[p] = sum_c feat[p][c]*w[c] + b (fp32).
 * sums[0..3] += {sum BCE, sum p*t, sum p, sum t} (caller zeroes sums).  C is a multiple of 8, <= 512. */
int plume_head_fwd(const void* feat, int ldf, const float* w, const float* b, const uint8_t* target,
                   float* logits, float* sums, long long pixels, int C, plume_stream_t stream);
/* loss = bce_weight * BCE_mean + dice_weight * (1 - (2*S_pt + eps)/(S_p + S_t + eps)); written to
 * loss_out[0] (and its parts to [1], [2]). */
int plume_head_loss(const float* sums, long long pixels, float bce_weight, float dice_weight,
                    float dice_eps, float* loss_out, plume_stream_t stream);
/* dfeat[p][c] = dlogit[p]*w[c]; dw[c] += sum_p dlogit[p]*feat[p][c]; db += sum_p dlogit[p]
 * (fp32 accumulators, caller zeroes).  grad_scale multiplies the loss gradient (1/world for DP). */
int plume_head_bwd(const void* feat, int ldf, const float* w, const float* logits,
                   const uint8_t* target, const float* sums, float bce_weight, float dice_weight,
                   float dice_eps, float grad_scale, void* dfeat, int lddf, float* dw, float* db,
                   long long pixels, int C, plume_stream_t stream);

/* Fused Adam over one flat fp32 parameter buffer (PyTorch semantics, no weight decay / amsgrad):
 * m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g*g; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).
 * Hyper-parameters are doubles so that (1-beta) and the bias corrections are formed exactly as
 * torch.optim.Adam forms them before rounding to fp32. */
int plume_adam(float* param, const float* grad, float* m, float* v, long long n, double lr, double beta1,
               double beta2, double eps, int step, float grad_scale, plume_stream_t stream);

/* Same update with the step-dependent coefficients in device memory (for CUDA-graph replay of a whole
 * training step): coef[8] = {lr/(1-b1^t), b1, b2, 1-b1, 1-b2, eps, 1/sqrt(1-b2^t), grad_scale}. */
int plume_adam_dev(float* param, const float* grad, float* m, float* v, long long n, const float* coef,
                   plume_stream_t stream);

/* ---- producers fused with the BatchNorm-backward reduction ----------------------------------------------------
 * The gradient these two kernels write (dx of the pool backward, dfeat of the head backward) is the `da` of the
 * BatchNorm layer in front of them.  The _bn variants also accumulate that layer's sum_g / sum_gx (exactly what
 * plume_bn_bwd_reduce(da, y, ...) would add, computed from the rounded gradient they store), reading y once instead
 * of running a separate pass over da and y.  Arguments: those of the plain function, then y / ldy / scale / shift /
 * mean / invstd / relu / sum_g / sum_gx with plume_bn_bwd_reduce's meaning. */
int plume_maxpool2x2_bwd_bn(const void* dy, int lddy, const uint8_t* argmax, const void* dskip, int lddskip,
                            void* dx, int lddx, const void* y, int ldy, const float* scale, const float* shift,
                            const float* mean, const float* invstd, int relu, float* sum_g, float* sum_gx, int N,
                            int H, int W, int C, plume_stream_t stream);
int plume_head_bwd_bn(const void* feat, int ldf, const float* w, const float* logits, const uint8_t* target,
                      const float* sums, float bce_weight, float dice_weight, float dice_eps, float grad_scale,
                      void* dfeat, int lddf, float* dw, float* db, const void* y, int ldy, const float* scale,
                      const float* shift, const float* mean, const float* invstd, int relu, float* sum_g,
                      float* sum_gx, long long pixels, int C, plume_stream_t stream);

/* ---- bf16x3 high-precision mode (BASELINE.json's north_star: "1e-3 (tf32 mode)") ----------------------------
 * Every activation value is stored as hi + lo, two bf16 numbers (hi = bf16(v), lo = bf16(v - hi): 16 significant
 * bits, 4 bytes per value like fp32), in two channel planes per pixel: the pixel stride `ld*` (bf16 elements, a
 * multiple of 16) covers both planes, hi of channel c at [p*ld + c], lo at [p*ld + ld/2 + c].  Each plane is an
 * ordinary NHWC bf16 tensor, so the GEMMs run the same tcgen05 kind::f16 MMAs in three passes into one fp32
 * accumulator, x*w = x_hi*w_hi + x_hi*w_lo + x_lo*w_hi (the dropped lo*lo term is below 2^-17 relative; a single
 * kind::tf32 pass keeps 11 bits and measures 2.6e-3 on the 23-layer network, above the 1e-3 bar).  The bf16 operand
 * copies of the weights are the hi matrix followed by the lo matrix (plume_pack_batch, kind | 2).
 * Same arguments and semantics as the functions without the suffix.  plume_pad_channels_x3 and
 * plume_extract_tiles_x3 read plain bf16 (the input bands) and write the split format. */
int plume_pad_channels_x3(const void* in, int Cs, void* out, int Cd, long long pixels,
                       plume_stream_t stream);
int plume_scale_shift_act_x3(const void* y, int ldy, const float* scale, const float* shift, int relu,
                          void* a, int lda, long long pixels, int C, plume_stream_t stream);
int plume_scale_shift_act_pool_x3(const void* y, int ldy, const float* scale, const float* shift,
                               int relu, void* skip, int ldskip, void* pooled, int ldpooled,
                               uint8_t* argmax, int N, int H, int W, int C, plume_stream_t stream);
int plume_maxpool2x2_fwd_x3(const void* x, int ldx, void* y, int ldy, uint8_t* argmax, int N, int H,
                         int W, int C, plume_stream_t stream);
int plume_maxpool2x2_bwd_x3(const void* dy, int lddy, const uint8_t* argmax, const void* dskip,
                         int lddskip, void* dx, int lddx, int N, int H, int W, int C,
                         plume_stream_t stream);
int plume_bn_bwd_reduce_x3(const void* da, int ldda, const void* y, int ldy, const float* scale,
                        const float* shift, const float* mean, const float* invstd, int relu,
                        float* sum_g, float* sum_gx, long long pixels, int C, plume_stream_t stream);
int plume_bn_bwd_apply_x3(const void* da, int ldda, const void* y, int ldy, const float* scale,
                       const float* shift, const float* mean, const float* invstd, int relu,
                       const float* sum_g, const float* sum_gx, void* dy, int lddy, float* sum_dy,
                       float* dgamma, float* dbeta, int accumulate, long long pixels, int C,
                       plume_stream_t stream);
int plume_relu_bwd_x3(const void* da, int ldda, const void* a, int lda, void* dy, int lddy,
                   float* sum_dy, long long pixels, int C, plume_stream_t stream);
int plume_channel_sum_x3(const void* x, int ldx, float* out, long long pixels, int C,
                      plume_stream_t stream);
int plume_head_fwd_x3(const void* feat, int ldf, const float* w, const float* b, const uint8_t* target,
                   float* logits, float* sums, long long pixels, int C, plume_stream_t stream);
int plume_head_bwd_x3(const void* feat, int ldf, const float* w, const float* logits,
                   const uint8_t* target, const float* sums, float bce_weight, float dice_weight,
                   float dice_eps, float grad_scale, void* dfeat, int lddf, float* dw, float* db,
                   long long pixels, int C, plume_stream_t stream);
int plume_extract_tiles_x3(const void* scene, int Hs, int Ws, int Cs, const int* ys, const int* xs,
                        int count, int T, void* tiles, int Cd, plume_stream_t stream);

int plume_maxpool2x2_bwd_bn_x3(const void* dy, int lddy, const uint8_t* argmax, const void* dskip, int lddskip,
                               void* dx, int lddx, const void* y, int ldy, const float* scale, const float* shift,
                               const float* mean, const float* invstd, int relu, float* sum_g, float* sum_gx, int N,
                               int H, int W, int C, plume_stream_t stream);
int plume_head_bwd_bn_x3(const void* feat, int ldf, const float* w, const float* logits, const uint8_t* target,
                         const float* sums, float bce_weight, float dice_weight, float dice_eps, float grad_scale,
                         void* dfeat, int lddf, float* dw, float* db, const void* y, int ldy, const float* scale,
                         const float* shift, const float* mean, const float* invstd, int relu, float* sum_g,
                         float* sum_gx, long long pixels, int C, plume_stream_t stream);
int plume_conv3x3_fwd_x3(const void* x, int ldx, const void* w, const float* scale, const float* shift,
                         int relu, void* y, int ldy, double* stat_sum, double* stat_sq, int N, int H, int W,
                         int Cin, int Cout, plume_stream_t stream);
int plume_conv3x3_dgrad_x3(const void* dy, int lddy, const void* w_dgrad, void* dx, int lddx, int N, int H,
                           int W, int Cin, int Cout, plume_stream_t stream);
int plume_conv3x3_wgrad_x3(const void* x, int ldx, const void* dy, int lddy, float* dw, int accumulate,
                           void* workspace, size_t workspace_bytes, int N, int H, int W, int Cin, int Cout,
                           plume_stream_t stream);
int plume_convT2x2_concat_fwd_x3(const void* x, int ldx, const void* w, const float* bias, void* u, int ldu,
                                 int N, int H, int W, int Cin, int Cout, plume_stream_t stream);
int plume_convT2x2_dgrad_x3(const void* du, int lddu, const void* w_dgrad, void* dx, int lddx, int N, int H,
                            int W, int Cin, int Cout, plume_stream_t stream);
int plume_convT2x2_wgrad_x3(const void* x, int ldx, const void* du, int lddu, float* dw, int accumulate,
                            void* workspace, size_t workspace_bytes, int N, int H, int W, int Cin, int Cout,
                            plume_stream_t stream);

/* Flat fp32 <-> bf16 casts (round to nearest even): gradient buckets compressed for the data-parallel all-reduce
 * (opt-in, PLUME_GRAD_COMM=bf16).  `in` / `out` 32-byte (fp32) and 16-byte (bf16) aligned. */
int plume_cast_f32_bf16(const float* in, void* out, long long n, plume_stream_t stream);
int plume_cast_bf16_f32(const void* in, float* out, long long n, plume_stream_t stream);

/* ---- tiled large-scene inference -------------------------------------------------------------- */
/* Cut `count` tiles of T x T (NHWC bf16, Cd channels, zero padded past Cs and past the scene edge)
 * out of a scene [Hs][Ws][Cs] bf16; tile k covers origin (ys[k], xs[k]) given as int32 device arrays. */
int plume_extract_tiles(const void* scene, int Hs, int Ws, int Cs, const int* ys, const int* xs,
                        int count, int T, void* tiles, int Cd, plume_stream_t stream);
/* Overlap-stitch by centre crop: each tile contributes the pixels at least `margin` away from its
 * border (or up to the scene edge), thresholded: mask = logit >= logit_threshold (uint8 0/1).
 * Optionally also writes the stitched probabilities (fp32, may be NULL). */
int plume_stitch_threshold(const float* logits, const int* ys, const int* xs, int count, int T,
                           int margin, float logit_threshold, uint8_t* mask, float* prob, int Hs,
                           int Ws, plume_stream_t stream);

/* ---- label geometry: plume hulls -> masks (replaces plume_selector.py:88-116 in_hull / find_plume_aod) ---- */
/* masks[k][r][c] (uint8, `count` windows of Hm x Wm) = 1 where pixel (x, y) = (xs[k] + c, ys[k] + r) lies inside
 * or on the boundary of any of the n convex polygons, else 0.  ys / xs are int32 device arrays (NULL = one window
 * at the origin, i.e. a whole-scene mask).  Polygon i has vertices verts_xy[2*j], verts_xy[2*j+1] for
 * j in [poly_offsets[i], poly_offsets[i+1]), int32 pixel coordinates in counter-clockwise order, and the bounding
 * box bbox[4*i .. 4*i+3] = (xmin, ymin, xmax, ymax).  The test is exact (int64 cross products), the same decision
 * as the reference's Delaunay find_simplex(p) >= 0 for integer coordinates.  All pointers are device memory. */
int plume_rasterize_hulls(const int* verts_xy, const int* poly_offsets, const int* bbox, int n_polys,
                          const int* ys, const int* xs, int count, int Hm, int Wm, uint8_t* masks,
                          plume_stream_t stream);

/* ---- fire -> pixel geolocation (replaces plume_identifier_gaussian_profile.py:85-106, the per-fire search) ---- */
/* out_row_col[2*f], [2*f+1] = row, column of the pixel nearest to fire f among the pixels whose latitude and
 * longitude lie strictly inside the +-half_box_deg box around the fire (haversine distance in float64, first
 * pixel in row-major order on ties), or -1, -1 when the box holds no pixel.  lats / lons are float64 [H][W];
 * workspace is device memory of plume_locate_fires_workspace_bytes(n_fires) bytes.  The reference's edge filter
 * (:108-114) is applied by the host wrapper.  All pointers are device memory. */
size_t plume_locate_fires_workspace_bytes(int n_fires);
int plume_locate_fires(const double* lats, const double* lons, int H, int W, const double* fire_lat,
                       const double* fire_lon, int n_fires, double half_box_deg, void* workspace,
                       size_t workspace_bytes, int* out_row_col, plume_stream_t stream);

/* ---- threshold sweep (replaces plume_identifier_gaussian_profile.py:142-202) ------------------------------ */
/* masks[t][y][x] (uint8 0/1) = binary_dilation(binary_erosion(aod > thresholds[t])) with the cross-shaped
 * footprint (erosion treats pixels beyond the border as set, dilation as unset); aod float32 [H][W]; decides exactly
 * like numpy's float64 comparison against the float64 thresholds (T <= 64), every threshold from one read of the image. */
int plume_threshold_masks(const float* aod, int H, int W, const double* thresholds, int T, uint8_t* masks,
                          plume_stream_t stream);
/* The same for a float64 image (the reference's MAIAC AOD is int16 * 0.001 = float64, tools.py:88): compared in float64.
 * Every aod-taking call of this section has an _f64 twin. */
int plume_threshold_masks_f64(const double* aod, int H, int W, const double* thresholds, int T, uint8_t* masks,
                              plume_stream_t stream);
/* 8-connected components of every mask plane.  labels[t][i] = -1 for background, else the smallest row-major
 * pixel index of the pixel's component (a canonical label); sizes[t][i] = pixel count of the component whose
 * canonical label is i (0 elsewhere).  labels / sizes are int32 [T][H][W]. */
int plume_label_components(const uint8_t* masks, int T, int H, int W, int* labels, int* sizes,
                           plume_stream_t stream);
/* extents[t][f] = size of the component nearest to fire f = (fire_row_col[2f], fire_row_col[2f+1]) inside the
 * (2 win + 1)^2 window around it (Euclidean pixel distance, first pixel in row-major window order on ties), 0 if
 * the window holds no component: find_plume_extents / extract_label of the reference. */
int plume_fire_extents(const int* labels, const int* sizes, int T, int H, int W, const int* fire_row_col,
                       int n_fires, int win, int* extents, plume_stream_t stream);
/* Bit-plane form of the same sweep: masks packed 32 pixels per word, components found over runs of set bits, no
 * dense label plane.  bits is uint32 [T][H][ceil(W / 32)], bit i of word s <-> pixel x = 32 s + i, bits beyond W zero.
 *   plume_threshold_mask_bits : the masks of plume_threshold_masks as bit planes (any T; fp32 comparison against the
 *                               thresholds rounded down to float32, which decides exactly like the float64 comparison).
 *   plume_pack_mask_bits      : byte masks [T][H][W] (non-zero = set) -> bit planes.
 *   plume_bits_extents        : plume_label_components + plume_fire_extents on bit planes: extents[t][f] only.
 *   plume_sweep_extents       : generate_mask_dict + find_plume_extents (gaussian_profile.py:142-179) in one call,
 *                               aod float32 [H][W] -> extents int32 [T][n_fires]; the bit planes live in the workspace.
 * workspace: device memory of at least plume_sweep_workspace_bytes(H, W, T) bytes for both calls that take one. */
size_t plume_sweep_workspace_bytes(int H, int W, int T);
int plume_threshold_mask_bits(const float* aod, int H, int W, const double* thresholds, int T, uint32_t* bits,
                              plume_stream_t stream);
int plume_threshold_mask_bits_f64(const double* aod, int H, int W, const double* thresholds, int T, uint32_t* bits,
                                  plume_stream_t stream);
int plume_pack_mask_bits(const uint8_t* masks, int T, int H, int W, uint32_t* bits, plume_stream_t stream);
int plume_bits_extents(const uint32_t* bits, int T, int H, int W, const int* fire_row_col, int n_fires, int win,
                       void* workspace, size_t workspace_bytes, int* extents, plume_stream_t stream);
/* The component nearest to each fire in the plane chosen for it -- find_plume_mask (gaussian_profile.py:306-331:
 * label(mask), extract_label, labelled_mask == label) without relabelling: component_bits[f] is a bit plane
 * [H][ceil(W / 32)] holding the component nearest to fire f (same rule as plume_fire_extents) of plane plane_of_fire[f]
 * (all zero when the window holds none or the plane index is negative); stats[f] = {area, min_row, min_col,
 * max_row + 1, max_col + 1, root, 0, 0} int32.  workspace must be the one plume_bits_extents / plume_sweep_extents just
 * used on the same bits: it holds the labelling. */
int plume_fire_components(const uint32_t* bits, int T, int H, int W, const int* fire_row_col, const int* plane_of_fire,
                          int n_fires, int win, const void* workspace, size_t workspace_bytes, uint32_t* component_bits,
                          int* stats, plume_stream_t stream);
int plume_sweep_extents(const float* aod, int H, int W, const double* thresholds, int T, const int* fire_row_col,
                        int n_fires, int win, void* workspace, size_t workspace_bytes, int* extents,
                        plume_stream_t stream);
int plume_sweep_extents_f64(const double* aod, int H, int W, const double* thresholds, int T, const int* fire_row_col,
                            int n_fires, int win, void* workspace, size_t workspace_bytes, int* extents,
                            plume_stream_t stream);

/* ---- nearest-valid fill (replaces plume_identifier_gaussian_profile.py:451-461, interpolate_aod_nearest) ----------
 * out[y][x] = aod[y][x] where aod != null_value, else the value of the nearest pixel (Euclidean pixel distance) whose
 * value != null_value -- scipy's NearestNDInterpolator over the valid pixels, evaluated on the whole grid.  Among
 * equidistant valid pixels the first in row-major order is taken (scipy's kd-tree returns one of them in traversal
 * order).  NaN counts as valid (NaN != null_value).  An image without any valid pixel is returned unchanged; the host
 * wrapper raises, as the reference does.  aod / out float32 (or float64: _f64) [H][W], out must not alias aod;
 * workspace: device memory of plume_fill_nearest_workspace_bytes(H, W) bytes. */
size_t plume_fill_nearest_workspace_bytes(int H, int W);
int plume_fill_nearest(const float* aod, int H, int W, float null_value, void* workspace, size_t workspace_bytes,
                       float* out, plume_stream_t stream);
int plume_fill_nearest_f64(const double* aod, int H, int W, double null_value, void* workspace, size_t workspace_bytes,
                           double* out, plume_stream_t stream);

/* ---- UTM projection and nearest-neighbour swath -> grid resampling ------------
 * Replaces /root/reference/src/features/tools.py:9-64 (class utm_resampler), which calls pyproj and
 * pyresample.kd_tree.resample_nearest(radius_of_influence=10000).  All coordinates fp64, degrees / metres.
 * plume_utm_zone_histogram: hist64[z] = number of longitudes whose UTM zone (tools.py:27-28) is z.
 * plume_utm_forward / _inverse: WGS84 UTM of `zone` (k0 0.9996, false easting 500 km, no false northing).
 * plume_resample_nearest_index: the target area is x_size x y_size cells over the extent (outer edges), row 0 at
 * max_y; out_idx[row][col] = flat index of the swath pixel nearest (3-D Cartesian, sphere R = 6370997 m) to the
 * cell centre if closer than `radius` metres, else -1; ties -> smallest index.  `workspace`: device memory of
 * plume_resample_workspace_bytes(...) bytes.  plume_gather_fill: out[i] = idx[i] >= 0 ? src[idx[i]] : fill_value,
 * elements of 4 (float32) or 8 (float64) bytes. */
int plume_utm_zone_histogram(const double* lon, long long n, int* hist64, plume_stream_t stream);
/* Geolocation half of read_modis_aod (tools.py:97-128): lat / lon (float64 [ny][nx], degrees) of the grid
 * x = linspace(x_start, x_stop, nx), y = linspace(y_start, y_stop, ny) metres on the MODIS sinusoidal sphere of `radius`
 * (6371007.181): latitude = y / R, longitude = x / (R cos(latitude)), wrapped into [-180, 180].  The HDF4 parsing of
 * read_modis_aod (pyhdf) is not rebuilt. */
int plume_sinusoidal_grid_latlon(double x_start, double x_stop, double y_start, double y_stop, int ny, int nx,
                                 double radius, double* lat, double* lon, plume_stream_t stream);
int plume_utm_forward(const double* lat, const double* lon, long long n, int zone, double* x, double* y,
                      plume_stream_t stream);
int plume_utm_inverse(const double* x, const double* y, long long n, int zone, double* lat, double* lon,
                      plume_stream_t stream);
size_t plume_resample_workspace_bytes(int n_src, double min_x, double min_y, double max_x, double max_y,
                                      double radius);
int plume_resample_nearest_index(const double* src_lat, const double* src_lon, int n_src, int zone, double min_x,
                                 double min_y, double max_x, double max_y, int x_size, int y_size, double radius,
                                 void* workspace, size_t workspace_bytes, int* out_idx, plume_stream_t stream);
int plume_gather_fill(const void* src, int elem_bytes, const int* idx, long long n, double fill_value, void* out,
                      plume_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PLUME_B200_H_ */
