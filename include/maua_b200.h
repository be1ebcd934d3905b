/*
 * maua_b200.h -- C ABI of libmaua_b200.so: the B200-native (sm_100a) VGG-19 neural-style inner loop
 * of JCBrouwer/maua-style.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; these entry points are what a
 * ctypes binding inside the reference's loss.py / models.py / optim.py would call (INTEGRATION.md shows
 * the stub).  Each entry cites the reference code (relative to the reference repo root) it replaces.
 *
 * Conventions
 *   - every function returns 0 on success or a negative maua_status code and never throws;
 *     maua_last_error() returns a thread-local human readable message for the last failure;
 *   - all pointers are DEVICE pointers owned by the caller unless stated otherwise, fp32, 16-byte aligned;
 *   - work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); no call synchronises
 *     the device unless stated otherwise;
 *   - images at the API boundary are NCHW [B,3,H,W] like the reference (BGR, 0-255, mean-subtracted,
 *     load.py:21-32); feature maps inside the library are NHWC with values rounded to TF32;
 *   - there is no CPU fallback: on a non-sm_100 device every compute entry fails with MAUA_ERR_ARCH.
 */
#ifndef MAUA_B200_H
#define MAUA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MAUA_API __attribute__((visibility("default")))
#else
#define MAUA_API
#endif

typedef void* maua_stream_t; /* cudaStream_t */

typedef enum maua_status {
    MAUA_STATUS_OK = 0,
    MAUA_STATUS_BAD_ARG = -1,
    MAUA_STATUS_CUDA = -2,
    MAUA_STATUS_ARCH = -3,
    MAUA_STATUS_OOM = -4,
    MAUA_STATUS_STATE = -5
} maua_status;

/* implementation selector of the GEMM-shaped kernels: 0 = tcgen05/TMEM/TMA (product path),
 * 1 = naive SIMT cross-check (tests / debugging only; never selected by the library itself). */
#define MAUA_IMPL_TC 0
#define MAUA_IMPL_REF 1
/* tests only: tcgen05 path with the CTA grouping forced -- single CTAs (cta_group::1) or CTA pairs (cta_group::2);
 * MAUA_IMPL_TC picks per launch. */
#define MAUA_IMPL_TC_1CTA 2
#define MAUA_IMPL_TC_2CTA 3
/* Exact-arithmetic mode: every GEMM-shaped launch runs with un-rounded FP32 operands on the CUDA cores (chunk sums in
 * fp32 FFMA, chunks added in fp64 -- csrc/conv_fp32.cu), the Gram / covariance is centred like loss.py:87-89 and summed
 * in fp64, and no activation / gradient / weight is rounded to TF32 anywhere.  Everything else -- plan logic, ReLU sign
 * bitmaps, max-pool arg-max recomputation, folded StyleLoss backward, pooling, image-side tail, optimizers -- is the
 * code the product path runs.  ~30x slower than MAUA_IMPL_TC; it exists to show (tests) that the product path differs
 * from the reference's fp32 results by TF32 operand rounding and nothing else, and as a "reference-grade" option
 * (MAUA_PRECISION=fp32). */
#define MAUA_IMPL_FP32 4

MAUA_API int maua_abi_version(void);
MAUA_API const char* maua_last_error(void);
/* 0 if `device` is an sm_100 GPU usable by this library, MAUA_STATUS_ARCH otherwise. */
MAUA_API int maua_device_check(int device);

/* ------------------------------------------------------------------------------------------------
 * Layout / weight preparation (once per model load; replaces nothing in the reference, which keeps
 * NCHW tensors and OIHW weights for cuDNN -- models.py:351-363)
 * ---------------------------------------------------------------------------------------------- */
/* w_oihw [Cout][Cin][3][3] -> GEMM layout, TF32-rounded.
 *   dgrad = 0: out[co][tap*Cin + ci]  = w[co][ci][ky][kx]      (forward, tap = ky*3+kx)
 *   dgrad = 1: out[ci][tap*Cout + co] = w[co][ci][2-ky][2-kx]  (input-gradient: rotated + transposed) */
MAUA_API int maua_prep_conv_weights(const float* w_oihw, float* out, int cout, int cin, int dgrad,
                                    maua_stream_t stream);
/* Same layouts with the TF32 rounding optional (round_tf32 = 0: the operands of MAUA_IMPL_FP32). */
MAUA_API int maua_prep_conv_weights_ex(const float* w_oihw, float* out, int cout, int cin, int dgrad, int round_tf32,
                                       maua_stream_t stream);
/* Host-side launch plan of a conv3x3 / pointwise / aux-GEMM launch on a device with `sms` SMs (no GPU needed): plan6 =
 * {BN, MT, CTA-group size, whole tiles, tail tiles, items per tail tile}.  tail_mode selects how the last, partial wave of the
 * persistent grid is computed: 0 whole tiles, 1 K-split (parts meet in a workspace, deterministic order), 2 two half-N items
 * per tile (the default of a plan: independent items, no hand-over). */
MAUA_API int maua_conv_tile_plan(int h, int w, int cin, int cout, int ntaps, int k2, int sms, int tail_mode, int* plan6);
MAUA_API int maua_nchw_to_nhwc(const float* src, float* dst, int b, int c, int h, int w, int round_tf32,
                               maua_stream_t stream);
MAUA_API int maua_nhwc_to_nchw(const float* src, float* dst, int b, int c, int h, int w, maua_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * VGG feature stack kernels -- models.py:116-132 build_sequential (Conv2d 3x3 pad 1 -> ReLU(inplace)
 * -> MaxPool2d/AvgPool2d(2,2)) and their autograd backward (optim.py:213), weights frozen
 * (models.py:443-445) so input-gradients only.
 * ---------------------------------------------------------------------------------------------- */
/* y = [relu](conv3x3(x) + bias); x NHWC [B,H,W,Cin] (Cin % 32 == 0), y NHWC [B,H,W,Cout] (Cout % 64 == 0),
 * wg from maua_prep_conv_weights(dgrad=0).  Output rounded to TF32. */
MAUA_API int maua_conv3x3_fwd(const float* x, const float* wg, const float* bias, float* y, int b, int h, int w,
                              int cin, int cout, int relu, int impl, maua_stream_t stream);
/* gx = (conv3x3_dgrad(gy) [+ aux_f @ aux_d^T + aux_bias] [+ cont_coef*(cont_f - cont_t)]) * (mask_src > 0)
 *   gy NHWC [B,H,W,Cout]; wd from maua_prep_conv_weights(dgrad=1); gx NHWC [B,H,W,Cin].
 *   mask_src (optional) is the post-ReLU activation that fed this conv (ReLU backward of the layer below).
 *   aux_* (optional): StyleLoss backward folded in as extra GEMM k-steps -- aux_f NHWC [B,H,W,Cin] is the
 *   tapped feature map, aux_d [Cin][Cin] the scaled symmetric (G - A) matrix, aux_bias [Cin] the covariance
 *   mean correction (loss.py:87-89).  cont_* (optional): ContentLoss gradient (loss.py:53-59), cont_coef is
 *   a device scalar.  gy may be NULL (wd too) when only the loss terms contribute (last tap layer). */
MAUA_API int maua_conv3x3_dgrad(const float* gy, const float* wd, float* gx, int b, int h, int w, int cout, int cin,
                                const float* mask_src, const float* aux_f, const float* aux_d,
                                const float* aux_bias, const float* cont_f, const float* cont_t,
                                const float* cont_coef, int round_tf32, int impl, maua_stream_t stream);
/* Sign bitmap of an NHWC activation x [npix][c] (c % 32 == 0): bits[pixel * (c/32) + ch/32] bit (ch % 32) = x > 0.  The
 * plan keeps one per conv layer (written by the forward epilogue) as the ReLU mask of the backward pass: autograd's
 * threshold_backward reads the fp32 activation instead (models.py:130 ReLU(inplace)), 32x the bytes. */
MAUA_API int maua_relu_mask_bits(const float* x, uint32_t* bits, long npix, int c, maua_stream_t stream);
/* maua_conv3x3_dgrad with the ReLU mask given as a sign bitmap; this variant runs the TMA-store epilogue
 * (registers -> swizzled shared-memory box -> cp.async.bulk.tensor store). */
MAUA_API int maua_conv3x3_dgrad_bits(const float* gy, const float* wd, float* gx, int b, int h, int w, int cout, int cin,
                                     const uint32_t* mask_bits, const float* aux_f, const float* aux_d,
                                     const float* aux_bias, int round_tf32, int impl, maua_stream_t stream);
/* First layer (Cin = 3): image NCHW [B,3,H,W] -> NHWC [B,H,W,Cout], bias + ReLU; w is plain OIHW fp32. */
MAUA_API int maua_conv_first_fwd(const float* img, const float* w_oihw, const float* bias, float* y, int b, int h,
                                 int w, int cout, maua_stream_t stream);
/* First layer dgrad + image-side tail: gimg NCHW [B,3,H,W] = dgrad(gy) + tv_coef * dTV/dimg
 *   + temp_coef * w * (img*w - temp_target).  tv_coef / temp_coef are device scalars (NULL = term absent).
 *   Runs as a per-pixel tensor-core contraction (64 channels -> 27 tap columns) followed by a 9-tap gather;
 *   workspace: maua_conv_first_dgrad_workspace_bytes(b, h, w) device bytes. */
MAUA_API size_t maua_conv_first_dgrad_workspace_bytes(int b, int h, int w);
MAUA_API int maua_conv_first_dgrad(const float* gy, const float* w_oihw, float* gimg, int b, int h, int w, int cout,
                                   const float* img, const float* tv_coef, const float* temp_target,
                                   const float* temp_weights, const float* temp_coef, void* workspace,
                                   maua_stream_t stream);
/* 2x2/2 pooling on NHWC, floor semantics; avg = 0 max (first maximum wins ties), 1 average. */
MAUA_API int maua_pool2x2_fwd(const float* x, float* y, int b, int h, int w, int c, int avg, maua_stream_t stream);
/* gx = unpool(gy) * (x > 0) [+ addend * (x > 0)]; x is the post-ReLU pre-pool activation. */
MAUA_API int maua_pool2x2_bwd(const float* x, const float* gy, const float* addend, float* gx, int b, int h, int w,
                              int c, int avg, int round_tf32, maua_stream_t stream);

/* k x k / stride 1 / pad k/2 convolution for k = 1, 3, 5 through the same tcgen05 implicit-GEMM kernel (k = 1: NIN's "cccp"
 * layers, models.py:85-110; k = 5: NIN conv2, models.py:90 -- a halo of 2 pixels, 5 box loads per channel chunk, 5 vertical taps
 * per box).  wg from maua_prep_conv_weights_k: forward [cout][k*k*cin] (tap-major), dgrad = rotated + transposed, optionally
 * TF32-rounded.  With the dgrad weights and cin / cout swapped the same call is the input gradient. */
MAUA_API int maua_prep_conv_weights_k(const float* w, float* out, int cout, int cin, int ks, int dgrad, int round_tf32,
                                      maua_stream_t stream);
MAUA_API int maua_conv_kxk_fwd(const float* x, const float* wg, const float* bias, float* y, int b, int h, int w, int cin,
                               int cout, int ks, int relu, int impl, maua_stream_t stream);

/* NIN backbone (models.py:74-113): the layer shapes that are not 3x3 / pad 1 or 1x1 GEMMs, as direct fp32 convolutions.
 * out NHWC [b][oh][ow][cout] = act(conv(in, w OIHW [cout][cin][ks][ks], stride, pad) + bias), oh = (h + 2 pad - ks) / stride + 1;
 * `in` is NHWC [b][h][w][cin], or the NCHW image when in_nchw (conv1: 11x11 / stride 4, models.py:83; conv2: 5x5 / pad 2, :90). */
MAUA_API int maua_conv_direct_fwd(const float* in, int in_nchw, const float* w, const float* bias, float* out, int b, int h,
                                  int wd, int cin, int cout, int ks, int stride, int pad, int relu, int round_tf32,
                                  maua_stream_t stream);
/* [cout][cin][ks][ks] -> [cin][cout][ks][ks] rotated by 180 degrees: with these weights and pad' = ks - 1 - pad the input gradient
 * of a stride-1 convolution is maua_conv_direct_fwd of the output gradient (autograd's conv backward-data, models.py:90). */
MAUA_API int maua_conv_direct_flip_weights(const float* w, float* out, int cout, int cin, int ks, maua_stream_t stream);
/* Input gradient of the strided, unpadded image layer: gout NHWC [b][oh][ow][cout] -> gimg NCHW [b][3][h][w] (models.py:83). */
MAUA_API int maua_conv_direct_dgrad_image(const float* gout, const float* w, float* gimg, int b, int h, int wd, int cout, int ks,
                                          int stride, maua_stream_t stream);
/* MaxPool2d / AvgPool2d((3,3), (2,2), (0,0), ceil_mode=True) on NHWC (models.py:77-80): output (h < 2 ? 0 : (h - 2) / 2 + 1), the
 * last window clipped at the border (the average divides by the clipped size); backward like maua_pool2x2_bwd, gather form over
 * the <= 4 windows that contain a pixel, first maximum in row-major window order. */
MAUA_API int maua_pool3x3_fwd(const float* x, float* y, int b, int h, int w, int c, int avg, maua_stream_t stream);
MAUA_API int maua_pool3x3_bwd(const float* x, const float* gy, const float* addend, float* gx, int b, int h, int w, int c,
                              int avg, int round_tf32, maua_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Loss kernels -- loss.py
 * ---------------------------------------------------------------------------------------------- */
/* GramMatrix.forward (loss.py:67-91) / nelement: gram[c][d] = sum_p X[p][c] X[p][d] / (C*P) for one image,
 * f NHWC [P][C] (P = H*W), optionally mean-centred (use_covariance).  Symmetric tensor-core SYRK,
 * split-K over pixels.  mean_out [C] (may be NULL unless use_covariance).  workspace: maua_gram_workspace_bytes. */
MAUA_API size_t maua_gram_workspace_bytes(int c);
MAUA_API int maua_gram(const float* f, long p, int c, int use_covariance, float* gram, float* mean_out,
                       void* workspace, int impl, maua_stream_t stream);
/* StyleLoss in "loss" mode (loss.py:141-181): mse = mean((gram - target)^2); *loss_out = value_scale * mse;
 * diff[c][d] = gram - target (kept for the backward). */
MAUA_API int maua_style_loss_fwd(const float* gram, const float* target, int c, float value_scale, float* loss_out,
                                 float* diff, void* workspace, maua_stream_t stream);
/* Backward coefficient matrix: aux_d = (*coef) * 4/(C^3 P) * diff (TF32-rounded), aux_bias = -aux_d @ mean
 * (covariance only, else NULL).  coef is a device scalar.  Feed to maua_conv3x3_dgrad(aux_*). */
MAUA_API int maua_style_loss_bwd_prep(const float* diff, const float* mean, int c, long p, const float* coef,
                                      float* aux_d, float* aux_bias, maua_stream_t stream);
/* ContentLoss value (loss.py:53-59): *loss_out = value_scale * mean((x*weights? - target)^2).
 * weights (optional) has `plane` elements broadcast over n/plane leading slices (NCHW temporal loss). */
MAUA_API int maua_content_loss_fwd(const float* x, const float* weights, const float* target, long n, long plane,
                                   float value_scale, float* loss_out, void* workspace, maua_stream_t stream);
/* TVLoss value (loss.py:224-233) on an NCHW image. */
MAUA_API int maua_tv_loss_fwd(const float* img, int planes, int h, int w, float strength, float* loss_out,
                              void* workspace, maua_stream_t stream);
MAUA_API size_t maua_reduce_workspace_bytes(void);

/* ------------------------------------------------------------------------------------------------
 * Optimizer kernels -- optim.py:180-196 (torch.optim.Adam / torch.optim.LBFGS on the pastiche pixels)
 * ---------------------------------------------------------------------------------------------- */
MAUA_API int maua_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, float lr,
                            float beta1, float beta2, float eps, int step, maua_stream_t stream);
/* Same update with the 1-based step number read from device memory (*step_dev), so that a CUDA-graph replay of the
 * optimisation loop (optim.py:240) sees a fresh bias correction every iteration. */
MAUA_API int maua_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, float lr,
                                float beta1, float beta2, float eps, const int* step_dev, maua_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Image-side operations either side of the loop: with these the pastiche stays in
 * HBM from one scale / frame to the next and only 8-bit RGB crosses PCIe.  Images are NCHW planes [planes][H][W].
 * ---------------------------------------------------------------------------------------------- */
/* F.interpolate(x, mode="bilinear", align_corners=False) -- style.py:38-41, :47-49, :57-66 (img_img scale
 * transition), :205-212, :241-255, :284-286 (vid_img), load.py:211-213.  scale_h / scale_w: the `scale_factor` the
 * caller passed to F.interpolate (the source step is then 1/scale_factor, like ATen), or 0 when it passed `size`
 * (step = in/out).  The caller computes h_out = floor(h_in * scale_factor) itself. */
MAUA_API int maua_resize_bilinear(const float* src, float* dst, int planes, int h_in, int w_in, int h_out, int w_out,
                                  double scale_h, double scale_w, maua_stream_t stream);
/* F.grid_sample(x, grid, padding_mode="border") (bilinear, align_corners=False) -- style.py:223, :279: warps the
 * previous frame's pastiche along the optical flow.  grid: [h_out][w_out][2] normalised (x, y) in [-1, 1]. */
MAUA_API int maua_grid_sample_border(const float* src, const float* grid, float* dst, int planes, int h_in, int w_in,
                                     int h_out, int w_out, maua_stream_t stream);
/* load.preprocess (load.py:21-32): RGB -> BGR, 0-255, minus the mean pixel (103.939, 116.779, 123.68).
 * _u8: from a PIL image's bytes (HWC uint8, device memory); _f32: from a ToTensor()-style CHW float image in [0,1]. */
MAUA_API int maua_preprocess_u8(const uint8_t* rgb_hwc, float* bgr_chw, int h, int w, maua_stream_t stream);
MAUA_API int maua_preprocess_f32(const float* rgb_chw, float* bgr_chw, int h, int w, maua_stream_t stream);
/* load.deprocess (load.py:47-52) down to the bytes ToPILImage would hold: HWC uint8 RGB (device memory). */
MAUA_API int maua_deprocess_u8(const float* bgr_chw, uint8_t* rgb_hwc, int h, int w, maua_stream_t stream);
/* out = a * x + b * y -- style.py:290 `(1 - temporal_blend) * blend_image + temporal_blend * pastiche`.  out may alias. */
MAUA_API int maua_blend(const float* x, const float* y, float* out, long n, float a, float b, maua_stream_t stream);
/* utils.match_histogram (utils.py:88-151; called at style.py:24, :67, :71 and the video drivers): transfer of the
 * channel means and covariances of the style image(s) onto an image, `mean_s [ Qs Qt^-1 (x - mu_t) + mu_s ]` with
 * Q = the symmetric square root of the 3x3 channel covariance + eps*I (utils.get_histogram, utils.py:88-93).  Three
 * asynchronous steps, no host round trip:
 *   maua_image_moments      one pass over a CHW [3][h][w] image: moments[10] = { N, sum x_c (3), sum x_c x_d (6, upper
 *                           triangle row-major) } in fp64, reduced in a fixed order (workspace: maua_reduce_workspace_bytes)
 *   maua_hist_match_coefs   target moments + n_sources consecutive source moment sets -> affine[12] = { M (3x3 row-major),
 *                           b (3) } (replaces the two th.symeig / th.inverse / th.mm chains, utils.py:124-135)
 *   maua_color_affine       dst_c = sum_d M[c][d] src_d + b_c over the image (utils.py:135-138); dst may alias src
 * The reference also perturbs its inputs with 1e-3 * randn (utils.py:120-121); callers reproduce the effect on the
 * statistics by passing eps + 1e-6, the per-pixel output noise is not reproduced. */
MAUA_API int maua_image_moments(const float* img_chw, int h, int w, double* moments, void* workspace, maua_stream_t stream);
MAUA_API int maua_hist_match_coefs(const double* target_moments, const double* source_moments, int n_sources, double eps,
                                   float* affine, maua_stream_t stream);
MAUA_API int maua_color_affine(const float* src_chw, float* dst_chw, int h, int w, const float* affine, maua_stream_t stream);

typedef struct maua_lbfgs maua_lbfgs_t;
/* L-BFGS state for one n-element parameter vector (torch.optim.LBFGS semantics without line search:
 * history ring of `history` (s, y) pairs, lr, first-step t = min(1, 1/|g|_1) * lr, ys > 1e-10 update gate,
 * H_diag = ys / y.y, gtd > -tolerance_change halts).  Owns 2*history + 4 vectors of n floats. */
MAUA_API int maua_lbfgs_create(long n, int history, float lr, float tolerance_change, maua_lbfgs_t** out);
MAUA_API void maua_lbfgs_destroy(maua_lbfgs_t* s);
/* One L-BFGS iteration given the gradient at the current parameters: updates history, computes the two-loop
 * direction and applies param += t * d.  Fully asynchronous (all scalars stay on the device). */
MAUA_API int maua_lbfgs_step(maua_lbfgs_t* s, float* param, const float* grad, maua_stream_t stream);
/* Back to the state right after maua_lbfgs_create (empty history, first-step rule armed) without re-allocating the
 * history rings: what a caller that optimises many images of one size with the same state object does between images
 * (the reference constructs a new torch.optim.LBFGS per optimize() call, optim.py:180-191). */
MAUA_API int maua_lbfgs_reset(maua_lbfgs_t* s, maua_stream_t stream);
/* Debug / test read-back (synchronises): n_iter, history length, halted flag. */
MAUA_API int maua_lbfgs_query(maua_lbfgs_t* s, int* n_iter, int* hist_len, int* halted, maua_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Plan: the whole feval of optim.py:201-238 (net(pastiche) -> sum of module losses -> backward to the
 * image) for one network description, as a fixed launch sequence with library-owned workspaces.
 * ---------------------------------------------------------------------------------------------- */
typedef struct maua_plan maua_plan_t;

#define MAUA_MAX_LAYERS 32
#define MAUA_MAX_TAPS 16

typedef enum maua_tap_kind { MAUA_TAP_STYLE = 0, MAUA_TAP_CONTENT = 1 } maua_tap_kind;
/* MAUA_MODE_EXTERNAL (style taps only): the forward pass keeps the tap's feature map but computes nothing for it; the
 * caller forms the loss from the features of SEVERAL plans (the frames of an img_vid window, loss.py:141-181 with B > 1:
 * per-frame static Grams and the [B*C, B*C] dynamic Gram) and hands this frame's backward GEMM term back with
 * maua_plan_set_tap_fold before maua_plan_backward. */
typedef enum maua_tap_mode { MAUA_MODE_NONE = 0, MAUA_MODE_CAPTURE = 1, MAUA_MODE_LOSS = 2, MAUA_MODE_EXTERNAL = 3 } maua_tap_mode;

/* Network description: models.py:135-139 channel_list entry truncated after the last tapped ReLU
 * (models.py:382).  channels[i] > 0: conv3x3(channels[i]) + ReLU;  channels[i] == 0: 2x2 pool. */
typedef struct maua_net_desc {
    int n_entries;
    int channels[MAUA_MAX_LAYERS];
    int avg_pool;                       /* models.py:119-122 */
    const float* weights[MAUA_MAX_LAYERS]; /* per conv entry: OIHW fp32 device pointer (copied + transformed) */
    const float* biases[MAUA_MAX_LAYERS];
    int n_taps;
    int tap_relu_index[MAUA_MAX_TAPS];  /* 0-based index of the ReLU (= conv count - 1) the loss module follows */
    int tap_kind[MAUA_MAX_TAPS];        /* maua_tap_kind; taps must be ordered by relu index (content before style
                                           at the same index, as models.py:411-431 inserts them) */
    /* Channel-pruned VGG-16 (models.py:136 "VGG-16p": 24, 22, 41, 51, 108, 89, 111, 184, 276, 228, 512...): the caller
     * zero-pads every conv's weights / bias to a channel count the tcgen05 kernels tile (multiples of 64) and names the
     * REAL count here; the Gram 1/(C*H*W), the nn.MSELoss means and the
     * gradient scales use it, so padded and unpadded networks give the same losses and gradients.  0 = channels[i]. */
    int norm_channels[MAUA_MAX_LAYERS];
    /* NIN (models.py:74-113, nin_dict :140-172): per conv entry 0 = 3x3 / pad 1 (VGG), 1 = 1x1 ("cccp" layers), 5 = 5x5 / pad 2
     * (conv2), 11 = 11x11 / stride 4 / no padding (conv1, image layer only).  pool_kind 0 = 2x2 / stride 2 (floor),
     * 1 = 3x3 / stride 2 / ceil_mode (models.py:77-80). */
    int conv_kind[MAUA_MAX_LAYERS];
    int pool_kind;
} maua_net_desc;

/* Per-call, per-tap state: targets are caller-owned tensors so the Python loss modules can expose them
 * as `.target` (loss.py:40,124-126). */
typedef struct maua_tap_io {
    int mode;                 /* maua_tap_mode */
    int use_covariance;       /* style only */
    float value_scale;        /* style: strength*(1+vsf) ; content: strength   (B = 1) */
    float capture_weight;     /* style capture: target (+)= capture_weight * gram   (blend_weight, loss.py:148-151) */
    int capture_accumulate;   /* 0: overwrite target, 1: add */
    float* target;            /* style: [C][C]; content: NHWC [H][W][C] at the tap resolution */
    long target_elems;        /* content: elements of `target` (0 = not captured yet); shape mismatch => skipped */
} maua_tap_io;

typedef struct maua_image_io {
    /* TVLoss (module index 0 when tv_weight > 0, models.py:369-373) */
    int tv_mode;              /* maua_tap_mode: NONE or LOSS (value is assigned on every forward, loss.py:232) */
    float tv_strength;
    /* temporal ContentLoss on the image (models.py:375-379, loss.py:46-54) */
    int temporal_mode;        /* NONE / CAPTURE / LOSS */
    float temporal_strength;
    float* temporal_target;   /* NCHW [3][H][W] */
    long temporal_target_elems;
    const float* temporal_weights; /* [H][W] or NULL */
} maua_image_io;

MAUA_API int maua_plan_create(int device, const maua_net_desc* desc, maua_plan_t** out);
/* One stage of the layer-wise multidevice split (models.py:503-566 ModelParallel / setup_multi_device): a plan over
 * entries [entry_begin, entry_end) of `desc` on `device`.  A stage must begin with a conv entry.  Tap indices stay
 * global: taps outside the range are ignored by this stage.  Weight pointers may live on another device. */
MAUA_API int maua_plan_create_stage(int device, const maua_net_desc* desc, int entry_begin, int entry_end,
                                    maua_plan_t** out);
MAUA_API void maua_plan_destroy(maua_plan_t* plan);
/* Bytes of device memory the plan currently owns (weights + workspaces), for capacity planning. */
/* Number of times the plan (re)allocated one of its workspaces (activation arena, sign bitmaps, gradient buffers): they grow
 * with the image size and keep their size afterwards.  A CUDA graph captured from maua_plan_forward / _backward bakes their
 * addresses in; a caller that replays captured iterations (optim.GraphedIteration) re-captures when this number changed. */
MAUA_API int maua_plan_workspace_generation(const maua_plan_t* plan);
MAUA_API size_t maua_plan_device_bytes(const maua_plan_t* plan);

/* net(image): forward to the last tap.  image NCHW [1,3,H,W].  For taps in CAPTURE mode updates `target`;
 * for taps in LOSS mode writes the module loss value into losses_out[tap] (device, n_taps + 2 floats:
 * [0..n_taps) taps in order, [n_taps] TV, [n_taps+1] temporal).  Entries of modules not in LOSS mode are 0.
 * keep_for_backward != 0 retains activations so maua_plan_backward can follow. */
MAUA_API int maua_plan_forward(maua_plan_t* plan, const float* image, int h, int w, const maua_tap_io* taps,
                               const maua_image_io* image_io, float* losses_out, int keep_for_backward,
                               maua_stream_t stream);
/* Backward of the last forward: grad_coefs (device, n_taps + 2 floats) are the per-module gradient weights
 * (upstream dL/dloss_i already combined with strength / ScaleGradients semantics by
 * maua_loss_grad_coefs); writes d(sum)/d(image) NCHW into grad_image. */
MAUA_API int maua_plan_backward(maua_plan_t* plan, const float* grad_coefs, float* grad_image,
                                maua_stream_t stream);
/* Stage variants (models.py:517-525 ModelParallel.forward and the autograd mirror of its `.to(device)` hops).
 *   forward : `input` is the NCHW image (entry_begin == 0) or the NHWC [h][w][C] activation handed over by the previous
 *             stage (h, w = its extent); boundary_out (NULL for the last stage) receives this stage's last activation --
 *             it may be memory of the NEXT stage's device (peer access enabled, maua_enable_peer_access): the producing
 *             conv epilogue / pool kernel stores into it directly over NVLink, there is no separate copy.
 *   backward: grad_top (NULL for the last stage) is d loss / d boundary_out as written by the next stage's backward;
 *             grad_out is the NCHW image gradient (entry_begin == 0) or the NHWC gradient w.r.t. `input`, again possibly
 *             peer memory (the previous stage's grad_top).
 * Ordering between the stages' streams is the caller's job (events), exactly like the reference's stream-ordered
 * `.to(device)`. */
MAUA_API int maua_plan_forward_stage(maua_plan_t* plan, const float* input, int h, int w, const maua_tap_io* taps,
                                     const maua_image_io* image_io, float* losses_out, int keep_for_backward,
                                     float* boundary_out, maua_stream_t stream);
MAUA_API int maua_plan_backward_stage(maua_plan_t* plan, const float* grad_coefs, const float* grad_top,
                                      float* grad_out, maua_stream_t stream);
/* cudaDeviceEnablePeerAccess in both directions; MAUA_STATUS_CUDA if the devices are not peers. */
MAUA_API int maua_enable_peer_access(int device_a, int device_b);
/* Extent of the stage's last activation for an input of h x w (what boundary_out must hold): NHWC [*oh][*ow][*oc]. */
MAUA_API int maua_plan_stage_output_shape(const maua_plan_t* plan, int h, int w, int* oh, int* ow, int* oc);

/* ScaleGradients (loss.py:10-20) + strength bookkeeping on the device: for module i with upstream gradient
 * up[i], coef[i] = sum over its loss terms of  normalize ? sg(up*term_scale)*strength^2 : up*term_scale,
 * sg(x) = x/(|x|+1e-8).  term scales: style {strength, vsf*strength (if vsf>0)}, content {strength},
 * TV {strength, never normalised}. */
MAUA_API int maua_loss_grad_coefs(const float* upstream, float* coefs, int n, const float* strength,
                                  const float* vsf, const int* normalize, const int* kind,
                                  maua_stream_t stream);

/* Copy the Gram matrix of style tap `tap` from the last forward (normalised, [C][C]) into dst (may be NULL to
 * query *c only). */
MAUA_API int maua_plan_tap_gram(maua_plan_t* plan, int tap, float* dst, int* c, maua_stream_t stream);
/* Copy the feature map of tap `tap` from the last forward (NHWC [H_l][W_l][C]) into dst (NULL: query the shape). */
MAUA_API int maua_plan_tap_feature(maua_plan_t* plan, int tap, float* dst, int* h, int* w, int* c,
                                   maua_stream_t stream);
/* The same copy into a wider matrix: pixel p of the tap goes to dst + p * dst_pixel_stride (channels contiguous).  With
 * dst = X + b*C and stride B*C the B frames of an img_vid window are laid side by side as the [H_l*W_l][B*C] matrix whose
 * Gram is the reference's dynamic [B*C, B*C] Gram (loss.py:164-168) and whose diagonal blocks are the per-frame static
 * Grams (loss.py:143-144).  Asynchronous on `stream`, graph-capturable. */
MAUA_API int maua_plan_tap_feature_strided(maua_plan_t* plan, int tap, float* dst, long dst_pixel_stride,
                                           maua_stream_t stream);
/* Backward term of a style tap in MAUA_MODE_EXTERNAL: the gradient w.r.t. the tapped feature map gets
 * + sum_k in2[pixel][k] * w2[c][k] (+ bias[c]), folded into the dgrad GEMM of the layer above as k2/32 extra k-steps.
 * in2: NHWC [H_l][W_l][k2] (TF32-rounded values), w2: [C][k2] (TF32-rounded), bias: [C] or NULL; k2 % 32 == 0.  Replaces
 * the four backward torch.mm of loss.py:141-181 for B > 1.  Must be called after every forward that used the mode. */
MAUA_API int maua_plan_set_tap_fold(maua_plan_t* plan, int tap, const float* in2, int k2, const float* w2,
                                    const float* bias);
/* Copy the output of stack entry `entry` (conv entries: the post-ReLU activation, models.py:129-130; pool entries: the
 * pooled map) from the last forward, NHWC [*h][*w][*c]; dst NULL queries the shape.  These are the tensors the backward
 * pass derives its ReLU masks and max-pool arg-max decisions from (tests: decision-level comparison with the oracle). */
MAUA_API int maua_plan_entry_output(maua_plan_t* plan, int entry, float* dst, int* h, int* w, int* c, int* is_pool,
                                    maua_stream_t stream);
/* Switch the GEMM-shaped kernels of this plan between MAUA_IMPL_TC (default), MAUA_IMPL_REF / _1CTA / _2CTA (tests) and
 * MAUA_IMPL_FP32 (exact arithmetic; allocates un-rounded GEMM-layout weight copies on first use). */
MAUA_API int maua_plan_set_impl(maua_plan_t* plan, int impl);
/* Fused pooling: when enabled, MaxPool2d / AvgPool2d(2,2) (models.py:119-122) is computed in the epilogue of the
 * convolution that produces its input instead of a separate pass over the activation (same arithmetic, bit for bit).
 * Off by default in this release (also switched on by MAUA_FUSE_POOL=1 in the environment at plan creation). */
MAUA_API int maua_plan_set_fuse_pool(maua_plan_t* plan, int enable);
/* Loss modules of the forward pass on a side stream: the GramMatrix / StyleLoss (loss.py:67-91, 141-157) and ContentLoss MSE
 * (loss.py:42-64) kernels of a tap only read that tap's feature map, so they can run next to the following convolutions
 * (models.py:403-431 splices them between the layers; autograd orders them the same way).  mode -1 = automatic (on up to
 * 640 x 640 pixels, where conv launches leave SMs idle), 0 = off, 1 = on; MAUA_SIDE_STREAM in the environment at plan creation.
 * Forked with an event after the producing conv and joined once at the end of maua_plan_forward: works eagerly and inside
 * stream capture; results are bit-identical (same kernels, same order per tap). */
MAUA_API int maua_plan_set_side_stream(maua_plan_t* plan, int mode);
/* K-split of the last partial wave of conv tiles (see maua_conv_tile_plan).  Off by default: on B200 the hand-over of
 * the partial accumulators costs more than the idle SMs it recovers (profiles/r02_splitk_ab.txt); results agree with the
 * unsplit plan to fp32 summation order. */
MAUA_API int maua_plan_set_splitk(maua_plan_t* plan, int enable);
/* Tail handling of the persistent conv kernels: 0 whole tiles, 1 K-split (parts meet in the last part's CTA through flags), 2
 * half-N items, 3 = 2 plus: a launch with FEWER tiles than SMs (the deep layers at <= 512^2) is K-split over all SMs, every part
 * dumps its raw accumulators and a small reduce kernel behind it sums them and runs the fused epilogue (MAUA_CONV_TAIL in the
 * environment at plan creation).  Half-N items compute exactly the sums of the whole tile, so results do not change. */
MAUA_API int maua_plan_set_conv_tail(maua_plan_t* plan, int mode);
/* Per-launch timing for roofline reports: when enabled, a CUDA event is recorded on the caller's stream after every
 * launch of forward / backward.  maua_plan_profile_json synchronises the stream and writes a JSON array
 * [{"name","layer","ms","flops","bytes"}...] (algorithmic FLOPs / bytes per launch) for the last forward + backward
 * into buf; returns the number of bytes needed (> cap: nothing written) or -1. */
MAUA_API int maua_plan_set_profile(maua_plan_t* plan, int enable);
MAUA_API long maua_plan_profile_json(maua_plan_t* plan, char* buf, long cap, maua_stream_t stream);
/* Number of kernel launches issued by the last forward / backward (for bench.py's gpu_launches). */
MAUA_API int maua_plan_last_launches(const maua_plan_t* plan, int* forward, int* backward);

#ifdef __cplusplus
}
#endif
#endif /* MAUA_B200_H */
