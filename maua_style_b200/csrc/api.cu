// extern "C" surface of libmaua_b200.so for the per-kernel entry points (include/maua_b200.h).
// The plan-level entry points live in plan.cu.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "conv_tc.cuh"
#include "gram.cuh"
#include "maua_b200.h"
#include "pointwise.cuh"

namespace maua {

static thread_local char g_last_error[1024] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int check_device_arch(int device) {
    // two attribute queries, cached per device: cudaGetDeviceProperties costs 60-270 ms per call on a busy B200 (measured:
    // profiles/r03_multires_trace.txt, "load_model" of a warm job) and this check runs for every network that is built
    static int ok_major[64] = {0};
    if (device >= 0 && device < 64 && ok_major[device] == 10) return MAUA_OK;
    int major = 0, minor = 0;
    cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_last_error("cudaDeviceGetAttribute(%d) failed: %s (no CUDA device? this library has no CPU fallback)",
                       device, cudaGetErrorString(e));
        return MAUA_ERR_CUDA;
    }
    if (major != 10) {
        set_last_error("device %d is sm_%d%d; libmaua_b200 is built for sm_100a (B200) only", device, major, minor);
        return MAUA_ERR_ARCH;
    }
    if (device >= 0 && device < 64) ok_major[device] = major;
    return MAUA_OK;
}

static int current_device_ok() {
    static thread_local int checked_dev = -1;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_last_error("cudaGetDevice failed: %s (no CUDA device? this library has no CPU fallback)",
                       cudaGetErrorString(e));
        return MAUA_ERR_CUDA;
    }
    if (dev == checked_dev) return MAUA_OK;
    int rc = check_device_arch(dev);
    if (rc == MAUA_OK) checked_dev = dev;
    return rc;
}

bool pdl_enabled(int kind) {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MAUA_PDL"); v = e ? atoi(e) : (PDL_CONV | PDL_GRAM | PDL_EDGE); }
    return (v & kind) != 0;
}

ReduceScratch scratch_from_workspace(void* ws) {
    ReduceScratch rs;
    rs.counter = reinterpret_cast<unsigned int*>(ws);
    rs.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 256);
    rs.max_blocks = 148 * 8;
    return rs;
}

}  // namespace maua

using namespace maua;

#define MAUA_ENTRY_GUARD()                      \
    do {                                        \
        int _rc = current_device_ok();          \
        if (_rc != MAUA_OK) return _rc;         \
    } while (0)

extern "C" {

MAUA_API int maua_abi_version(void) { return 1; }
MAUA_API const char* maua_last_error(void) { return g_last_error; }
MAUA_API int maua_device_check(int device) { return check_device_arch(device); }

MAUA_API int maua_prep_conv_weights(const float* w, float* out, int cout, int cin, int dgrad, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(w && out && cout > 0 && cin > 0, "maua_prep_conv_weights: bad arguments");
    return prep_weights_launch(w, out, cout, cin, dgrad, 1, (cudaStream_t)stream);
}
MAUA_API int maua_prep_conv_weights_ex(const float* w, float* out, int cout, int cin, int dgrad, int round_tf32,
                                       maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(w && out && cout > 0 && cin > 0, "maua_prep_conv_weights_ex: bad arguments");
    return prep_weights_launch(w, out, cout, cin, dgrad, round_tf32, (cudaStream_t)stream);
}
MAUA_API int maua_nchw_to_nhwc(const float* src, float* dst, int b, int c, int h, int w, int round_tf32,
                               maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(src && dst && b > 0 && c > 0 && h > 0 && w > 0, "maua_nchw_to_nhwc: bad arguments");
    return nchw_to_nhwc_launch(src, dst, b, c, h, w, round_tf32, (cudaStream_t)stream);
}
MAUA_API int maua_nhwc_to_nchw(const float* src, float* dst, int b, int c, int h, int w, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(src && dst && b > 0 && c > 0 && h > 0 && w > 0, "maua_nhwc_to_nchw: bad arguments");
    return nhwc_to_nchw_launch(src, dst, b, c, h, w, (cudaStream_t)stream);
}

MAUA_API int maua_conv3x3_fwd(const float* x, const float* wg, const float* bias, float* y, int b, int h, int w,
                              int cin, int cout, int relu, int impl, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(x && wg && y, "maua_conv3x3_fwd: null pointer");
    ConvArgs a;
    a.B = b; a.H = h; a.W = w; a.Cin = cin; a.Cout = cout; a.ntaps = 9;
    a.in = x; a.wg = wg;
    a.ep.out = y; a.ep.bias = bias; a.ep.relu = relu; a.ep.round = impl == MAUA_IMPL_FP32 ? 0 : 1;
    return conv_dispatch(a, impl, (cudaStream_t)stream);
}

MAUA_API int maua_prep_conv_weights_k(const float* w, float* out, int cout, int cin, int ks, int dgrad, int round_tf32,
                                      maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(w && out && cout > 0 && cin > 0 && (ks == 1 || ks == 3 || ks == 5), "maua_prep_conv_weights_k: bad arguments");
    return prep_weights_launch(w, out, cout, cin, dgrad, round_tf32, (cudaStream_t)stream, ks * ks);
}
MAUA_API int maua_conv_kxk_fwd(const float* x, const float* wg, const float* bias, float* y, int b, int h, int w, int cin,
                               int cout, int ks, int relu, int impl, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(x && wg && y && (ks == 1 || ks == 3 || ks == 5), "maua_conv_kxk_fwd: bad arguments");
    MAUA_REQUIRE(impl != MAUA_IMPL_FP32 || ks != 5, "maua_conv_kxk_fwd: the exact-arithmetic 5x5 convolution is maua_conv_direct_fwd");
    ConvArgs a;
    a.B = b; a.H = h; a.W = w; a.Cin = cin; a.Cout = cout; a.ntaps = ks * ks;
    a.in = x; a.wg = wg;
    a.ep.out = y; a.ep.bias = bias; a.ep.relu = relu; a.ep.round = impl == MAUA_IMPL_FP32 ? 0 : 1;
    return conv_dispatch(a, impl, (cudaStream_t)stream);
}

MAUA_API int maua_conv3x3_dgrad(const float* gy, const float* wd, float* gx, int b, int h, int w, int cout, int cin,
                                const float* mask_src, const float* aux_f, const float* aux_d,
                                const float* aux_bias, const float* cont_f, const float* cont_t,
                                const float* cont_coef, int round_tf32, int impl, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(gx, "maua_conv3x3_dgrad: null output");
    MAUA_REQUIRE((gy != nullptr) == (wd != nullptr), "maua_conv3x3_dgrad: gy and wd must both be given or both NULL");
    MAUA_REQUIRE((aux_f != nullptr) == (aux_d != nullptr), "maua_conv3x3_dgrad: aux_f and aux_d go together");
    MAUA_REQUIRE(!cont_f || (cont_t && cont_coef), "maua_conv3x3_dgrad: cont_f needs cont_t and cont_coef");
    ConvArgs a;
    a.B = b; a.H = h; a.W = w;
    a.Cin = cout;  // GEMM K per tap = channels of the incoming gradient
    a.Cout = cin;  // GEMM N = channels of the produced gradient
    a.ntaps = gy ? 9 : 0;
    a.in = gy; a.wg = wd;
    if (aux_f) { a.K2 = cin; a.in2 = aux_f; a.w2 = aux_d; }
    a.ep.out = gx; a.ep.bias = aux_bias; a.ep.mask_src = mask_src;
    a.ep.cont_f = cont_f; a.ep.cont_t = cont_t; a.ep.cont_coef = cont_coef;
    a.ep.relu = 0; a.ep.round = round_tf32;
    if (!gy && !aux_f) {
        MAUA_REQUIRE(false, "maua_conv3x3_dgrad: nothing to compute (no gy and no aux term)");
    }
    return conv_dispatch(a, impl, (cudaStream_t)stream);
}

MAUA_API int maua_conv_tile_plan(int h, int w, int cin, int cout, int ntaps, int k2, int sms, int tail_mode, int* plan6) {
    MAUA_REQUIRE(plan6 && h > 0 && w > 0 && cout > 0 && sms > 0, "maua_conv_tile_plan: bad arguments");
    ConvArgs a;
    a.B = 1; a.H = h; a.W = w; a.Cin = cin; a.Cout = cout; a.ntaps = ntaps; a.K2 = k2;
    // K-split plans assume the workspace exists (the plan owns one)
    static float dummy_ws; static unsigned int dummy_flag;
    if (tail_mode == 1 || tail_mode == 3) { a.splitk_ws = &dummy_ws; a.splitk_flags = &dummy_flag; }
    a.tail_mode = tail_mode;
    conv_tile_plan(a, sms, tail_mode >= 1 && tail_mode <= 3 ? tail_mode : 0, plan6, plan6 + 1, plan6 + 2, plan6 + 3, plan6 + 4, plan6 + 5);
    return MAUA_OK;
}

MAUA_API int maua_relu_mask_bits(const float* x, uint32_t* bits, long npix, int c, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(x && bits && npix > 0 && c > 0, "maua_relu_mask_bits: bad arguments");
    return relu_mask_bits_launch(x, bits, npix, c, (cudaStream_t)stream);
}

MAUA_API int maua_conv3x3_dgrad_bits(const float* gy, const float* wd, float* gx, int b, int h, int w, int cout, int cin,
                                     const uint32_t* mask_bits, const float* aux_f, const float* aux_d,
                                     const float* aux_bias, int round_tf32, int impl, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(gx, "maua_conv3x3_dgrad_bits: null output");
    MAUA_REQUIRE((gy != nullptr) == (wd != nullptr), "maua_conv3x3_dgrad_bits: gy and wd must both be given or both NULL");
    MAUA_REQUIRE((aux_f != nullptr) == (aux_d != nullptr), "maua_conv3x3_dgrad_bits: aux_f and aux_d go together");
    MAUA_REQUIRE(gy || aux_f, "maua_conv3x3_dgrad_bits: nothing to compute (no gy and no aux term)");
    ConvArgs a;
    a.B = b; a.H = h; a.W = w;
    a.Cin = cout; a.Cout = cin;
    a.ntaps = gy ? 9 : 0;
    a.in = gy; a.wg = wd;
    if (aux_f) { a.K2 = cin; a.in2 = aux_f; a.w2 = aux_d; }
    a.ep.out = gx; a.ep.bias = aux_bias; a.ep.mask_bits = mask_bits;
    a.ep.relu = 0; a.ep.round = round_tf32;
    return conv_dispatch(a, impl, (cudaStream_t)stream);
}

MAUA_API int maua_conv_first_fwd(const float* img, const float* w_oihw, const float* bias, float* y, int b, int h,
                                 int w, int cout, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(img && w_oihw && y, "maua_conv_first_fwd: null pointer");
    return conv_first_fwd_launch(img, w_oihw, bias, y, nullptr, b, h, w, cout, 1, (cudaStream_t)stream);
}
MAUA_API int maua_conv_first_dgrad(const float* gy, const float* w_oihw, float* gimg, int b, int h, int w, int cout,
                                   const float* img, const float* tv_coef, const float* temp_target,
                                   const float* temp_weights, const float* temp_coef, void* workspace,
                                   maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(gy && w_oihw && gimg, "maua_conv_first_dgrad: null pointer");
    MAUA_REQUIRE(!(tv_coef || temp_coef) || img, "maua_conv_first_dgrad: TV / temporal terms need the image");
    MAUA_REQUIRE(workspace, "maua_conv_first_dgrad: workspace (maua_conv_first_dgrad_workspace_bytes) is required");
    ImageTail t;
    t.img = img; t.tv_coef = tv_coef; t.temp_target = temp_target; t.temp_weights = temp_weights;
    t.temp_coef = temp_target ? temp_coef : nullptr;
    float* T = reinterpret_cast<float*>(workspace);
    float* wt = T + (((size_t)b * h * w * 32 + 63) & ~size_t(63));
    int rc = conv_first_dgrad_prep_weights(w_oihw, wt, cout, 1, (cudaStream_t)stream);
    if (rc) return rc;
    return conv_first_dgrad_launch(gy, wt, gimg, b, h, w, cout, t, T, MAUA_IMPL_TC, (cudaStream_t)stream);
}

MAUA_API size_t maua_conv_first_dgrad_workspace_bytes(int b, int h, int w) {
    return conv_first_dgrad_workspace_bytes(b, h, w) + 256;
}

MAUA_API int maua_pool2x2_fwd(const float* x, float* y, int b, int h, int w, int c, int avg, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(x && y, "maua_pool2x2_fwd: null pointer");
    return pool_fwd_launch(x, y, b, h, w, c, avg, 1, (cudaStream_t)stream);
}
MAUA_API int maua_pool2x2_bwd(const float* x, const float* gy, const float* addend, float* gx, int b, int h, int w,
                              int c, int avg, int round_tf32, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(x && gy && gx, "maua_pool2x2_bwd: null pointer");
    return pool_bwd_launch(x, gy, addend, gx, b, h, w, c, avg, round_tf32, (cudaStream_t)stream);
}

MAUA_API int maua_conv_direct_fwd(const float* in, int in_nchw, const float* w, const float* bias, float* out, int b, int h,
                                  int wd, int cin, int cout, int ks, int stride, int pad, int relu, int round_tf32,
                                  maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    return conv_gen_fwd_launch(in, in_nchw, w, bias, out, nullptr, b, h, wd, cin, cout, ks, stride, pad, relu, round_tf32,
                               (cudaStream_t)stream);
}
MAUA_API int maua_conv_direct_flip_weights(const float* w, float* out, int cout, int cin, int ks, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(w && out, "maua_conv_direct_flip_weights: null pointer");
    return conv_gen_flip_weights_launch(w, out, cout, cin, ks, (cudaStream_t)stream);
}
MAUA_API int maua_conv_direct_dgrad_image(const float* gout, const float* w, float* gimg, int b, int h, int wd, int cout, int ks,
                                          int stride, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    ImageTail tail;
    return conv_gen_dgrad_img_launch(gout, w, gimg, b, h, wd, cout, ks, stride, tail, (cudaStream_t)stream);
}
MAUA_API int maua_pool3x3_fwd(const float* x, float* y, int b, int h, int w, int c, int avg, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(x && y, "maua_pool3x3_fwd: null pointer");
    return pool3_fwd_launch(x, y, b, h, w, c, avg, 0, (cudaStream_t)stream);
}
MAUA_API int maua_pool3x3_bwd(const float* x, const float* gy, const float* addend, float* gx, int b, int h, int w, int c,
                              int avg, int round_tf32, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(x && gy && gx, "maua_pool3x3_bwd: null pointer");
    return pool3_bwd_launch(x, gy, addend, gx, b, h, w, c, avg, round_tf32, (cudaStream_t)stream);
}

MAUA_API size_t maua_reduce_workspace_bytes(void) { return 256 + sizeof(double) * 148 * 8 * 4; }

MAUA_API size_t maua_gram_workspace_bytes(int c) { return gram_workspace_bytes(c); }

MAUA_API int maua_gram(const float* f, long p, int c, int use_covariance, float* gram, float* mean_out,
                       void* workspace, int impl, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(f && gram && workspace, "maua_gram: null pointer");
    MAUA_REQUIRE(!use_covariance || mean_out, "maua_gram: covariance needs mean_out");
    return gram_launch(f, p, c, use_covariance, gram, mean_out, workspace, impl, (cudaStream_t)stream);
}
MAUA_API int maua_style_loss_fwd(const float* gram, const float* target, int c, float value_scale, float* loss_out,
                                 float* diff, void* workspace, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(gram && target && loss_out && diff && workspace, "maua_style_loss_fwd: null pointer");
    return style_loss_fwd_launch(gram, target, c, value_scale, loss_out, diff, scratch_from_workspace(workspace),
                                 (cudaStream_t)stream);
}
MAUA_API int maua_style_loss_bwd_prep(const float* diff, const float* mean, int c, long p, const float* coef,
                                      float* aux_d, float* aux_bias, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(diff && coef && aux_d, "maua_style_loss_bwd_prep: null pointer");
    MAUA_REQUIRE((mean != nullptr) == (aux_bias != nullptr), "maua_style_loss_bwd_prep: mean and aux_bias go together");
    return style_loss_bwd_prep_launch(diff, mean, c, p, coef, aux_d, aux_bias, (cudaStream_t)stream);
}
MAUA_API int maua_content_loss_fwd(const float* x, const float* weights, const float* target, long n, long plane,
                                   float value_scale, float* loss_out, void* workspace, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(x && target && loss_out && workspace && n > 0, "maua_content_loss_fwd: bad arguments");
    const float scale = value_scale / (float)n;
    ReduceScratch rs = scratch_from_workspace(workspace);
    if (!weights && n % 4 == 0) return mse_value_launch(x, target, n, scale, loss_out, rs, (cudaStream_t)stream);
    return wmse_value_launch(x, weights, target, n, plane > 0 ? plane : n, scale, loss_out, rs, (cudaStream_t)stream);
}
MAUA_API int maua_tv_loss_fwd(const float* img, int planes, int h, int w, float strength, float* loss_out,
                              void* workspace, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(img && loss_out && workspace, "maua_tv_loss_fwd: null pointer");
    return tv_value_launch(img, planes, h, w, strength, loss_out, scratch_from_workspace(workspace),
                           (cudaStream_t)stream);
}

MAUA_API int maua_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, float lr,
                            float beta1, float beta2, float eps, int step, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(param && grad && exp_avg && exp_avg_sq && n > 0, "maua_adam_step: bad arguments");
    return adam_launch(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, nullptr, (cudaStream_t)stream);
}
MAUA_API int maua_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, float lr,
                                float beta1, float beta2, float eps, const int* step_dev, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(param && grad && exp_avg && exp_avg_sq && step_dev && n > 0, "maua_adam_step_dev: bad arguments");
    return adam_launch(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, 0, step_dev, (cudaStream_t)stream);
}

// ---- image-side operations either side of the loop (image_ops.cu) ----
MAUA_API int maua_resize_bilinear(const float* src, float* dst, int planes, int h_in, int w_in, int h_out, int w_out,
                                  double scale_h, double scale_w, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(src && dst && planes > 0 && h_in > 0 && w_in > 0 && h_out > 0 && w_out > 0,
                 "maua_resize_bilinear: bad arguments (planes=%d %dx%d -> %dx%d)", planes, h_in, w_in, h_out, w_out);
    // ATen compute_scales_value: 1 / scale_factor when the caller gave one (F.interpolate(scale_factor=...) without
    // recompute_scale_factor), else in / out
    const float rh = scale_h > 0.0 ? (float)(1.0 / scale_h) : (float)h_in / (float)h_out;
    const float rw = scale_w > 0.0 ? (float)(1.0 / scale_w) : (float)w_in / (float)w_out;
    return resize_bilinear_launch(src, dst, planes, h_in, w_in, h_out, w_out, rh, rw, (cudaStream_t)stream);
}
MAUA_API int maua_grid_sample_border(const float* src, const float* grid, float* dst, int planes, int h_in, int w_in,
                                     int h_out, int w_out, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(src && grid && dst && planes > 0 && h_in > 0 && w_in > 0 && h_out > 0 && w_out > 0,
                 "maua_grid_sample_border: bad arguments");
    MAUA_REQUIRE((reinterpret_cast<uintptr_t>(grid) & 7) == 0, "maua_grid_sample_border: grid must be 8-byte aligned");
    return grid_sample_border_launch(src, grid, dst, planes, h_in, w_in, h_out, w_out, (cudaStream_t)stream);
}
MAUA_API int maua_preprocess_u8(const uint8_t* rgb_hwc, float* bgr_chw, int h, int w, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(rgb_hwc && bgr_chw && h > 0 && w > 0, "maua_preprocess_u8: bad arguments");
    return preprocess_u8_launch(rgb_hwc, bgr_chw, (long)h * w, (cudaStream_t)stream);
}
MAUA_API int maua_preprocess_f32(const float* rgb_chw, float* bgr_chw, int h, int w, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(rgb_chw && bgr_chw && h > 0 && w > 0, "maua_preprocess_f32: bad arguments");
    return preprocess_f32_launch(rgb_chw, bgr_chw, (long)h * w, (cudaStream_t)stream);
}
MAUA_API int maua_deprocess_u8(const float* bgr_chw, uint8_t* rgb_hwc, int h, int w, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(rgb_hwc && bgr_chw && h > 0 && w > 0, "maua_deprocess_u8: bad arguments");
    return deprocess_u8_launch(bgr_chw, rgb_hwc, (long)h * w, (cudaStream_t)stream);
}
MAUA_API int maua_blend(const float* x, const float* y, float* out, long n, float a, float b, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(x && y && out && n > 0, "maua_blend: bad arguments");
    return blend_launch(x, y, out, n, a, b, (cudaStream_t)stream);
}
MAUA_API int maua_image_moments(const float* img_chw, int h, int w, double* moments, void* workspace, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(img_chw && moments && workspace && h > 0 && w > 0, "maua_image_moments: bad arguments");
    MAUA_REQUIRE((reinterpret_cast<uintptr_t>(moments) & 7) == 0, "maua_image_moments: moments must be 8-byte aligned");
    return image_moments_launch(img_chw, (long)h * w, moments, scratch_from_workspace(workspace), (cudaStream_t)stream);
}
MAUA_API int maua_hist_match_coefs(const double* target_moments, const double* source_moments, int n_sources, double eps,
                                   float* affine, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(target_moments && source_moments && affine && n_sources > 0 && eps >= 0.0,
                 "maua_hist_match_coefs: bad arguments (n_sources=%d eps=%g)", n_sources, eps);
    return hist_match_coefs_launch(target_moments, source_moments, n_sources, eps, affine, (cudaStream_t)stream);
}
MAUA_API int maua_color_affine(const float* src_chw, float* dst_chw, int h, int w, const float* affine, maua_stream_t stream) {
    MAUA_ENTRY_GUARD();
    MAUA_REQUIRE(src_chw && dst_chw && affine && h > 0 && w > 0, "maua_color_affine: bad arguments");
    return color_affine_launch(src_chw, dst_chw, (long)h * w, affine, (cudaStream_t)stream);
}

}  // extern "C"
