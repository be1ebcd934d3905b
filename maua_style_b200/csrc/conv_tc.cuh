// Host-visible interface of the tcgen05 implicit-GEMM convolution (conv_tc.cu) and of the
// direct edge kernels (conv_edge.cu).  Internal to libmaua_b200.so; the C ABI is in
// include/maua_b200.h.
#pragma once
#include "common.cuh"

namespace maua {

// Fused epilogue applied to every accumulator element v = acc[pixel][n], in this order:
//   v += bias[n]                         (forward bias; covariance rank-1 correction in backward)
//   v += *cont_coef * (cont_f - cont_t)  (ContentLoss gradient, loss.py:53-59)
//   v += addend                          (gradient arriving from another branch, e.g. un-pooled)
//   v  = max(v, 0)                       (ReLU, models.py:130)                       if relu
//   v  = mask_src > 0 ? v : 0            (ReLU backward through the *previous* layer) if mask_src / mask_bits
//   v  = round_tf32(v)                   (operand rounding for the consuming MMA)     if round
// optionally followed by the 2x2 / stride-2 pooling of the finished tile (models.py:119-122; pool_out, opt-in), and,
// for forward layers, mask_out receives the sign bitmap (v > 0) of the result: 1 bit per element, word
// [pixel * (Cout/32) + n/32], bit n % 32 -- what the dgrad of the layer above reads as mask_bits instead of re-reading
// the fp32 activation (32x less mask traffic in the backward pass).
struct ConvEpilogue {
    float* out = nullptr;             // NHWC [B][H][W][Cout]
    float* out2 = nullptr;            // optional second copy of the output (layer-wise split: the next stage's input,
                                      // usually peer memory reached over NVLink -- stored tile by tile from the epilogue)
    const float* bias = nullptr;      // [Cout]
    const float* mask_src = nullptr;  // NHWC like out (fp32 activation; slower than mask_bits)
    const uint32_t* mask_bits = nullptr;  // sign bitmap of the activation being masked (see above)
    uint32_t* mask_out = nullptr;     // sign bitmap of this launch's output
    const float* cont_f = nullptr;    // NHWC like out
    const float* cont_t = nullptr;    // NHWC like out
    const float* cont_coef = nullptr; // device scalar
    const float* addend = nullptr;    // NHWC like out
    int relu = 0;
    int round = 1;
    // Fused 2x2 pooling (tcgen05 path with the TMA-store epilogue only): the pooled map [B][H/2][W/2][Cout] is written
    // next to the full-resolution output from the same registers (max: the window's four pixels sit in lanes l, l^1,
    // l^16, l^17 of one epilogue warp).  Same arithmetic as pool_fwd_kernel, bit for bit.
    float* pool_out = nullptr;
    int pool_avg = 0;
    // Optional, max pooling: arg-max codes of the pooled map, one byte per 4 channels ([B][H/2][W/2][Cout/4], two bits per
    // channel, first maximum in row-major window order like pool_bwd_kernel recomputes it) -- lets the backward pass un-pool
    // from the codes and the sign bitmap instead of re-reading the pre-pool activation (pool_bwd_codes_kernel).
    uint8_t* pool_codes = nullptr;
};

// out[b][h][w][n] = epilogue( sum_{tap,c} in[b][h+dy][w+dx][c] * wg[n][tap*Cin + c]
//                           + sum_{c2}    in2[b][h][w][c2]     * w2[n][c2] )
// ntaps is 9 (3x3, pad 1), 1 (pointwise) or 0 (main term absent, aux GEMM only).
struct ConvArgs {
    int B = 1, H = 0, W = 0;
    int Cin = 0, Cout = 0, ntaps = 9;
    int force_cg = 0;            // 0: pick single CTAs or CTA pairs per launch; 1 / 2: force (tests)
    const float* in = nullptr;   // NHWC [B][H][W][Cin], values tf32-representable
    const float* wg = nullptr;   // [Cout][ntaps*Cin], tf32-rounded
    int K2 = 0;
    const float* in2 = nullptr;  // NHWC [B][H][W][K2]
    const float* w2 = nullptr;   // [Cout][K2]
    ConvEpilogue ep;
    // Split-K of the last (partial) wave of tiles (tcgen05 path): partial accumulators travel through this workspace.
    // Optional: without it every tile is computed by one CTA (pair).  Sizes: conv_splitk_ws_bytes() / conv_splitk_flag_words();
    // the flag words must be zero before the first launch and are left zero by every launch.
    float* splitk_ws = nullptr;
    unsigned int* splitk_flags = nullptr;
    // How the last partial wave of tiles is computed: 0 = whole tiles, 1 = K-split (needs the workspace above), 2 = as two
    // half-N items per tile (independent, no hand-over).
    int tail_mode = 0;
    // Optional: weights of the NEXT convolution of the iteration (frozen data).  The tcgen05 kernel's spare warp asks L2 to fetch
    // them (cp.async.bulk.prefetch.L2) while this launch runs, so that the next launch's first weight tiles -- which sit at the
    // head of its k-loop chain -- do not come from HBM (the optimizer's history sweep flushes L2 once per iteration).
    const void* prefetch = nullptr;
    size_t prefetch_bytes = 0;  // multiple of 16
};
void conv_tile_plan(const ConvArgs& a, int sms, int tail_mode, int* bn, int* mt, int* cg, int* full_tiles, int* split_tiles,
                    int* split);
size_t conv_splitk_ws_bytes();
size_t conv_splitk_flag_words();

// sign bitmap of an NHWC activation: bits[pixel * (C/32) + c/32] bit (c % 32) = x[pixel][c] > 0   (C % 32 == 0)
int relu_mask_bits_launch(const float* x, uint32_t* bits, long npix, int C, cudaStream_t st);

int conv_tc_launch(const ConvArgs& a, cudaStream_t st);   // tcgen05 / TMEM / TMA path
int conv_ref_launch(const ConvArgs& a, cudaStream_t st);  // naive SIMT cross-check (debug only)
int conv_fp32_launch(const ConvArgs& a, cudaStream_t st); // exact arithmetic: FP32 operands, fp64 chunk sums (conv_fp32.cu)

// impl: MAUA_IMPL_* of include/maua_b200.h (0 tcgen05, 1 naive SIMT, 2 / 3 tcgen05 with forced CTA grouping, 4 exact FP32)
inline int conv_dispatch(ConvArgs a, int impl, cudaStream_t st) {
    if (impl == 1) return conv_ref_launch(a, st);
    if (impl == 4) return conv_fp32_launch(a, st);
    a.force_cg = impl == 2 ? 1 : (impl == 3 ? 2 : 0);
    return conv_tc_launch(a, st);
}

// conv1_1 forward: NCHW 3-channel image -> NHWC Cout, bias + ReLU, fp32 FFMA (K = 27).  Optionally also the TVLoss value
// of the image (loss.py:229-233): *out = strength * (sum |x[h]-x[h-1]| + sum |x[w]-x[w-1]|), deterministic grid sum.
struct ConvFirstTV {
    float strength = 0.f;
    float* out = nullptr;
    double* partials = nullptr;      // ReduceScratch of the caller
    unsigned int* counter = nullptr;
    int max_blocks = 0;
};
int conv_first_fwd_launch(const float* img, const float* w /*[Cout][3][3][3]*/, const float* bias, float* out,
                          uint32_t* mask_out /*optional sign bitmap*/, int B, int H, int W, int Cout, int round,
                          cudaStream_t st, const ConvFirstTV* tv = nullptr, int exact = 0);
int make_tmap_nhwc(CUtensorMap* m, const float* ptr, int B, int H, int W, int C, int box_w, int box_h);

// conv1_1 dgrad (+ fused image-side tail): NHWC Cout gradient -> NCHW 3-channel image gradient,
// plus TV gradient and temporal ContentLoss gradient.
struct ImageTail {
    const float* img = nullptr;        // NCHW image (needed for TV / temporal)
    const float* tv_coef = nullptr;    // device scalar: upstream * strength (TVLoss, loss.py:224-233)
    const float* temp_target = nullptr;// NCHW warped previous frame
    const float* temp_weights = nullptr;// [H][W] reliability map or null
    const float* temp_coef = nullptr;  // device scalar: coef * 2 / numel
};
// wt: [32][Cout] from conv_first_dgrad_prep_weights; T: scratch of B*H*W*32 floats (the per-pixel tap contraction)
size_t conv_first_dgrad_workspace_bytes(int B, int H, int W);
int conv_first_dgrad_prep_weights(const float* w_oihw, float* wt, int Cout, int do_round, cudaStream_t st);
int conv_first_dgrad_launch(const float* gout, const float* wt, float* gimg, int B, int H, int W, int Cout,
                            const ImageTail& tail, float* T, int impl, cudaStream_t st);

// conv_gen.cu: direct fp32 convolutions and the 3x3 / 2 ceil_mode pooling of the NIN backbone (models.py:74-113)
// out NHWC [B][OH][OW][Cout] = act(conv(in, w [Cout][Cin][ks][ks]) + bias); `in` is NHWC, or the NCHW image when in_nchw
int conv_gen_fwd_launch(const float* in, int in_nchw, const float* w, const float* bias, float* out, uint32_t* mask_out, int B,
                        int H, int W, int Cin, int Cout, int ks, int stride, int pad, int relu, int round, cudaStream_t st);
// [Cout][Cin][ks][ks] -> [Cin][Cout][ks][ks] rotated by 180 degrees (weights of the stride-1 input-gradient convolution)
int conv_gen_flip_weights_launch(const float* w, float* out, int Cout, int Cin, int ks, cudaStream_t st);
// image-layer backward for a strided convolution without padding (+ TVLoss / temporal gradients); gout may be null
int conv_gen_dgrad_img_launch(const float* gout, const float* w, float* gimg, int B, int H, int W, int Cout, int ks, int stride,
                              const ImageTail& tail, cudaStream_t st);
// the image layer as a GEMM (product path): A [OH*OW][KP] = im2col(NCHW image), k = c*ks^2 + ky*ks + kx zero-padded to KP;
// GEMM weights wg [Cout][KP] and wt [KP][Cout] (TF32-rounded); col2im of T [OH*OW][KP] back to the NCHW image gradient + tail
int im2col_img_launch(const float* img, float* A, int H, int W, int ks, int stride, int KP, int do_round, cudaStream_t st);
int col2im_img_launch(const float* T, float* gimg, int H, int W, int ks, int stride, int KP, const ImageTail& tail, cudaStream_t st);
int im2col_weights_launch(const float* w, float* wg, float* wt, int Cout, int K, int KP, cudaStream_t st);
void pool3_out_extent(int H, int W, int* PH, int* PW);
int pool3_fwd_launch(const float* x, float* y, int B, int H, int W, int C, int avg, int do_round, cudaStream_t st);
int pool3_bwd_launch(const float* x, const float* gy, const float* addend, float* gx, int B, int H, int W, int C, int avg,
                     int do_round, cudaStream_t st);

}  // namespace maua
