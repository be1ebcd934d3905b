// Launchers of the memory-bound kernels (pointwise.cu) and the L-BFGS vector kernels (lbfgs.cu).
#pragma once
#include "common.cuh"

namespace maua {

// scratch for deterministic grid-wide sums: one double per block per value + a self-resetting counter
struct ReduceScratch {
    double* partials = nullptr;
    unsigned int* counter = nullptr;  // must be zero-initialised once
    int max_blocks = 0;
};

int nchw_to_nhwc_launch(const float* src, float* dst, int B, int C, int H, int W, int do_round, cudaStream_t st);
int nhwc_to_nchw_launch(const float* src, float* dst, int B, int C, int H, int W, cudaStream_t st);
int prep_weights_launch(const float* w, float* out, int Cout, int Cin, int dgrad, int do_round, cudaStream_t st, int taps = 9);
// do_round: TF32-round the averaged values (operands of the next tensor-core conv); max pooling never needs it
// codes (optional, max pooling): arg-max codes of the pooled map, see ConvEpilogue::pool_codes
int pool_fwd_launch(const float* x, float* y, int B, int H, int W, int C, int avg, int do_round, cudaStream_t st,
                    uint8_t* codes = nullptr);
int pool_bwd_launch(const float* x, const float* gy, const float* addend, float* gx, int B, int H, int W, int C,
                    int avg, int do_round, cudaStream_t st);
// the same un-pooling from the forward pass's arg-max codes and the sign bitmap of the pre-pool activation
// (bits[pixel * (C/32) + c/32] bit c%32 = x > 0) instead of the activation itself: 5.2 instead of 9 bytes per element
int pool_bwd_codes_launch(const uint32_t* bits, const uint8_t* codes, const float* gy, const float* addend, float* gx, int B,
                          int H, int W, int C, int avg, int do_round, cudaStream_t st);

// argmax with "first maximum in row-major window order wins" (ATen max_pool2d behaviour)
__device__ __forceinline__ int first_argmax(float a0, float a1, float a2, float a3) {
    int k = 0;
    float m = a0;
    if (a1 > m) { m = a1; k = 1; }
    if (a2 > m) { m = a2; k = 2; }
    if (a3 > m) { m = a3; k = 3; }
    return k;
}
int tv_value_launch(const float* x, int planes, int H, int W, float strength, float* loss_out, ReduceScratch rs,
                    cudaStream_t st);
int mse_value_launch(const float* x, const float* t, long n, float scale, float* loss_out, ReduceScratch rs,
                     cudaStream_t st);
int wmse_value_launch(const float* x, const float* wts, const float* t, long n, long plane, float scale,
                      float* loss_out, ReduceScratch rs, cudaStream_t st);
int channel_mean_launch(const float* x, long P, int C, float* mean_out, double* scratch, int scratch_blocks,
                        cudaStream_t st);
// image_ops.cu: scale transition, video warp, pre/post-processing
int resize_bilinear_launch(const float* src, float* dst, int planes, int Hin, int Win, int Hout, int Wout, float rh,
                           float rw, cudaStream_t st);
int grid_sample_border_launch(const float* src, const float* grid, float* dst, int planes, int Hin, int Win, int Hout,
                              int Wout, cudaStream_t st);
int preprocess_u8_launch(const uint8_t* rgb, float* out, long npix, cudaStream_t st);
int preprocess_f32_launch(const float* rgb, float* out, long npix, cudaStream_t st);
int deprocess_u8_launch(const float* bgr, uint8_t* rgb, long npix, cudaStream_t st);
int blend_launch(const float* x, const float* y, float* out, long n, float a, float b, cudaStream_t st);
// utils.match_histogram as moments -> 3x3 coefficient solve -> affine colour map
int image_moments_launch(const float* img, long npix, double* moments, ReduceScratch rs, cudaStream_t st);
int hist_match_coefs_launch(const double* target_m, const double* source_m, int n_sources, double eps, float* affine,
                            cudaStream_t st);
int color_affine_launch(const float* src, float* dst, long npix, const float* affine, cudaStream_t st);
int adam_launch(float* p, const float* g, float* m, float* v, long n, float lr, float b1, float b2, float eps,
                int step, const int* step_dev, cudaStream_t st);

}  // namespace maua
