// Image-side kernels either side of the optimisation loop: the scale transition
// of the multi-resolution driver, the video warp, and the pre/post-processing of load.py -- so the pastiche stays
// resident in HBM between scales / frames and only 8-bit RGB crosses PCIe.
//
//   resize_bilinear   F.interpolate(x, scale_factor=s | size=(h,w), mode="bilinear", align_corners=False)
//                     reference style.py:38-41, :47-49, :57-66, :205-212, :241-243, :253-255, :284-286, load.py:211-213
//   grid_sample       F.grid_sample(x, grid, padding_mode="border")  (bilinear, align_corners=False)
//                     reference style.py:223, :279
//   preprocess        load.py:21-32  (ToTensor()*255 -> RGB->BGR -> subtract the BGR mean)
//   deprocess         load.py:47-52  (add the mean -> BGR->RGB -> /255 -> clamp -> ToPILImage's mul(255).byte())
//   blend             style.py:290   pastiche = (1 - temporal_blend) * blend_image + temporal_blend * pastiche
//
// All are memory-bound, one thread per output pixel (x fastest => coalesced stores, gathers with 2-D locality), grid
// capped at a multiple of the SM count with a grid-stride loop.  The arithmetic follows ATen's published kernels
// (UpSampleBilinear2d / GridSampler) operation by operation in fp32 with explicit rounding intrinsics; the fused
// multiply-adds sit where the ATen CPU build has them, which makes the results bit-identical to torch's CPU ops.
#include "pointwise.cuh"

namespace maua {

namespace {

constexpr int kThreads = 256;
inline int grid_for(long n_items, int per_sm = 8) {
    long blocks = (n_items + kThreads - 1) / kThreads;
    long cap = 148L * per_sm;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// ATen area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false)
__device__ __forceinline__ float src_index(float scale, int dst) {
    const float s = __fmaf_rn(scale, (float)dst + 0.5f, -0.5f);
    return s < 0.f ? 0.f : s;
}

__global__ void __launch_bounds__(kThreads)
resize_bilinear_kernel(const float* __restrict__ src, float* __restrict__ dst, int planes, int Hin, int Win, int Hout,
                       int Wout, float rh, float rw) {
    const long npix = (long)Hout * Wout;
    const bool small = Hout + Wout <= 128;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < npix; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % Wout), y = (int)(i / Wout);
        const float hr = src_index(rh, y), wr = src_index(rw, x);
        // guard_index_and_lambda (CPU) / the CUDA kernel's h1, h1p, lambda: identical for in-range indices
        int h1 = (int)hr, w1 = (int)wr;
        h1 = h1 < Hin - 1 ? h1 : Hin - 1;
        w1 = w1 < Win - 1 ? w1 : Win - 1;
        const int h1p = h1 < Hin - 1 ? 1 : 0, w1p = w1 < Win - 1 ? 1 : 0;
        float hl1 = hr - (float)h1, wl1 = wr - (float)w1;
        hl1 = fminf(fmaxf(hl1, 0.f), 1.f);
        wl1 = fminf(fmaxf(wl1, 0.f), 1.f);
        const float hl0 = 1.f - hl1, wl0 = 1.f - wl1;
        const long o00 = (long)h1 * Win + w1, o01 = o00 + w1p, o10 = o00 + (long)h1p * Win, o11 = o10 + w1p;
        if (small) {
            // ATen sends small outputs (out_h + out_w <= 128, _use_vectorized_kernel_cond_2d) to a kernel that multiplies
            // the four weights out first; same fused-operation placement as that build
            const float w00 = __fmul_rn(hl0, wl0), w01 = __fmul_rn(hl0, wl1), w10 = __fmul_rn(hl1, wl0), w11 = __fmul_rn(hl1, wl1);
            for (int p = 0; p < planes; ++p) {
                const float* s = src + (long)p * Hin * Win;
                float acc = __fmaf_rn(__ldg(s + o00), w00, __fmul_rn(__ldg(s + o01), w01));
                acc = __fmaf_rn(__ldg(s + o10), w10, acc);
                dst[(long)p * npix + i] = __fmaf_rn(__ldg(s + o11), w11, acc);
            }
            continue;
        }
        for (int p = 0; p < planes; ++p) {
            const float* s = src + (long)p * Hin * Win;
            // fma(v0, w0, round(v1 * w1)): where the ATen CPU build fuses w0*v0 + w1*v1 (bit-exact against it)
            const float top = __fmaf_rn(__ldg(s + o00), wl0, __fmul_rn(__ldg(s + o01), wl1));
            const float bot = __fmaf_rn(__ldg(s + o10), wl0, __fmul_rn(__ldg(s + o11), wl1));
            dst[(long)p * npix + i] = __fmaf_rn(top, hl0, __fmul_rn(bot, hl1));
        }
    }
}

// ATen grid_sampler_2d, interpolation bilinear, padding border, align_corners false.  grid is [Hout][Wout][2] (x, y in [-1,1]).
__global__ void __launch_bounds__(kThreads)
grid_sample_border_kernel(const float* __restrict__ src, const float* __restrict__ grid, float* __restrict__ dst,
                          int planes, int Hin, int Win, int Hout, int Wout) {
    const long npix = (long)Hout * Wout;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < npix; i += (long)gridDim.x * blockDim.x) {
        const float2 g = __ldg(reinterpret_cast<const float2*>(grid) + i);
        // unnormalize: (coord + 1) * (size / 2) - 0.5 as one fma (ATen CPU GridSampler) ; border: clip to [0, size - 1]
        float ix = __fmaf_rn(__fadd_rn(g.x, 1.f), (float)Win * 0.5f, -0.5f);
        float iy = __fmaf_rn(__fadd_rn(g.y, 1.f), (float)Hin * 0.5f, -0.5f);
        ix = fminf((float)(Win - 1), fmaxf(ix, 0.f));
        iy = fminf((float)(Hin - 1), fmaxf(iy, 0.f));
        const float fx = floorf(ix), fy = floorf(iy);
        const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
        const float wx = ix - fx, ex = 1.f - wx, ny = iy - fy, sy = 1.f - ny;
        const float nw = __fmul_rn(sy, ex), ne = __fmul_rn(sy, wx), sw = __fmul_rn(ny, ex), se = __fmul_rn(ny, wx);
        const bool bx1 = x1 < Win, by1 = y1 < Hin;  // x0, y0 are in bounds after the clip
        for (int p = 0; p < planes; ++p) {
            const float* s = src + (long)p * Hin * Win;
            float v = __fmul_rn(__ldg(s + (long)y0 * Win + x0), nw);
            v = __fmaf_rn(bx1 ? __ldg(s + (long)y0 * Win + x1) : 0.f, ne, v);
            v = __fmaf_rn(by1 ? __ldg(s + (long)y1 * Win + x0) : 0.f, sw, v);
            v = __fmaf_rn(bx1 && by1 ? __ldg(s + (long)y1 * Win + x1) : 0.f, se, v);
            dst[(long)p * npix + i] = v;
        }
    }
}

__constant__ float kMeanBGR[3] = {103.939f, 116.779f, 123.68f};

// u8 RGB HWC (a PIL image's bytes) -> fp32 BGR CHW, 0-255, mean-subtracted.  ToTensor divides by 255 and the
// reference multiplies by 255 again (load.py:32): kept, so the result is bit-identical to the reference's.
__global__ void __launch_bounds__(kThreads)
preprocess_u8_kernel(const uint8_t* __restrict__ rgb, float* __restrict__ out, long npix) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < npix; i += (long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {  // c = output (BGR) channel; source channel 2 - c
            const float v = __fmul_rn(__fdiv_rn((float)rgb[i * 3 + (2 - c)], 255.f), 255.f);
            out[(long)c * npix + i] = __fsub_rn(v, kMeanBGR[c]);
        }
    }
}

// fp32 RGB CHW in [0,1] (ToTensor output, or load.py:22-25's "random" image) -> BGR CHW 0-255 mean-subtracted.
__global__ void __launch_bounds__(kThreads)
preprocess_f32_kernel(const float* __restrict__ rgb, float* __restrict__ out, long npix) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < npix; i += (long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            out[(long)c * npix + i] = __fsub_rn(__fmul_rn(rgb[(long)(2 - c) * npix + i], 255.f), kMeanBGR[c]);
    }
}

// fp32 BGR CHW mean-subtracted -> u8 RGB HWC: (x + mean) / 255, clamp to [0,1], ToPILImage = mul(255) and a truncating
// byte cast (load.py:47-52).
__global__ void __launch_bounds__(kThreads)
deprocess_u8_kernel(const float* __restrict__ bgr, uint8_t* __restrict__ rgb, long npix) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < npix; i += (long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {  // c = source (BGR) channel; destination channel 2 - c
            float v = __fdiv_rn(__fsub_rn(bgr[(long)c * npix + i], -kMeanBGR[c]), 255.f);
            v = fminf(fmaxf(v, 0.f), 1.f);
            rgb[i * 3 + (2 - c)] = (uint8_t)(int)__fmul_rn(v, 255.f);
        }
    }
}

// out = a * x + b * y, each product rounded (torch evaluates the two scalar multiplies and the add separately)
__global__ void __launch_bounds__(kThreads)
blend_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, long n, float a, float b) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = __fadd_rn(__fmul_rn(a, x[i]), __fmul_rn(b, y[i]));
}

}  // namespace

int resize_bilinear_launch(const float* src, float* dst, int planes, int Hin, int Win, int Hout, int Wout, float rh,
                           float rw, cudaStream_t st) {
    resize_bilinear_kernel<<<grid_for((long)Hout * Wout), kThreads, 0, st>>>(src, dst, planes, Hin, Win, Hout, Wout, rh, rw);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}
int grid_sample_border_launch(const float* src, const float* grid, float* dst, int planes, int Hin, int Win, int Hout,
                              int Wout, cudaStream_t st) {
    grid_sample_border_kernel<<<grid_for((long)Hout * Wout), kThreads, 0, st>>>(src, grid, dst, planes, Hin, Win, Hout, Wout);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}
int preprocess_u8_launch(const uint8_t* rgb, float* out, long npix, cudaStream_t st) {
    preprocess_u8_kernel<<<grid_for(npix), kThreads, 0, st>>>(rgb, out, npix);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}
int preprocess_f32_launch(const float* rgb, float* out, long npix, cudaStream_t st) {
    preprocess_f32_kernel<<<grid_for(npix), kThreads, 0, st>>>(rgb, out, npix);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}
int deprocess_u8_launch(const float* bgr, uint8_t* rgb, long npix, cudaStream_t st) {
    deprocess_u8_kernel<<<grid_for(npix), kThreads, 0, st>>>(bgr, rgb, npix);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}
int blend_launch(const float* x, const float* y, float* out, long n, float a, float b, cudaStream_t st) {
    blend_kernel<<<grid_for(n), kThreads, 0, st>>>(x, y, out, n, a, b);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

}  // namespace maua
