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
//   match_histogram   utils.py:88-151  colour-statistics transfer (section 8f rank 3) as three launches:
//                     image_moments (one HBM pass per image: N, sum x_c, sum x_c x_d in fp64, deterministic),
//                     hist_match_coefs (one thread: 3x3 symmetric eigen-decompositions -> the affine map),
//                     color_affine (one HBM pass: y = M x + b)
//
// All are memory-bound, one thread per output pixel (x fastest => coalesced stores, gathers with 2-D locality), grid
// capped at a multiple of the SM count with a grid-stride loop.  The arithmetic follows ATen's published kernels
// (UpSampleBilinear2d / GridSampler) operation by operation in fp32 with explicit rounding intrinsics; the fused
// multiply-adds sit where the ATen CPU build has them, which makes the results bit-identical to torch's CPU ops.
#include "pointwise.cuh"
#include "reduce.cuh"

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

// ---- utils.match_histogram (utils.py:88-151) ---------------------------------------------------------------------
// moments[10] = { N, S0, S1, S2, S00, S01, S02, S11, S12, S22 } of a CHW 3-plane image, accumulated in fp64 (exact
// products of fp32 values) and reduced in a fixed order (reduce.cuh), so the result is run-to-run identical.
__global__ void __launch_bounds__(kReduceThreads)
image_moments_kernel(const float* __restrict__ img, long npix, int vec4, double* __restrict__ moments, double* partials,
                     unsigned int* counter) {
    double a[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] = 0.0;
    auto add = [&](float x0, float x1, float x2) {
        const double d0 = x0, d1 = x1, d2 = x2;
        a[0] += d0; a[1] += d1; a[2] += d2;
        a[3] = fma(d0, d0, a[3]); a[4] = fma(d0, d1, a[4]); a[5] = fma(d0, d2, a[5]);
        a[6] = fma(d1, d1, a[6]); a[7] = fma(d1, d2, a[7]); a[8] = fma(d2, d2, a[8]);
    };
    const long stride = (long)gridDim.x * blockDim.x, first = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (vec4) {
        const float4* p0 = reinterpret_cast<const float4*>(img);
        const float4* p1 = reinterpret_cast<const float4*>(img + npix);
        const float4* p2 = reinterpret_cast<const float4*>(img + 2 * npix);
        for (long i = first; i < npix / 4; i += stride) {
            const float4 u = __ldg(p0 + i), v = __ldg(p1 + i), w = __ldg(p2 + i);
            add(u.x, v.x, w.x); add(u.y, v.y, w.y); add(u.z, v.z, w.z); add(u.w, v.w, w.w);
        }
    } else {
        for (long i = first; i < npix; i += stride) add(__ldg(img + i), __ldg(img + npix + i), __ldg(img + 2 * npix + i));
    }
    double tot[9];
    if (grid_sum<9>(a, partials, counter, tot)) {
        moments[0] = (double)npix;
#pragma unroll
        for (int k = 0; k < 9; ++k) moments[1 + k] = tot[k];
    }
}

// cyclic Jacobi for a symmetric 3x3 matrix: a -> eigenvalues w, eigenvectors in the columns of v (fp64, one thread)
__device__ void sym3_eigen(const double (&c)[3][3], double (&w)[3], double (&v)[3][3]) {
    double a[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { a[i][j] = c[i][j]; v[i][j] = i == j ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 32; ++sweep) {
        const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
        if (off <= 1e-34 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = cs * akp - sn * akq;
                    a[k][q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = cs * apk - sn * aqk;
                    a[q][k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < 3; ++k) {  // V <- V J
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = cs * vkp - sn * vkq;
                    v[k][q] = sn * vkp + cs * vkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) w[i] = a[i][i];
}

// mean and covariance (+ eps on the diagonal) from the raw moments: utils.get_histogram (utils.py:88-93)
__device__ void stats_from_moments(const double* m, double eps, double (&mu)[3], double (&c)[3][3]) {
    const double n = m[0];
    for (int i = 0; i < 3; ++i) mu[i] = m[1 + i] / n;
    const double s[3][3] = {{m[4], m[5], m[6]}, {m[5], m[7], m[8]}, {m[6], m[8], m[9]}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c[i][j] = s[i][j] / n - mu[i] * mu[j] + (i == j ? eps : 0.0);
}

// f(C) = V diag(f(w)) V^T with f = sqrt (inverse = 0) or 1/sqrt (inverse = 1); negative eigenvalues count as 0
// (utils.py:125-126 `Et[Et != Et] = 0`), and are floored for the inverse
__device__ void sym3_sqrt(const double (&c)[3][3], int inverse, double (&q)[3][3]) {
    double w[3], v[3][3];
    sym3_eigen(c, w, v);
    for (int k = 0; k < 3; ++k) {
        const double r = sqrt(w[k] > 0.0 ? w[k] : 0.0);
        w[k] = inverse ? 1.0 / (r > 1e-150 ? r : 1e-150) : r;
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) q[i][j] = v[i][0] * w[0] * v[j][0] + v[i][1] * w[1] * v[j][1] + v[i][2] * w[2] * v[j][2];
}

// affine[12] = { M row-major, b }: mean over the sources of  Qs Qt^-1 (x - mu_t) + mu_s  (utils.py:112-143)
__global__ void hist_match_coefs_kernel(const double* __restrict__ target_m, const double* __restrict__ source_m,
                                        int n_sources, double eps, float* __restrict__ affine) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double mu_t[3], ct[3][3], qti[3][3], mbar[3][3] = {}, bbar[3] = {};
    stats_from_moments(target_m, eps, mu_t, ct);
    sym3_sqrt(ct, 1, qti);
    for (int s = 0; s < n_sources; ++s) {
        double mu_s[3], cs[3][3], qs[3][3];
        stats_from_moments(source_m + 10 * s, eps, mu_s, cs);
        sym3_sqrt(cs, 0, qs);
        for (int i = 0; i < 3; ++i) {
            double row[3], dot = 0.0;
            for (int j = 0; j < 3; ++j) {
                row[j] = qs[i][0] * qti[0][j] + qs[i][1] * qti[1][j] + qs[i][2] * qti[2][j];
                mbar[i][j] += row[j] / n_sources;
                dot += row[j] * mu_t[j];
            }
            bbar[i] += (mu_s[i] - dot) / n_sources;
        }
    }
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) affine[3 * i + j] = (float)mbar[i][j];
        affine[9 + i] = (float)bbar[i];
    }
}

// y_c = M[c][0] x_0 + M[c][1] x_1 + M[c][2] x_2 + b_c on CHW planes (dst may alias src)
__global__ void __launch_bounds__(kThreads)
color_affine_kernel(const float* src, float* dst, long npix, int vec4, const float* __restrict__ affine) {
    float m[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) m[k] = __ldg(affine + k);
    auto map = [&](float x0, float x1, float x2, int c) {
        return __fmaf_rn(m[3 * c + 2], x2, __fmaf_rn(m[3 * c + 1], x1, __fmaf_rn(m[3 * c], x0, m[9 + c])));
    };
    const long stride = (long)gridDim.x * blockDim.x, first = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (vec4) {
        for (long i = first; i < npix / 4; i += stride) {
            const float4 u = reinterpret_cast<const float4*>(src)[i], v = reinterpret_cast<const float4*>(src + npix)[i],
                         w = reinterpret_cast<const float4*>(src + 2 * npix)[i];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                reinterpret_cast<float4*>(dst + c * npix)[i] =
                    make_float4(map(u.x, v.x, w.x, c), map(u.y, v.y, w.y, c), map(u.z, v.z, w.z, c), map(u.w, v.w, w.w, c));
        }
    } else {
        for (long i = first; i < npix; i += stride) {
            const float x0 = src[i], x1 = src[npix + i], x2 = src[2 * npix + i];
#pragma unroll
            for (int c = 0; c < 3; ++c) dst[c * npix + i] = map(x0, x1, x2, c);
        }
    }
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

int image_moments_launch(const float* img, long npix, double* moments, ReduceScratch rs, cudaStream_t st) {
    const int vec4 = (npix % 4 == 0) && ((reinterpret_cast<uintptr_t>(img) & 15) == 0);
    long blocks = ((vec4 ? npix / 4 : npix) + kReduceThreads - 1) / kReduceThreads;
    const long cap = 148L * 3;  // 9 partial sums per block must fit the shared reduce scratch (4 per block x 148 x 8)
    const int grid = (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
    MAUA_REQUIRE((long)grid * 9 <= (long)rs.max_blocks * 4, "reduce scratch too small");
    image_moments_kernel<<<grid, kReduceThreads, 0, st>>>(img, npix, vec4, moments, rs.partials, rs.counter);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}
int hist_match_coefs_launch(const double* target_m, const double* source_m, int n_sources, double eps, float* affine,
                            cudaStream_t st) {
    hist_match_coefs_kernel<<<1, 32, 0, st>>>(target_m, source_m, n_sources, eps, affine);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}
int color_affine_launch(const float* src, float* dst, long npix, const float* affine, cudaStream_t st) {
    const int vec4 = (npix % 4 == 0) && (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0);
    color_affine_kernel<<<grid_for(vec4 ? npix / 4 : npix), kThreads, 0, st>>>(src, dst, npix, vec4, affine);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

}  // namespace maua
