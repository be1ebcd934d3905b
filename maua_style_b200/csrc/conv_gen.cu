// Direct convolutions and the 3x3 / stride-2 / ceil_mode pooling of the NIN backbone (reference models.py:74-113): the layer
// shapes that are not the 3x3 / pad-1 and 1x1 GEMMs of conv_tc.cu --
//     conv1  3 -> 96, 11x11, stride 4, no padding, on the NCHW image      (models.py:83)
//     conv2 96 -> 256, 5x5, stride 1, padding 2                            (models.py:90)
//     pool   MaxPool2d / AvgPool2d((3, 3), (2, 2), (0, 0), ceil_mode=True) (models.py:77-80)
// and their backward passes (input gradients only: the weights are frozen, models.py:443-445).
//
// NIN is what the stock config/scaling-img.json selects above 4096 px, where these two layers see 1/16 and 1/64 of the
// image's pixels.  They run on the CUDA cores in fp32 -- products of one staged chunk accumulate with FFMA in fp32, chunk
// sums are added in fp64, like conv_fp32.cu -- and feed the tcgen05 kernels (1x1 and 3x3 layers) TF32-rounded outputs.
// The same kernels serve the product path and the exact-arithmetic mode.
#include "conv_tc.cuh"
#include "pointwise.cuh"

namespace maua {

namespace {

constexpr int GT_P = 8;    // output tile: 8 x 8 pixels
constexpr int GT_N = 64;   // x 64 output channels
constexpr int G_IN_MAX = 4608;   // staged input patch (floats): 39 x 39 x 3 (11x11 / 4) or 12 x 12 x 8 (5x5 / 1)
constexpr int G_W_MAX = 11 * 8 * GT_N;  // one kernel row of weights for one channel chunk

struct GenConv {
    int B, H, W, Cin, Cout, ks, stride, pad, OH, OW;
    int in_nchw;         // input is the NCHW image
    int ck;              // channels per staged chunk
    const float* in;
    const float* w;      // [Cout][Cin][ks][ks]
    const float* bias;   // [Cout] or null
    float* out;          // NHWC [B][OH][OW][Cout]
    int relu, round;
};

// thread (pg, cg): 4 consecutive output pixels of one tile row x 4 output channels
__global__ void __launch_bounds__(256)
conv_gen_kernel(const GenConv a) {
    __shared__ float in_s[G_IN_MAX];
    __shared__ __align__(16) float w_s[G_W_MAX];
    const int tiles_w = (a.OW + GT_P - 1) / GT_P, tiles_h = (a.OH + GT_P - 1) / GT_P;
    const int n_tiles = (a.Cout + GT_N - 1) / GT_N;
    const long total = (long)a.B * tiles_h * tiles_w * n_tiles;
    const int cg = threadIdx.x & 15, pg = threadIdx.x >> 4;
    const int prow = pg >> 1, pcol = (pg & 1) * 4;
    const int span = (GT_P - 1) * a.stride + a.ks;  // input rows / columns one output tile reads
    for (long tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int nt = tile % n_tiles;
        long pt = tile / n_tiles;
        const int tw = pt % tiles_w;
        pt /= tiles_w;
        const int th = pt % tiles_h;
        const int b = pt / tiles_h;
        const int oh0 = th * GT_P, ow0 = tw * GT_P, n0 = nt * GT_N;
        const int ih0 = oh0 * a.stride - a.pad, iw0 = ow0 * a.stride - a.pad;
        double accd[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) accd[i][j] = 0.0;
#pragma unroll 1
        for (int c0 = 0; c0 < a.Cin; c0 += a.ck) {
            __syncthreads();
            // input patch [ci][r][x]
            for (int i = threadIdx.x; i < a.ck * span * span; i += blockDim.x) {
                int ci, r, x;
                if (a.in_nchw) { x = i % span; r = (i / span) % span; ci = i / (span * span); }
                else { ci = i % a.ck; x = (i / a.ck) % span; r = i / (a.ck * span); }
                const int hh = ih0 + r, ww = iw0 + x, c = c0 + ci;
                float v = 0.f;
                if (hh >= 0 && hh < a.H && ww >= 0 && ww < a.W && c < a.Cin)
                    v = a.in_nchw ? __ldg(a.in + (((long)b * a.Cin + c) * a.H + hh) * a.W + ww)
                                  : __ldg(a.in + (((long)b * a.H + hh) * a.W + ww) * a.Cin + c);
                in_s[(ci * span + r) * span + x] = v;
            }
#pragma unroll 1
            for (int ky = 0; ky < a.ks; ++ky) {
                __syncthreads();
                // one kernel row of weights [kx][ci][n]
                for (int i = threadIdx.x; i < a.ks * a.ck * GT_N; i += blockDim.x) {
                    const int n = i % GT_N;
                    const int ci = (i / GT_N) % a.ck;
                    const int kx = i / (GT_N * a.ck);
                    float v = 0.f;
                    if (n0 + n < a.Cout && c0 + ci < a.Cin)
                        v = __ldg(a.w + (((long)(n0 + n) * a.Cin + c0 + ci) * a.ks + ky) * a.ks + kx);
                    w_s[i] = v;
                }
                __syncthreads();
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 1
                for (int kx = 0; kx < a.ks; ++kx)
#pragma unroll 1
                    for (int ci = 0; ci < a.ck; ++ci) {
                        const float4 wv = *reinterpret_cast<const float4*>(&w_s[(kx * a.ck + ci) * GT_N + cg * 4]);
                        const float* row = in_s + (ci * span + prow * a.stride + ky) * span + kx;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float xv = row[(pcol + i) * a.stride];
                            acc[i][0] = fmaf(xv, wv.x, acc[i][0]);
                            acc[i][1] = fmaf(xv, wv.y, acc[i][1]);
                            acc[i][2] = fmaf(xv, wv.z, acc[i][2]);
                            acc[i][3] = fmaf(xv, wv.w, acc[i][3]);
                        }
                    }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) accd[i][j] += (double)acc[i][j];
            }
        }
        const int oh = oh0 + prow;
        if (oh >= a.OH) continue;
#pragma unroll 1
        for (int i = 0; i < 4; ++i) {
            const int ow = ow0 + pcol + i;
            if (ow >= a.OW) break;
            const long pix = ((long)b * a.OH + oh) * a.OW + ow;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + cg * 4 + j;
                if (n >= a.Cout) break;
                double vd = accd[i][j];
                if (a.bias) vd += (double)a.bias[n];
                float v = (float)vd;
                if (a.relu) v = fmaxf(v, 0.f);
                if (a.round) v = round_tf32(v);
                a.out[pix * a.Cout + n] = v;
            }
        }
    }
}

// Backward of the image layer (models.py:83, 11x11 / 4): gimg[c][y][x] = sum over the <= 3 x 3 output positions whose window
// covers (y, x) and over the output channels, + the TVLoss / temporal ContentLoss gradients of the image (ImageTail).
__global__ void __launch_bounds__(128)
conv_gen_dgrad_img_kernel(const float* __restrict__ gout /*NHWC [B][OH][OW][Cout]*/, const float* __restrict__ w /*[Cout][3][ks][ks]*/,
                          float* __restrict__ gimg, int B, int H, int W, int Cout, int ks, int stride, int OH, int OW,
                          ImageTail tail) {
    const long HW = (long)H * W;
    const long total = (long)B * HW;
    const float tvc = tail.tv_coef ? *tail.tv_coef : 0.f;
    const float tpc = tail.temp_coef ? *tail.temp_coef : 0.f;
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < total; p += (long)gridDim.x * blockDim.x) {
        const int x = p % W;
        const int y = (p / W) % H;
        const int b = p / HW;
        double acc[3] = {0.0, 0.0, 0.0};
        if (gout) {
            const int oy_lo = y - ks + 1 > 0 ? (y - ks + 1 + stride - 1) / stride : 0;
            const int ox_lo = x - ks + 1 > 0 ? (x - ks + 1 + stride - 1) / stride : 0;
            for (int oy = oy_lo; oy < OH && oy * stride <= y; ++oy)
                for (int ox = ox_lo; ox < OW && ox * stride <= x; ++ox) {
                    const int ky = y - oy * stride, kx = x - ox * stride;
                    const float* g = gout + (((long)b * OH + oy) * OW + ox) * Cout;
                    float part[3] = {0.f, 0.f, 0.f};
                    for (int co = 0; co < Cout; ++co) {
                        const float gv = g[co];
                        const float* wp = w + (((long)co * 3) * ks + ky) * ks + kx;
                        part[0] = fmaf(gv, __ldg(wp), part[0]);
                        part[1] = fmaf(gv, __ldg(wp + ks * ks), part[1]);
                        part[2] = fmaf(gv, __ldg(wp + 2 * ks * ks), part[2]);
                    }
                    acc[0] += part[0]; acc[1] += part[1]; acc[2] += part[2];
                }
        }
        const float wt = (tail.temp_coef && tail.temp_weights) ? tail.temp_weights[(long)y * W + x] : 1.f;
        for (int ci = 0; ci < 3; ++ci) {
            const long idx = ((long)b * 3 + ci) * HW + (long)y * W + x;
            float v = (float)acc[ci];
            if (tail.tv_coef) {  // loss.py:229-233: strength * (sum |x[h+1] - x[h]| + sum |x[w+1] - x[w]|), sign(0) = 0
                const float c0 = tail.img[idx];
                float s = 0.f;
                if (y > 0) s += (float)((c0 - tail.img[idx - W] > 0.f) - (c0 - tail.img[idx - W] < 0.f));
                if (y + 1 < H) s -= (float)((tail.img[idx + W] - c0 > 0.f) - (tail.img[idx + W] - c0 < 0.f));
                if (x > 0) s += (float)((c0 - tail.img[idx - 1] > 0.f) - (c0 - tail.img[idx - 1] < 0.f));
                if (x + 1 < W) s -= (float)((tail.img[idx + 1] - c0 > 0.f) - (tail.img[idx + 1] - c0 < 0.f));
                v += tvc * s;
            }
            if (tail.temp_coef) v += tpc * wt * (tail.img[idx] * wt - tail.temp_target[idx]);
            gimg[idx] = v;
        }
    }
}

// ---- the image layer as a GEMM (product path): im2col of the NCHW image, pointwise tcgen05 GEMM, and col2im for its gradient ----
// A[p][k], p = output position, k = c * ks^2 + ky * ks + kx (the OIHW order of a weight row), zero-padded to KP columns
__global__ void im2col_img_kernel(const float* __restrict__ img, float* __restrict__ A, int H, int W, int ks, int stride, int OH,
                                  int OW, int K, int KP, int do_round) {
    // one thread per (position, image row of the window): ks contiguous pixels of the image -> ks contiguous columns of A;
    // the thread with row index 3 * ks writes the zero padding K .. KP - 1
    const int rows = 3 * ks + 1;
    const long total = (long)OH * OW * rows;
    const long HW = (long)H * W;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int r = i % rows;
        const long p = i / rows;
        float* dst = A + p * KP;
        if (r == 3 * ks) {
            for (int k = K; k < KP; ++k) dst[k] = 0.f;
            continue;
        }
        const int ox = p % OW, oy = p / OW;
        const int c = r / ks, ky = r - c * ks;
        const float* src = img + c * HW + (long)(oy * stride + ky) * W + ox * stride;
        dst += r * ks;
        for (int kx = 0; kx < ks; ++kx) {
            const float v = __ldg(src + kx);
            dst[kx] = do_round ? round_tf32(v) : v;
        }
    }
}

// gimg[c][y][x] = sum over the output positions whose window covers (y, x) of T[position][c * ks^2 + ky * ks + kx], + image tail
__global__ void __launch_bounds__(256)
col2im_img_kernel(const float* __restrict__ T, float* __restrict__ gimg, int H, int W, int ks, int stride, int OH, int OW, int KP,
                  ImageTail tail) {
    const long HW = (long)H * W;
    const int kk = ks * ks;
    const float tvc = tail.tv_coef ? *tail.tv_coef : 0.f;
    const float tpc = tail.temp_coef ? *tail.temp_coef : 0.f;
    for (long p = blockIdx.x * (long)blockDim.x + threadIdx.x; p < HW; p += (long)gridDim.x * blockDim.x) {
        const int x = p % W, y = p / W;
        float acc[3] = {0.f, 0.f, 0.f};
        if (T) {
            const int oy_lo = y - ks + 1 > 0 ? (y - ks + 1 + stride - 1) / stride : 0;
            const int ox_lo = x - ks + 1 > 0 ? (x - ks + 1 + stride - 1) / stride : 0;
            for (int oy = oy_lo; oy < OH && oy * stride <= y; ++oy)
                for (int ox = ox_lo; ox < OW && ox * stride <= x; ++ox) {
                    const float* t = T + ((long)oy * OW + ox) * KP + (y - oy * stride) * ks + (x - ox * stride);
                    acc[0] += __ldg(t); acc[1] += __ldg(t + kk); acc[2] += __ldg(t + 2 * kk);
                }
        }
        const float wt = (tail.temp_coef && tail.temp_weights) ? tail.temp_weights[p] : 1.f;
        for (int ci = 0; ci < 3; ++ci) {
            const long idx = ci * HW + p;
            float v = acc[ci];
            if (tail.tv_coef) {
                const float c0 = tail.img[idx];
                float s = 0.f;
                if (y > 0) s += (float)((c0 - tail.img[idx - W] > 0.f) - (c0 - tail.img[idx - W] < 0.f));
                if (y + 1 < H) s -= (float)((tail.img[idx + W] - c0 > 0.f) - (tail.img[idx + W] - c0 < 0.f));
                if (x > 0) s += (float)((c0 - tail.img[idx - 1] > 0.f) - (c0 - tail.img[idx - 1] < 0.f));
                if (x + 1 < W) s -= (float)((tail.img[idx + 1] - c0 > 0.f) - (tail.img[idx + 1] - c0 < 0.f));
                v += tvc * s;
            }
            if (tail.temp_coef) v += tpc * wt * (tail.img[idx] * wt - tail.temp_target[idx]);
            gimg[idx] = v;
        }
    }
}

// GEMM weights of the image layer: wg[co][k] = w[co][k] (k < K, else 0) and its transpose wt[k][co], TF32-rounded
__global__ void im2col_weights_kernel(const float* __restrict__ w, float* __restrict__ wg, float* __restrict__ wt, int Cout, int K,
                                      int KP) {
    const long total = (long)Cout * KP;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int k = i % KP;
        const int co = i / KP;
        const float v = k < K ? round_tf32(w[(long)co * K + k]) : 0.f;
        wg[i] = v;
        wt[(long)k * Cout + co] = v;
    }
}

// [Cout][Cin][ks][ks] -> [Cin][Cout][ks][ks] rotated by 180 degrees: the weights with which the stride-1 input gradient is again a
// direct convolution (pad' = ks - 1 - pad)
__global__ void flip_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int ks) {
    const long total = (long)Cout * Cin * ks * ks;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int kx = i % ks;
        const int ky = (i / ks) % ks;
        const int ci = (i / (ks * ks)) % Cin;
        const int co = i / ((long)ks * ks * Cin);
        out[(((long)ci * Cout + co) * ks + (ks - 1 - ky)) * ks + (ks - 1 - kx)] = w[i];
    }
}

// ---- MaxPool2d / AvgPool2d((3,3), (2,2), (0,0), ceil_mode=True) on NHWC --------------------------------------------------------
// Window (ph, pw) covers rows 2ph .. min(2ph + 3, H) - 1: the last window of a ceil_mode pool is clipped, and the average divides by
// the clipped size (ATen avg_pool2d with padding 0).  Max pooling: first maximum in row-major window order (ATen).
__global__ void pool3_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, int B, int H, int W, int C4, int PH, int PW,
                                 int avg, int do_round) {
    const long total = (long)B * PH * PW * C4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = i % C4;
        long r = i / C4;
        const int pw = r % PW;
        r /= PW;
        const int ph = r % PH;
        const int b = r / PH;
        const int h1 = min(2 * ph + 3, H), w1 = min(2 * pw + 3, W);
        float4 o = avg ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int h = 2 * ph; h < h1; ++h)
            for (int w = 2 * pw; w < w1; ++w) {
                const float4 v = x[(((long)b * H + h) * W + w) * C4 + c];
                if (avg) { o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w; }
                else { o.x = fmaxf(o.x, v.x); o.y = fmaxf(o.y, v.y); o.z = fmaxf(o.z, v.z); o.w = fmaxf(o.w, v.w); }
            }
        if (avg) {
            const float d = (float)((h1 - 2 * ph) * (w1 - 2 * pw));
            o.x /= d; o.y /= d; o.z /= d; o.w /= d;
            if (do_round) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        }
        y[i] = o;
    }
}

// gx[h][w][c] = (sum over the <= 2 x 2 windows that contain (h, w) of: gy[window] if (h, w) is the window's arg-max [max] or
// gy[window] / window size [avg]) * (x > 0) [+ addend * (x > 0)].  Gather form: no atomics, deterministic.  Four channels per thread.
__global__ void pool3_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ gy, const float4* __restrict__ addend,
                                 float4* __restrict__ gx, int B, int H, int W, int C4, int PH, int PW, int avg, int do_round) {
    const long total = (long)B * H * W * C4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = i % C4;
        long r = i / C4;
        const int w = r % W;
        r /= W;
        const int h = r % H;
        const int b = r / H;
        const float4 xv = x[i];
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (xv.x > 0.f || xv.y > 0.f || xv.z > 0.f || xv.w > 0.f) {
            for (int ph = max((h - 1) / 2, 0); ph < PH && 2 * ph <= h; ++ph) {
                if (h >= 2 * ph + 3) continue;
                for (int pw = max((w - 1) / 2, 0); pw < PW && 2 * pw <= w; ++pw) {
                    if (w >= 2 * pw + 3) continue;
                    const float4 gv = gy[(((long)b * PH + ph) * PW + pw) * C4 + c];
                    const int h1 = min(2 * ph + 3, H), w1 = min(2 * pw + 3, W);
                    if (avg) {
                        const float inv = 1.f / (float)((h1 - 2 * ph) * (w1 - 2 * pw));
                        g.x += gv.x * inv; g.y += gv.y * inv; g.z += gv.z * inv; g.w += gv.w * inv;
                    } else {
                        // is (h, w) the first maximum of this window? (per channel)
                        bool wx = true, wy = true, wz = true, ww_ = true;
                        for (int hh = 2 * ph; hh < h1; ++hh)
                            for (int ww = 2 * pw; ww < w1; ++ww) {
                                if (hh == h && ww == w) continue;
                                const float4 v = x[(((long)b * H + hh) * W + ww) * C4 + c];
                                if (hh < h || (hh == h && ww < w)) {
                                    wx = wx && !(v.x >= xv.x); wy = wy && !(v.y >= xv.y); wz = wz && !(v.z >= xv.z); ww_ = ww_ && !(v.w >= xv.w);
                                } else {
                                    wx = wx && !(v.x > xv.x); wy = wy && !(v.y > xv.y); wz = wz && !(v.z > xv.z); ww_ = ww_ && !(v.w > xv.w);
                                }
                            }
                        if (wx) g.x += gv.x;
                        if (wy) g.y += gv.y;
                        if (wz) g.z += gv.z;
                        if (ww_) g.w += gv.w;
                    }
                }
            }
            if (addend) {
                const float4 a = addend[i];
                g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
            }
        }
        g.x = xv.x > 0.f ? g.x : 0.f; g.y = xv.y > 0.f ? g.y : 0.f; g.z = xv.z > 0.f ? g.z : 0.f; g.w = xv.w > 0.f ? g.w : 0.f;
        if (do_round) { g.x = round_tf32(g.x); g.y = round_tf32(g.y); g.z = round_tf32(g.z); g.w = round_tf32(g.w); }
        gx[i] = g;
    }
}

}  // namespace

int conv_gen_fwd_launch(const float* in, int in_nchw, const float* w, const float* bias, float* out, uint32_t* mask_out, int B,
                        int H, int W, int Cin, int Cout, int ks, int stride, int pad, int relu, int round, cudaStream_t st) {
    MAUA_REQUIRE(in && w && out, "conv_gen: null pointer");
    MAUA_REQUIRE(ks >= 1 && ks <= 11 && stride >= 1 && pad >= 0, "conv_gen: bad geometry k %d stride %d pad %d", ks, stride, pad);
    GenConv a;
    a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.ks = ks; a.stride = stride; a.pad = pad;
    a.OH = (H + 2 * pad - ks) / stride + 1;
    a.OW = (W + 2 * pad - ks) / stride + 1;
    MAUA_REQUIRE(H + 2 * pad >= ks && W + 2 * pad >= ks, "conv_gen: %dx%d input is smaller than the %dx%d kernel", H, W, ks, ks);
    a.in_nchw = in_nchw; a.in = in; a.w = w; a.bias = bias; a.out = out; a.relu = relu; a.round = round;
    const int span = (GT_P - 1) * stride + ks;
    int ck = G_IN_MAX / (span * span);
    if (ck > 8) ck = 8;
    if (ck > Cin) ck = Cin;
    MAUA_REQUIRE(ck >= 1 && ks * ck * GT_N <= G_W_MAX, "conv_gen: kernel %dx%d stride %d does not fit the staging buffers", ks, ks, stride);
    a.ck = ck;
    const long tiles = (long)B * ((a.OH + GT_P - 1) / GT_P) * ((a.OW + GT_P - 1) / GT_P) * ((Cout + GT_N - 1) / GT_N);
    conv_gen_kernel<<<(int)(tiles > 148L * 16 ? 148L * 16 : tiles), 256, 0, st>>>(a);
    MAUA_CUDA_CHECK(cudaGetLastError());
    if (mask_out) return relu_mask_bits_launch(out, mask_out, (long)B * a.OH * a.OW, Cout, st);
    return MAUA_OK;
}

int conv_gen_flip_weights_launch(const float* w, float* out, int Cout, int Cin, int ks, cudaStream_t st) {
    const long total = (long)Cout * Cin * ks * ks;
    flip_weights_kernel<<<(int)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256), 256, 0, st>>>(w, out, Cout, Cin, ks);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

int conv_gen_dgrad_img_launch(const float* gout, const float* w, float* gimg, int B, int H, int W, int Cout, int ks, int stride,
                              const ImageTail& tail, cudaStream_t st) {
    MAUA_REQUIRE(gimg && (gout == nullptr || w), "conv_gen_dgrad_img: null pointer");
    const int OH = (H - ks) / stride + 1, OW = (W - ks) / stride + 1;
    const long total = (long)B * H * W;
    conv_gen_dgrad_img_kernel<<<(int)((total + 127) / 128 > 148L * 32 ? 148L * 32 : (total + 127) / 128), 128, 0, st>>>(
        gout, w, gimg, B, H, W, Cout, ks, stride, OH, OW, tail);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

int im2col_img_launch(const float* img, float* A, int H, int W, int ks, int stride, int KP, int do_round, cudaStream_t st) {
    MAUA_REQUIRE(img && A && H >= ks && W >= ks && KP >= 3 * ks * ks, "im2col: bad arguments");
    const int OH = (H - ks) / stride + 1, OW = (W - ks) / stride + 1;
    const long total = (long)OH * OW * (3 * ks + 1);
    im2col_img_kernel<<<(int)((total + 255) / 256 > 148L * 64 ? 148L * 64 : (total + 255) / 256), 256, 0, st>>>(
        img, A, H, W, ks, stride, OH, OW, 3 * ks * ks, KP, do_round);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

int col2im_img_launch(const float* T, float* gimg, int H, int W, int ks, int stride, int KP, const ImageTail& tail, cudaStream_t st) {
    MAUA_REQUIRE(gimg, "col2im: null output");
    const int OH = (H - ks) / stride + 1, OW = (W - ks) / stride + 1;
    const long total = (long)H * W;
    col2im_img_kernel<<<(int)((total + 255) / 256 > 148L * 32 ? 148L * 32 : (total + 255) / 256), 256, 0, st>>>(
        T, gimg, H, W, ks, stride, OH, OW, KP, tail);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

int im2col_weights_launch(const float* w, float* wg, float* wt, int Cout, int K, int KP, cudaStream_t st) {
    const long total = (long)Cout * KP;
    im2col_weights_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(w, wg, wt, Cout, K, KP);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

void pool3_out_extent(int H, int W, int* PH, int* PW) {
    // ATen pooling_output_shape with ceil_mode, pad 0: floor((H - 3 + 1) / 2) + 1
    // (n = 2 is one clipped window; n = 1 gives 0: ATen refuses it as "output size is too small"; with kernel 3 / stride 2 the
    //  last window always starts inside the input, so ATen's (o - 1) * stride >= n correction never fires)
    auto f = [](int n) { return n < 2 ? 0 : (n - 2) / 2 + 1; };
    *PH = f(H); *PW = f(W);
}

int pool3_fwd_launch(const float* x, float* y, int B, int H, int W, int C, int avg, int do_round, cudaStream_t st) {
    MAUA_REQUIRE(C % 4 == 0, "pool: C %% 4 != 0");
    int PH, PW;
    pool3_out_extent(H, W, &PH, &PW);
    const long total = (long)B * PH * PW * (C / 4);
    pool3_fwd_kernel<<<(int)((total + 255) / 256 > 148L * 16 ? 148L * 16 : (total + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), B, H, W, C / 4, PH, PW, avg, do_round);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

int pool3_bwd_launch(const float* x, const float* gy, const float* addend, float* gx, int B, int H, int W, int C, int avg,
                     int do_round, cudaStream_t st) {
    int PH, PW;
    pool3_out_extent(H, W, &PH, &PW);
    MAUA_REQUIRE(C % 4 == 0, "pool: C %% 4 != 0");
    const long total = (long)B * H * W * (C / 4);
    pool3_bwd_kernel<<<(int)((total + 255) / 256 > 148L * 32 ? 148L * 32 : (total + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(gy), reinterpret_cast<const float4*>(addend),
        reinterpret_cast<float4*>(gx), B, H, W, C / 4, PH, PW, avg, do_round);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

}  // namespace maua
