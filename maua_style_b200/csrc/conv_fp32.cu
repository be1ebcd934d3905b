// Exact-arithmetic twin of conv_tc.cu: the same ConvArgs / ConvEpilogue contract executed with FP32 operands on the
// CUDA cores (no TF32 operand rounding, no tensor-core accumulate).  MAUA_IMPL_FP32 plans run every GEMM-shaped launch
// through this kernel, so that the plan logic around it -- ReLU sign bitmaps, max-pool arg-max recomputation, the folded
// StyleLoss backward, the content / addend epilogue terms, pool scatter -- can be checked against the reference's fp32
// results (models.py:116-132, optim.py:201-221 autograd) to the 1e-3 the north_star states, also under max pooling where
// the TF32 path's error is dominated by arg-max flips.
//
// Arithmetic: every output is the sum over (tap, channel) of fp32 products.  Products of one 8-channel chunk (72 terms for
// a 3x3 layer) are accumulated with FFMA in fp32, and the chunk sums are added in fp64, so the result does not depend on
// the length of K and is within about one fp32 ulp of the exact sum -- closer to the fp64 truth than any pure-fp32
// summation order, i.e. it does not add arg-max flips of its own.
//
// Tile: 8 x 16 pixels x 64 output channels per 256-thread block; a thread owns 8 consecutive pixels of one tile row and 4
// channels.  Works for any channel counts (edges are predicated), which the tensor-core path does not.
#include "conv_tc.cuh"
#include "pointwise.cuh"

namespace maua {

namespace {

constexpr int FT_H = 8, FT_W = 16, FT_N = 64, FCK = 8;
constexpr int IN_W = FT_W + 2 + 2;  // padded row of the staged halo
constexpr int WS_N = FT_N + 4;      // padded channel row of the staged weights (conflict-free stores, 16-byte rows)

struct Fp32Term {
    const float* in;  // NHWC [B][H][W][C]
    const float* w;   // [Cout][taps * C]
    int C, taps;
};

__global__ void __launch_bounds__(256)
conv_fp32_kernel(const ConvArgs a) {
    __shared__ __align__(16) float in_s[FCK][FT_H + 2][IN_W];
    __shared__ __align__(16) float w_s[9][FCK][WS_N];
    const int tiles_w = (a.W + FT_W - 1) / FT_W, tiles_h = (a.H + FT_H - 1) / FT_H;
    const int n_tiles = (a.Cout + FT_N - 1) / FT_N;
    const long total_tiles = (long)a.B * tiles_h * tiles_w * n_tiles;
    const int cg = threadIdx.x & 15;   // channel group: 4 channels
    const int pg = threadIdx.x >> 4;   // pixel group: row pg / 2, columns (pg % 2) * 8 .. + 7
    const int prow = pg >> 1, pcol = (pg & 1) * 8;
    const ConvEpilogue& ep = a.ep;
    const float ccoef = ep.cont_f ? *ep.cont_coef : 0.f;

    for (long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % n_tiles;
        long pt = tile / n_tiles;
        const int tw = pt % tiles_w;
        pt /= tiles_w;
        const int th = pt % tiles_h;
        const int b = pt / tiles_h;
        const int h0 = th * FT_H, w0 = tw * FT_W, n0 = nt * FT_N;

        double accd[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) accd[i][j] = 0.0;

        Fp32Term terms[2] = {{a.in, a.wg, a.ntaps > 0 ? a.Cin : 0, a.ntaps}, {a.in2, a.w2, a.K2, 1}};
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
            const Fp32Term tm = terms[t];
            if (tm.C <= 0) continue;
            const int halo = tm.taps == 9 ? 1 : 0;
            const int rows = FT_H + 2 * halo, cols = FT_W + 2 * halo;
#pragma unroll 1
            for (int c0 = 0; c0 < tm.C; c0 += FCK) {
                __syncthreads();  // the previous chunk has been consumed
                for (int i = threadIdx.x; i < rows * cols * FCK; i += blockDim.x) {
                    const int ci = i % FCK;
                    const int x = (i / FCK) % cols;
                    const int r = i / (FCK * cols);
                    const int hh = h0 + r - halo, ww = w0 + x - halo;
                    float v = 0.f;
                    if (hh >= 0 && hh < a.H && ww >= 0 && ww < a.W && c0 + ci < tm.C)
                        v = __ldg(tm.in + (((long)b * a.H + hh) * a.W + ww) * tm.C + c0 + ci);
                    in_s[ci][r][x] = v;
                }
                for (int i = threadIdx.x; i < tm.taps * FT_N * FCK; i += blockDim.x) {
                    const int ci = i % FCK;
                    const int n = (i / FCK) % FT_N;
                    const int tap = i / (FCK * FT_N);
                    float v = 0.f;
                    if (n0 + n < a.Cout && c0 + ci < tm.C)
                        v = __ldg(tm.w + (long)(n0 + n) * tm.taps * tm.C + (long)tap * tm.C + c0 + ci);
                    w_s[tap][ci][n] = v;
                }
                __syncthreads();
                float acc[8][4];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
                if (tm.taps == 9) {
#pragma unroll 2
                    for (int ci = 0; ci < FCK; ++ci)
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
                            float xin[10];
#pragma unroll
                            for (int k = 0; k < 10; ++k) xin[k] = in_s[ci][prow + dy][pcol + k];
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx) {
                                const float4 wv = *reinterpret_cast<const float4*>(&w_s[dy * 3 + dx][ci][cg * 4]);
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    acc[i][0] = fmaf(xin[i + dx], wv.x, acc[i][0]);
                                    acc[i][1] = fmaf(xin[i + dx], wv.y, acc[i][1]);
                                    acc[i][2] = fmaf(xin[i + dx], wv.z, acc[i][2]);
                                    acc[i][3] = fmaf(xin[i + dx], wv.w, acc[i][3]);
                                }
                            }
                        }
                } else {
#pragma unroll
                    for (int ci = 0; ci < FCK; ++ci) {
                        const float4 wv = *reinterpret_cast<const float4*>(&w_s[0][ci][cg * 4]);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float xv = in_s[ci][prow][pcol + i];
                            acc[i][0] = fmaf(xv, wv.x, acc[i][0]);
                            acc[i][1] = fmaf(xv, wv.y, acc[i][1]);
                            acc[i][2] = fmaf(xv, wv.z, acc[i][2]);
                            acc[i][3] = fmaf(xv, wv.w, acc[i][3]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) accd[i][j] += (double)acc[i][j];
            }
        }

        // fused epilogue, same order as ConvEpilogue documents
        const int h = h0 + prow;
        if (h >= a.H) continue;
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
            const int w = w0 + pcol + i;
            if (w >= a.W) break;
            const long pix = ((long)b * a.H + h) * a.W + w;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + cg * 4 + j;
                if (n >= a.Cout) break;
                const long idx = pix * a.Cout + n;
                double vd = accd[i][j];
                if (ep.bias) vd += (double)ep.bias[n];
                float v = (float)vd;
                if (ep.cont_f) v += ccoef * (ep.cont_f[idx] - ep.cont_t[idx]);
                if (ep.addend) v += ep.addend[idx];
                if (ep.relu) v = fmaxf(v, 0.f);
                if (ep.mask_src) v = ep.mask_src[idx] > 0.f ? v : 0.f;
                if (ep.mask_bits) v = ((ep.mask_bits[pix * (a.Cout >> 5) + (n >> 5)] >> (n & 31)) & 1u) ? v : 0.f;
                if (ep.round) v = round_tf32(v);
                ep.out[idx] = v;
                if (ep.out2) ep.out2[idx] = v;
            }
        }
    }
}

}  // namespace

int conv_fp32_launch(const ConvArgs& a, cudaStream_t st) {
    MAUA_REQUIRE(a.ntaps == 9 || a.ntaps == 1 || a.ntaps == 0, "ntaps must be 9, 1 or 0 (got %d)", a.ntaps);
    MAUA_REQUIRE(a.ntaps > 0 || a.K2 > 0, "conv has neither a main nor an aux term");
    MAUA_REQUIRE(a.ep.out != nullptr, "null output pointer");
    MAUA_REQUIRE(a.B >= 1 && a.H >= 1 && a.W >= 1, "bad extent B=%d H=%d W=%d", a.B, a.H, a.W);
    MAUA_REQUIRE(!a.ep.mask_bits || a.Cout % 32 == 0, "sign bitmaps need Cout %% 32 == 0 (got %d)", a.Cout);
    const long tiles = (long)a.B * ((a.H + FT_H - 1) / FT_H) * ((a.W + FT_W - 1) / FT_W) * ((a.Cout + FT_N - 1) / FT_N);
    const int blocks = (int)(tiles > 148L * 16 ? 148L * 16 : tiles);
    conv_fp32_kernel<<<blocks, 256, 0, st>>>(a);
    MAUA_CUDA_CHECK(cudaGetLastError());
    if (a.ep.mask_out) {
        const int rc = relu_mask_bits_launch(a.ep.out, a.ep.mask_out, (long)a.B * a.H * a.W, a.Cout, st);
        if (rc) return rc;
    }
    if (a.ep.pool_out) return pool_fwd_launch(a.ep.out, a.ep.pool_out, a.B, a.H, a.W, a.Cout, a.ep.pool_avg, a.ep.round, st, a.ep.pool_codes);
    return MAUA_OK;
}

}  // namespace maua
