// The image-side edge of the VGG stack: conv1_1 (Cin = 3) forward and its input-gradient.
//
// conv1_1 has K = 27: far too thin for a tensor-core tile, and the layer is bound by writing (forward) or reading
// (backward) the 64-channel full-resolution map.  The image itself stays NCHW [B,3,H,W] (the layout of the reference's
// pastiche, optim.py:173), which is not TMA-friendly for W % 4 != 0.
//
// forward : direct fp32 FFMA2 kernel, thread = 8 pixels x 16 output channels, image halo staged in shared memory,
//           weights broadcast from shared memory, NHWC output written as float4 (bias + ReLU + TF32 rounding fused),
//           TVLoss value folded in (the tile already holds every pixel's neighbours).
// backward: two steps.  (1) a pointwise tensor-core GEMM (conv_tc.cu, ntaps = 1, Cin = 64, Cout = 32) contracts the
//           64 channels of the masked gradient with the 27 (tap, ci) weight columns per pixel:
//               T[p][tap*3 + ci] = sum_co gout[p][co] * W[co][ci][ky][kx]
//           (2) a memory-bound gather sums the 9 shifted taps,  gimg[ci][h][w] = sum_tap T[(h-ky+1, w-kx+1)][tap*3+ci],
//           and is also the fused "image-side tail" (SURVEY.md A.4 item 5): it adds the TVLoss gradient (loss.py:224-233)
//           and the temporal ContentLoss gradient (loss.py:46-54) while the image gradient is in registers, so
//           pastiche.grad is written exactly once.
#include <cstdlib>

#include "conv_tc.cuh"
#include "reduce.cuh"

namespace maua {

namespace {

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
constexpr int FG_STRIDE = 20;  // 16 channels + 4 pad floats per group: conflict-free float4 broadcast reads
constexpr int FT_W = 64;       // tile: 64 x 4 pixels, 128 threads = 8 pixel octets x 4 rows x 4 channel groups
constexpr int FT_H = 4;
constexpr int FI_W = 68;       // staged image row: 64 + 2 halo, padded to a float4 multiple
constexpr int FT_THREADS = 128;

// Thread = 8 consecutive pixels of one tile row x 16 output channels (channel c4*16 + g*4 + i for c4, i in 0..3, so that
// the four threads of a pixel write 64 contiguous bytes per store).  Per (ci, ky) a thread reads 10 image values and per
// tap 16 weights (four broadcast LDS.128) for 128 multiply-adds: twice the arithmetic per shared-memory byte of the
// round-1 kernel (4 pixels x 16 channels), which was bound by the weight reads, not by the FMA pipe.  PACKED issues the
// multiply-adds as FFMA2 (fma.rn.f32x2, sm_100): two round-to-nearest FMAs per instruction, bit-identical results.
// The TVLoss value (loss.py:229-233) is folded in when tv_out is given: the staged tile already holds every pixel's upper
// and left neighbour, so the image is read once per forward instead of twice.
template <bool PACKED>
__global__ void __launch_bounds__(FT_THREADS, 3)
conv_first_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ out, uint16_t* __restrict__ mask16, int B, int H, int W, int do_round,
                      float tv_strength, float* __restrict__ tv_out, double* tv_partials, unsigned int* tv_counter) {
    constexpr int Cout = 64, G = 4;
    __shared__ __align__(16) float ws[27 * G * FG_STRIDE];
    __shared__ float bs[Cout];
    __shared__ __align__(16) float inp[3][FT_H + 2][FI_W];
    for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
        const int co = i / 27, k = i % 27;
        const int c4 = co >> 4, g = (co >> 2) & 3, e = co & 3;
        ws[(k * G + g) * FG_STRIDE + c4 * 4 + e] = w[i];
    }
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) bs[i] = bias ? bias[i] : 0.f;
    pdl_wait();  // the (frozen) weights were staged while the optimizer kernel that writes the image was still draining
    pdl_trigger();

    const int g = threadIdx.x & 3;
    const int oct = (threadIdx.x >> 2) & 7;
    const int row = threadIdx.x >> 5;
    const int tiles_w = (W + FT_W - 1) / FT_W, tiles_h = (H + FT_H - 1) / FT_H;
    const long ntiles = (long)B * tiles_w * tiles_h;
    const long HW = (long)H * W;
    float tv_local = 0.f;

    for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tw = t % tiles_w;
        const int th = (t / tiles_w) % tiles_h;
        const int b = t / ((long)tiles_w * tiles_h);
        const int h0 = th * FT_H, w0 = tw * FT_W;
        __syncthreads();  // previous tile consumed (and, first time, weights staged)
        for (int i = threadIdx.x; i < 3 * (FT_H + 2) * (FT_W + 2); i += blockDim.x) {
            const int x = i % (FT_W + 2);
            const int r = (i / (FT_W + 2)) % (FT_H + 2);
            const int ci = i / ((FT_W + 2) * (FT_H + 2));
            const int hh = h0 + r - 1, ww = w0 + x - 1;
            inp[ci][r][x] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(img + ((long)b * 3 + ci) * HW + (long)hh * W + ww) : 0.f;
        }
        __syncthreads();

        const int h = h0 + row;
        if (tv_out && g < 3 && h < H) {  // TVLoss: channel ci = g of this thread's 8 pixels (row / column 0 have no neighbour)
            const float* cur = &inp[g][row + 1][8 * oct];
            const float* up = &inp[g][row][8 * oct];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int x = w0 + 8 * oct + j;
                if (x < W) {
                    const float v = cur[j + 1];
                    if (h > 0) tv_local += fabsf(v - up[j + 1]);
                    if (x > 0) tv_local += fabsf(v - cur[j]);
                }
            }
        }

        float2 acc[8][8];  // [pixel][channel pair]: channels c4*16 + g*4 + {0,1} and {2,3}
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                acc[j][2 * c4] = make_float2(bs[c4 * 16 + g * 4], bs[c4 * 16 + g * 4 + 1]);
                acc[j][2 * c4 + 1] = make_float2(bs[c4 * 16 + g * 4 + 2], bs[c4 * 16 + g * 4 + 3]);
            }
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float* ip = &inp[ci][row + ky][8 * oct];
                const float4 x0 = *reinterpret_cast<const float4*>(ip);
                const float4 x1 = *reinterpret_cast<const float4*>(ip + 4);
                const float2 x2 = *reinterpret_cast<const float2*>(ip + 8);
                const float xin[10] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y};
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4* wp = reinterpret_cast<const float4*>(ws + ((ci * 9 + ky * 3 + kx) * G + g) * FG_STRIDE);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 wv = wp[c4];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float xv = xin[j + kx];
                            if (PACKED) {
                                acc[j][2 * c4] = __ffma2_rn(make_float2(xv, xv), make_float2(wv.x, wv.y), acc[j][2 * c4]);
                                acc[j][2 * c4 + 1] = __ffma2_rn(make_float2(xv, xv), make_float2(wv.z, wv.w), acc[j][2 * c4 + 1]);
                            } else {
                                acc[j][2 * c4].x = fmaf(xv, wv.x, acc[j][2 * c4].x);
                                acc[j][2 * c4].y = fmaf(xv, wv.y, acc[j][2 * c4].y);
                                acc[j][2 * c4 + 1].x = fmaf(xv, wv.z, acc[j][2 * c4 + 1].x);
                                acc[j][2 * c4 + 1].y = fmaf(xv, wv.w, acc[j][2 * c4 + 1].y);
                            }
                        }
                    }
                }
            }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int x = w0 + 8 * oct + j;
            const bool valid = (h < H) && (x < W);
            const long pix = ((long)b * H + (valid ? h : 0)) * W + (valid ? x : 0);
            float* op = out + pix * Cout + g * 4;
            uint32_t mine = 0;  // nibble c4 = sign bits of this thread's channels c4*16 + g*4 + 0..3
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                float4 o;
                o.x = fmaxf(acc[j][2 * c4].x, 0.f); o.y = fmaxf(acc[j][2 * c4].y, 0.f);
                o.z = fmaxf(acc[j][2 * c4 + 1].x, 0.f); o.w = fmaxf(acc[j][2 * c4 + 1].y, 0.f);
                if (do_round) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
                if (valid) *reinterpret_cast<float4*>(op + c4 * 16) = o;
                mine |= ((o.x > 0.f ? 1u : 0u) | (o.y > 0.f ? 2u : 0u) | (o.z > 0.f ? 4u : 0u) | (o.w > 0.f ? 8u : 0u)) << (4 * c4);
            }
            // sign bitmap (the ReLU mask dgrad reads): half-word c4 of a pixel = channels c4*16 .. c4*16+15 = the nibbles c4
            // of the pixel's four threads; thread g assembles half-word g
            uint32_t hw = 0;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const uint32_t other = __shfl_sync(0xffffffffu, mine, (threadIdx.x & 28) | s);
                hw |= ((other >> (4 * g)) & 0xFu) << (4 * s);
            }
            if (mask16 && valid) mask16[pix * (Cout / 16) + g] = (uint16_t)hw;
        }
    }
    if (tv_out) {
        double v[1] = {(double)tv_local}, tot[1];
        if (grid_sum<1, FT_THREADS>(v, tv_partials, tv_counter, tot)) *tv_out = tv_strength * (float)tot[0];
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// wt[n][co], n = (ky*3 + kx)*3 + ci (27 rows, padded with zeros to 32): B operand of the per-pixel contraction
__global__ void first_dgrad_weights_kernel(const float* __restrict__ w, float* __restrict__ wt, int Cout, int do_round) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 32 * Cout) return;
    const int n = i / Cout, co = i % Cout;
    float v = 0.f;
    if (n < 27) {
        const int tap = n / 3, ci = n % 3;
        v = w[((long)co * 3 + ci) * 9 + tap];
        if (do_round) v = round_tf32(v);
    }
    wt[i] = v;
}

constexpr int GT = 16;      // gather tile: 16 x 16 output pixels per 256-thread block
constexpr int GH = GT + 2;
constexpr int GPS = 33;     // padded pixel stride (words): conflict-free scalar reads across adjacent pixels

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) - (v < 0.f); }

__global__ void __launch_bounds__(256)
conv_first_gather_kernel(const float* __restrict__ T /*NHWC [B][H][W][32]*/, float* __restrict__ gimg, int B, int H,
                         int W, ImageTail tail) {
    __shared__ float tile[GH * GH * GPS];
    pdl_wait();
    pdl_trigger();
    const int tiles_w = (W + GT - 1) / GT, tiles_h = (H + GT - 1) / GT;
    const long ntiles = (long)B * tiles_w * tiles_h;
    const long HW = (long)H * W;
    const float tvc = tail.tv_coef ? *tail.tv_coef : 0.f;
    const float tpc = tail.temp_coef ? *tail.temp_coef : 0.f;
    const int tx = threadIdx.x % GT, ty = threadIdx.x / GT;

    for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tw = t % tiles_w;
        const int th = (t / tiles_w) % tiles_h;
        const int b = t / ((long)tiles_w * tiles_h);
        const int h0 = th * GT, w0 = tw * GT;
        __syncthreads();
        // stage the 18 x 18 x 27(32) halo: 8 float4 per pixel, coalesced
        for (int i = threadIdx.x; i < GH * GH * 8; i += blockDim.x) {
            const int c4 = i & 7;
            const int pp = i >> 3;
            const int hy = pp / GH, hx = pp % GH;
            const int hh = h0 + hy - 1, ww = w0 + hx - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                v = __ldg(reinterpret_cast<const float4*>(T + (((long)b * H + hh) * W + ww) * 32) + c4);
            float* d = tile + pp * GPS + c4 * 4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        __syncthreads();
        const int h = h0 + ty, x = w0 + tx;
        if (h >= H || x >= W) continue;
        float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                // gimg[h][w] += T[h - ky + 1][w - kx + 1][tap]; halo coordinates add +1
                const float* p = tile + ((ty + 2 - ky) * GH + (tx + 2 - kx)) * GPS + (ky * 3 + kx) * 3;
                acc[0] += p[0]; acc[1] += p[1]; acc[2] += p[2];
            }
        const float wt = (tail.temp_coef && tail.temp_weights) ? tail.temp_weights[(long)h * W + x] : 1.f;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
            const long idx = ((long)b * 3 + ci) * HW + (long)h * W + x;
            float v = acc[ci];
            if (tail.tv_coef) {
                const float c0 = tail.img[idx];
                float s = 0.f;
                if (h > 0) s += sgn(c0 - tail.img[idx - W]);
                if (h + 1 < H) s -= sgn(tail.img[idx + W] - c0);
                if (x > 0) s += sgn(c0 - tail.img[idx - 1]);
                if (x + 1 < W) s -= sgn(tail.img[idx + 1] - c0);
                v += tvc * s;
            }
            if (tail.temp_coef) v += tpc * wt * (tail.img[idx] * wt - tail.temp_target[idx]);
            gimg[idx] = v;
        }
    }
}

}  // namespace

int conv_first_fwd_launch(const float* img, const float* w, const float* bias, float* out, uint32_t* mask_out, int B,
                          int H, int W, int Cout, int round, cudaStream_t st, const ConvFirstTV* tv) {
    MAUA_REQUIRE(Cout == 64, "conv_first_fwd: the image layer must have 64 output channels (got %d)", Cout);
    const long ntiles = (long)B * ((W + FT_W - 1) / FT_W) * ((H + FT_H - 1) / FT_H);
    long blocks = ntiles > 148L * 6 ? 148L * 6 : ntiles;
    MAUA_REQUIRE(!tv || (tv->out && tv->partials && tv->counter && blocks <= tv->max_blocks), "conv_first_fwd: bad TVLoss arguments");
    const char* f2 = getenv("MAUA_CONV1_FFMA2");  // packed FFMA2 by default (bit-identical to scalar FFMA, verified on B200)
    uint16_t* m16 = reinterpret_cast<uint16_t*>(mask_out);
    const float tvs = tv ? tv->strength : 0.f;
    float* tvo = tv ? tv->out : nullptr;
    double* tvp = tv ? tv->partials : nullptr;
    unsigned int* tvc = tv ? tv->counter : nullptr;
    if (!f2 || atoi(f2) != 0)
        MAUA_CUDA_CHECK(launch_pdl<PDL_EDGE>(conv_first_fwd_kernel<true>, dim3((unsigned)blocks), dim3(FT_THREADS), 0, st, img, w, bias, out, m16, B, H,
                                   W, round, tvs, tvo, tvp, tvc));
    else
        MAUA_CUDA_CHECK(launch_pdl<PDL_EDGE>(conv_first_fwd_kernel<false>, dim3((unsigned)blocks), dim3(FT_THREADS), 0, st, img, w, bias, out, m16, B, H,
                                   W, round, tvs, tvo, tvp, tvc));
    return MAUA_OK;
}

size_t conv_first_dgrad_workspace_bytes(int B, int H, int W) {
    return ((size_t)B * H * W * 32 + 32 * 64) * sizeof(float) + 256;
}

int conv_first_dgrad_prep_weights(const float* w, float* wt, int Cout, int do_round, cudaStream_t st) {
    MAUA_REQUIRE(Cout == 64, "conv_first_dgrad: the image layer must have 64 output channels (got %d)", Cout);
    first_dgrad_weights_kernel<<<(32 * Cout + 255) / 256, 256, 0, st>>>(w, wt, Cout, do_round);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

int conv_first_dgrad_launch(const float* gout, const float* wt, float* gimg, int B, int H, int W, int Cout,
                            const ImageTail& tail, float* T, int impl, cudaStream_t st) {
    MAUA_REQUIRE(Cout == 64, "conv_first_dgrad: the image layer must have 64 output channels (got %d)", Cout);
    MAUA_REQUIRE(T && (reinterpret_cast<uintptr_t>(T) & 15) == 0, "conv_first_dgrad: bad workspace");
    ConvArgs a;
    a.B = B; a.H = H; a.W = W; a.Cin = Cout; a.Cout = 32; a.ntaps = 1;
    a.in = gout; a.wg = wt;
    a.ep.out = T; a.ep.round = 0;
    int rc = conv_dispatch(a, impl, st);
    if (rc) return rc;
    const long ntiles = (long)B * ((W + GT - 1) / GT) * ((H + GT - 1) / GT);
    long blocks = ntiles > 148L * 8 ? 148L * 8 : ntiles;
    MAUA_CUDA_CHECK(launch_pdl<PDL_EDGE>(conv_first_gather_kernel, dim3((unsigned)blocks), dim3(256), 0, st, T, gimg, B, H, W, tail));
    return MAUA_OK;
}

}  // namespace maua
