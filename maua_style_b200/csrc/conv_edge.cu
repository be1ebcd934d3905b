// The image-side edge of the VGG stack: conv1_1 (Cin = 3) forward and its input-gradient.
//
// conv1_1 has K = 27: far too thin for a tensor-core tile and the layer is bound by writing (forward) or
// reading (backward) the 64-channel full-resolution map, so both directions are direct fp32 FFMA kernels
// with coalesced NHWC traffic.  The image itself stays NCHW [B,3,H,W] (the layout of the reference's
// pastiche, optim.py:173), which is not TMA-friendly for W % 4 != 0.
//
// The backward kernel is also the fused "image-side tail" (SURVEY.md A.4 item 5): it adds the TVLoss
// gradient (loss.py:224-233) and the temporal ContentLoss gradient (loss.py:46-54) while the image gradient
// is still in registers, so pastiche.grad is written exactly once.
#include "conv_tc.cuh"

namespace maua {

namespace {

// ------------------------------------------------------------------------------------------------
// forward: thread = 1 pixel x 16 output channels
// ------------------------------------------------------------------------------------------------
constexpr int FG_STRIDE = 20;  // 16 channels + 4 pad floats per group: conflict-free float4 broadcast reads

__global__ void __launch_bounds__(256)
conv_first_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ out, int B, int H, int W, int Cout, int do_round) {
    extern __shared__ float sm[];
    const int G = Cout / 16;
    float* ws = sm;                        // [27][G][FG_STRIDE]
    float* bs = sm + 27 * G * FG_STRIDE;   // [Cout]
    for (int i = threadIdx.x; i < 27 * Cout; i += blockDim.x) {
        const int co = i / 27, k = i % 27;
        ws[(k * G + co / 16) * FG_STRIDE + (co % 16)] = w[i];
    }
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) bs[i] = bias ? bias[i] : 0.f;
    __syncthreads();

    const int ppb = blockDim.x / G;
    const int g = threadIdx.x % G;
    const long HW = (long)H * W;
    const long npix = (long)B * HW;
    for (long pix = (long)blockIdx.x * ppb + threadIdx.x / G; pix < npix; pix += (long)gridDim.x * ppb) {
        const int b = pix / HW;
        const long r = pix - (long)b * HW;
        const int h = r / W, x = r % W;
        float xin[27];
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int hh = h + ky - 1, ww = x + kx - 1;
                    xin[ci * 9 + ky * 3 + kx] =
                        (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(img + ((long)b * 3 + ci) * HW + (long)hh * W + ww) : 0.f;
                }
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = bs[g * 16 + j];
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            const float4* wp = reinterpret_cast<const float4*>(ws + (k * G + g) * FG_STRIDE);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 wv = wp[j];
                acc[4 * j + 0] = fmaf(xin[k], wv.x, acc[4 * j + 0]);
                acc[4 * j + 1] = fmaf(xin[k], wv.y, acc[4 * j + 1]);
                acc[4 * j + 2] = fmaf(xin[k], wv.z, acc[4 * j + 2]);
                acc[4 * j + 3] = fmaf(xin[k], wv.w, acc[4 * j + 3]);
            }
        }
        float4* op = reinterpret_cast<float4*>(out + pix * Cout + g * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float4 o;
            o.x = fmaxf(acc[4 * j + 0], 0.f); o.y = fmaxf(acc[4 * j + 1], 0.f);
            o.z = fmaxf(acc[4 * j + 2], 0.f); o.w = fmaxf(acc[4 * j + 3], 0.f);
            if (do_round) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
            op[j] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward: block = 16 x 16 output pixels, 128 threads, thread = 2 vertically adjacent pixels x 3 channels.
// The 18 x 18 x Cout halo of the incoming gradient is staged in shared memory once (coalesced float4
// loads); pixel rows are padded by 4 floats so the per-thread float4 reads are bank-conflict free.
// ------------------------------------------------------------------------------------------------
constexpr int DT = 16;
constexpr int DH = DT + 2;

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) - (v < 0.f); }

__global__ void __launch_bounds__(128)
conv_first_dgrad_kernel(const float* __restrict__ gout, const float* __restrict__ w, float* __restrict__ gimg,
                        int B, int H, int W, int Cout, ImageTail tail) {
    extern __shared__ float sm[];
    const int PS = Cout + 4;               // padded pixel stride
    float* tile = sm;                      // [DH*DH][PS]
    float* wsd = sm + DH * DH * PS;        // [9][Cout][4]  (w for ci = 0,1,2 ; pad)
    for (int i = threadIdx.x; i < 9 * Cout; i += blockDim.x) {
        const int tap = i / Cout, co = i % Cout;
        const int ky = tap / 3, kx = tap % 3;
        float4 v;
        v.x = w[((long)co * 3 + 0) * 9 + ky * 3 + kx];
        v.y = w[((long)co * 3 + 1) * 9 + ky * 3 + kx];
        v.z = w[((long)co * 3 + 2) * 9 + ky * 3 + kx];
        v.w = 0.f;
        reinterpret_cast<float4*>(wsd)[i] = v;
    }
    const int tiles_w = (W + DT - 1) / DT, tiles_h = (H + DT - 1) / DT;
    const long ntiles = (long)B * tiles_w * tiles_h;
    const long HW = (long)H * W;
    const int C4 = Cout / 4;
    const float tvc = tail.tv_coef ? *tail.tv_coef : 0.f;
    const float tpc = tail.temp_coef ? *tail.temp_coef : 0.f;

    for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int tw = t % tiles_w;
        const int th = (t / tiles_w) % tiles_h;
        const int b = t / ((long)tiles_w * tiles_h);
        const int h0 = th * DT, w0 = tw * DT;
        __syncthreads();  // previous tile fully consumed (also orders the weight staging on the first pass)
        for (int i = threadIdx.x; i < DH * DH * C4; i += blockDim.x) {
            const int c4 = i % C4;
            const int pp = i / C4;
            const int hy = pp / DH, hx = pp % DH;
            const int hh = h0 + hy - 1, ww = w0 + hx - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (hh >= 0 && hh < H && ww >= 0 && ww < W)
                v = __ldg(reinterpret_cast<const float4*>(gout + (((long)b * H + hh) * W + ww) * Cout) + c4);
            *reinterpret_cast<float4*>(tile + pp * PS + c4 * 4) = v;
        }
        __syncthreads();

        const int tx = threadIdx.x % DT;
        const int ty2 = (threadIdx.x / DT) * 2;
        float acc[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                // gimg[h][w] += gout[h - ky + 1][w - kx + 1] * W[co][ci][ky][kx]; halo coords add +1
                const float* p0 = tile + ((ty2 + 2 - ky) * DH + (tx + 2 - kx)) * PS;
                const float* p1 = p0 + DH * PS;
                const float4* wp = reinterpret_cast<const float4*>(wsd) + (ky * 3 + kx) * Cout;
                for (int c = 0; c < Cout; c += 4) {
                    const float4 g0 = *reinterpret_cast<const float4*>(p0 + c);
                    const float4 g1 = *reinterpret_cast<const float4*>(p1 + c);
                    const float4 wa = wp[c], wb = wp[c + 1], wc = wp[c + 2], wd = wp[c + 3];
                    acc[0][0] += g0.x * wa.x + g0.y * wb.x + g0.z * wc.x + g0.w * wd.x;
                    acc[0][1] += g0.x * wa.y + g0.y * wb.y + g0.z * wc.y + g0.w * wd.y;
                    acc[0][2] += g0.x * wa.z + g0.y * wb.z + g0.z * wc.z + g0.w * wd.z;
                    acc[1][0] += g1.x * wa.x + g1.y * wb.x + g1.z * wc.x + g1.w * wd.x;
                    acc[1][1] += g1.x * wa.y + g1.y * wb.y + g1.z * wc.y + g1.w * wd.y;
                    acc[1][2] += g1.x * wa.z + g1.y * wb.z + g1.z * wc.z + g1.w * wd.z;
                }
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int h = h0 + ty2 + r, x = w0 + tx;
            if (h >= H || x >= W) continue;
            const float wt = (tail.temp_coef && tail.temp_weights) ? tail.temp_weights[(long)h * W + x] : 1.f;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const long idx = ((long)b * 3 + ci) * HW + (long)h * W + x;
                float v = acc[r][ci];
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
}

}  // namespace

int conv_first_fwd_launch(const float* img, const float* w, const float* bias, float* out, int B, int H, int W,
                          int Cout, int round, cudaStream_t st) {
    MAUA_REQUIRE(Cout % 16 == 0 && Cout <= 256 && 256 % (Cout / 16) == 0, "conv_first_fwd: unsupported Cout %d", Cout);
    const int G = Cout / 16;
    const size_t smem = (size_t)(27 * G * FG_STRIDE + Cout) * sizeof(float);
    const long npix = (long)B * H * W;
    const int ppb = 256 / G;
    long blocks = (npix + ppb - 1) / ppb;
    if (blocks > 148L * 8) blocks = 148L * 8;
    conv_first_fwd_kernel<<<(int)blocks, 256, smem, st>>>(img, w, bias, out, B, H, W, Cout, round);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

int conv_first_dgrad_launch(const float* gout, const float* w, float* gimg, int B, int H, int W, int Cout,
                            const ImageTail& tail, cudaStream_t st) {
    MAUA_REQUIRE(Cout % 4 == 0 && Cout <= 128, "conv_first_dgrad: unsupported Cout %d", Cout);
    const size_t smem = ((size_t)DH * DH * (Cout + 4) + (size_t)9 * Cout * 4) * sizeof(float);
    static unsigned long long attr_done = 0;
    MAUA_CUDA_CHECK(ensure_dynamic_smem(conv_first_dgrad_kernel, 200 * 1024, &attr_done));
    const long ntiles = (long)B * ((W + DT - 1) / DT) * ((H + DT - 1) / DT);
    long blocks = ntiles > 148L * 2 ? 148L * 2 : ntiles;
    conv_first_dgrad_kernel<<<(int)blocks, 128, smem, st>>>(gout, w, gimg, B, H, W, Cout, tail);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

}  // namespace maua
