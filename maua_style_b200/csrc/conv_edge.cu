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
        {
            // 3 x 6 x 66 halo values, 10 per thread: all loads are issued before the first store (an un-unrolled load -> store
            // loop serialises one global-load latency per element, which bounded this kernel in round 1)
            constexpr int kTot = 3 * (FT_H + 2) * (FT_W + 2);
            constexpr int kPer = (kTot + FT_THREADS - 1) / FT_THREADS;
            float tmp[kPer];
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int i = threadIdx.x + FT_THREADS * k;
                const int x = i % (FT_W + 2);
                const int r = (i / (FT_W + 2)) % (FT_H + 2);
                const int ci = i / ((FT_W + 2) * (FT_H + 2));
                const int hh = h0 + r - 1, ww = w0 + x - 1;
                tmp[k] = (i < kTot && hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(img + ((long)b * 3 + ci) * HW + (long)hh * W + ww) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int i = threadIdx.x + FT_THREADS * k;
                if (i < kTot) inp[i / ((FT_W + 2) * (FT_H + 2))][(i / (FT_W + 2)) % (FT_H + 2)][i % (FT_W + 2)] = tmp[k];
            }
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
// forward on the tensor core (the product path; the FFMA kernel above stays as the exact-arithmetic / cross-check path)
// ------------------------------------------------------------------------------------------------
// conv1_1 as a K = 27 -> 32 tcgen05 GEMM.  The FFMA kernel issues 864 FFMA2 + ~1300 other instructions per pixel quad and
// ran at 1.7 TB/s of output (ncu r02n: 44 % issue-slot utilisation, 168 registers, 18 % occupancy); here the 27 x 64
// multiply-adds of a pixel are one row of a 128 x 64 x 32 MMA and the CUDA cores only build the operand and run the epilogue.
//   A (im2col): a 128-pixel tile (16 wide x 8 high) gives 128 rows of 32 floats = exactly one 128-byte swizzle row each;
//      the producer warps read the staged image halo and write rows in the canonical K-major SWIZZLE_128B layout
//      (16-byte chunk c of row r at chunk c ^ (r % 8)), column k = ci*9 + ky*3 + kx, columns 27..31 zero.
//   B: the OIHW weights are already [64][27] row-major = K-major; padded to 32 columns, swizzled once per CTA.
//   Precision: the tensor core reads TF32 (11 significant bits) operands.  conv1_1 sees the raw image, and the optimizers
//      move it by less than 2^-20 of a pixel value per step (torch's L-BFGS starts with a step of 1 / |g|_1), so the image
//      is split EXACTLY into three TF32 numbers x = hi + mid + lo (11 + 11 + 2 bits) and the weights into hi + lo; the
//      product is accumulated as x_hi w_hi + x_mid w_hi + x_lo w_hi + x_hi w_lo: every bit of the image takes part, the
//      dropped terms are below 2^-22 relative -- the accuracy of the FFMA kernel for 16 instead of 4 MMAs per tile, still
//      far below the epilogue's time.  (With a two-way split of the image the 64 x 64 L-BFGS golden lost 43 dB: changes
//      of x below 2^-22 |x| were invisible to the first curvature pair.)
//   Warps: 0-3 producers (thread = pixel), 4 MMA issuer + TMEM owner, 5-8 epilogue (TMEM -> bias + ReLU + TF32 rounding ->
//      sign bitmap -> swizzled staging box -> TMA store, as in conv_tc.cu).  TMEM accumulators double-buffered (2 x 64
//      columns); two CTAs per SM hide each other's global-load and barrier latencies.
constexpr int C1_THREADS = 288;
constexpr int C1_A_BYTES = 128 * 128;   // one operand tile: 128 rows x 128 B
constexpr int C1_B_BYTES = 64 * 128;
constexpr int C1_IMG_W = 20;            // staged halo row: 18 used
constexpr int C1_SMEM = 3 * C1_A_BYTES + 2 * C1_B_BYTES + 2 * C1_A_BYTES /*staging boxes*/ + 3 * 10 * C1_IMG_W * 4 + 512 + 1024;

__global__ void __launch_bounds__(C1_THREADS, 2)
conv_first_tc_kernel(const __grid_constant__ CUtensorMap tmOut, const float* __restrict__ img, const float* __restrict__ w,
                     const float* __restrict__ bias, uint32_t* __restrict__ mask_out, int B, int H, int W, int do_round,
                     float tv_strength, float* __restrict__ tv_out, double* tv_partials, unsigned int* tv_counter) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                          // [hi | mid | lo]
    uint8_t* sB = sA + 3 * C1_A_BYTES;           // [hi | lo]
    uint8_t* sBox = sB + 2 * C1_B_BYTES;         // 2 staging boxes (1024-byte aligned: 3 x 16 KB + 2 x 8 KB before it)
    float* sImg = reinterpret_cast<float*>(sBox + 2 * C1_A_BYTES);  // [3][10][C1_IMG_W]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sImg + 3 * 10 * C1_IMG_W);
    uint64_t* a_full = bars;                     // producers -> MMA   (128 arrivals)
    uint64_t* a_empty = bars + 1;                // MMA commit -> producers
    uint64_t* tmem_full = bars + 2;              // [2] MMA commit -> epilogue
    uint64_t* tmem_empty = bars + 4;             // [2] epilogue (4 warps) -> MMA
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 6);
    float* tv_warp = reinterpret_cast<float*>(bars + 7);  // [4]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_w = (W + 15) / 16, tiles_h = (H + 7) / 8;
    const long ntiles = (long)B * tiles_w * tiles_h;
    const long HW = (long)H * W;

    // ---- prologue: barriers, TMEM, weights (frozen: may be read before the previous kernel has finished) ----
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmOut);
        mbar_init(a_full, 128);
        mbar_init(a_empty, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 4) { tmem_alloc(tmem_ptr_smem, 128); tmem_relinquish(); }
    for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {  // (row n, 16-byte chunk c) of B hi / lo
        const int n = i >> 3, c = i & 7;
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = 4 * c + e;
            // column 27 carries the bias: the A rows hold 1.0 there (hi tile only), so the tensor core adds it
            const float v = k < 27 ? __ldg(w + n * 27 + k) : (k == 27 && bias ? __ldg(bias + n) : 0.f);
            hi[e] = round_tf32(v);
            lo[e] = round_tf32(v - hi[e]);
        }
        const int off = n * 128 + ((c ^ (n & 7)) << 4);
        *reinterpret_cast<float4*>(sB + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(sB + C1_B_BYTES + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (threadIdx.x < 128) {  // K columns 28..31 of every A row stay zero for the whole kernel
        const int r = threadIdx.x;
        const int off = r * 128 + ((7 ^ (r & 7)) << 4);
        *reinterpret_cast<float4*>(sA + off) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(sA + C1_A_BYTES + off) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(sA + 2 * C1_A_BYTES + off) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();  // the image is written by the optimizer kernel of the previous iteration
    pdl_trigger();

    if (warp < 4) {
        // ===================== producers: stage the halo, build the im2col rows =====================
        const int r = threadIdx.x;          // pixel of the tile = A row
        const int hl = r >> 4, wl = r & 15;
        float tv_local = 0.f;
        uint32_t lt = 0;
        // The 3 x 10 x 18 halo of a tile is 540 values = up to 5 per producer thread.  They are fetched into registers one tile
        // ahead (all five loads in flight together, issued before the current tile's operand build), so the global-load
        // latency never sits on the per-tile critical path.
        constexpr int kPer = 5;
        float nxt[kPer];
        auto fetch = [&](long t, float (&dst)[kPer]) {
            const int tw = t % tiles_w;
            const int th = (t / tiles_w) % tiles_h;
            const int b = t / ((long)tiles_w * tiles_h);
            const int h0 = th * 8, w0 = tw * 16;
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int i = r + 128 * k;
                const int x = i % 18;
                const int y = (i / 18) % 10;
                const int ci = i / 180;
                const int hh = h0 + y - 1, ww = w0 + x - 1;
                dst[k] = (i < 540 && hh >= 0 && hh < H && ww >= 0 && ww < W)
                             ? __ldg(img + ((long)b * 3 + ci) * HW + (long)hh * W + ww) : 0.f;
            }
        };
        if ((long)blockIdx.x < ntiles) fetch(blockIdx.x, nxt);
        for (long t = blockIdx.x; t < ntiles; t += gridDim.x, ++lt) {
            const int tw = t % tiles_w;
            const int th = (t / tiles_w) % tiles_h;
            const int h0 = th * 8, w0 = tw * 16;
            named_bar_sync(3, 128);  // every producer has finished reading the previous tile's halo
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int i = r + 128 * k;
                if (i < 540) sImg[((i / 180) * 10 + (i / 18) % 10) * C1_IMG_W + i % 18] = nxt[k];
            }
            if (t + gridDim.x < ntiles) fetch(t + gridDim.x, nxt);  // in flight during this tile's build
            named_bar_sync(3, 128);
            float v[28];
#pragma unroll
            for (int ci = 0; ci < 3; ++ci)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) v[ci * 9 + ky * 3 + kx] = sImg[(ci * 10 + hl + ky) * C1_IMG_W + wl + kx];
            v[27] = 1.f;  // x 1.0 = hi 1.0 + mid 0 + lo 0: multiplies the bias column of B
            if (tv_out) {  // TVLoss value (loss.py:229-233) from the staged neighbours; row / column 0 have none
                const int h = h0 + hl, x = w0 + wl;
                if (h < H && x < W) {
#pragma unroll
                    for (int ci = 0; ci < 3; ++ci) {
                        const float c0 = v[ci * 9 + 4];
                        if (h > 0) tv_local += fabsf(c0 - v[ci * 9 + 1]);
                        if (x > 0) tv_local += fabsf(c0 - v[ci * 9 + 3]);
                    }
                }
            }
            mbar_wait(a_empty, (lt & 1) ^ 1);  // the MMAs of the previous tile have read the operand tile
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                float hi[4], mid[4], lo[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    hi[e] = round_tf32(v[4 * c + e]);
                    const float r1 = v[4 * c + e] - hi[e];  // exact
                    mid[e] = round_tf32(r1);
                    lo[e] = r1 - mid[e];                    // exact, <= 3 significant bits: a TF32 number
                }
                const int off = r * 128 + ((c ^ (r & 7)) << 4);
                *reinterpret_cast<float4*>(sA + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(sA + C1_A_BYTES + off) = make_float4(mid[0], mid[1], mid[2], mid[3]);
                *reinterpret_cast<float4*>(sA + 2 * C1_A_BYTES + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
            mbar_arrive(a_full);
        }
        if (tv_out) {
            tv_local = warp_sum(tv_local);
            if (lane == 0) tv_warp[warp] = tv_local;
        }
    } else if (warp == 4) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_tf32(128, 64, 0, 0);
        uint32_t lt = 0;
        for (long t = blockIdx.x; t < ntiles; t += gridDim.x, ++lt) {
            const int acc = lt & 1;
            mbar_wait(&tmem_empty[acc], ((lt >> 1) & 1) ^ 1);
            mbar_wait(a_full, lt & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tmem_base + acc * 64;
                const uint32_t a_hi = smem_u32(sA), a_mid = a_hi + C1_A_BYTES, a_lo = a_hi + 2 * C1_A_BYTES;
                const uint32_t b_hi = smem_u32(sB), b_lo = b_hi + C1_B_BYTES;
                // smallest terms first: the TMEM accumulator adds in fp32
                const uint32_t aa[4] = {a_lo, a_hi, a_mid, a_hi};
                const uint32_t bb[4] = {b_hi, b_lo, b_hi, b_hi};
#pragma unroll
                for (int term = 0; term < 4; ++term) {
                    const uint64_t adesc = make_smem_desc_sw128(aa[term], 16, 1024);
                    const uint64_t bdesc = make_smem_desc_sw128(bb[term], 16, 1024);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_tf32(d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (term | kk) ? 1u : 0u);
                }
                umma_commit(a_empty);
                umma_commit(&tmem_full[acc]);
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue =====================
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;
        const int hl = row >> 4, wl = row & 15;
        const bool issuer = (threadIdx.x == 160);
        uint32_t lt = 0, box = 0;
        for (long t = blockIdx.x; t < ntiles; t += gridDim.x, ++lt) {
            const int tw = t % tiles_w;
            const int th = (t / tiles_w) % tiles_h;
            const int b = t / ((long)tiles_w * tiles_h);
            const int h0 = th * 8, w0 = tw * 16;
            const int h = h0 + hl, x = w0 + wl;
            const bool valid = (h < H) && (x < W);
            const size_t pix = (static_cast<size_t>(b) * H + h) * W + x;
            const int acc = lt & 1;
            mbar_wait(&tmem_full[acc], (lt >> 1) & 1);
            tc_fence_after();
            const uint32_t t_base = tmem_base + acc * 64 + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < 64; c += 32, ++box) {
                float v[32];
                tmem_ld_x32(t_base + c, v);
                uint32_t bits = 0;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float o = fmaxf(v[i], 0.f);
                    if (do_round) o = round_tf32(o);
                    v[i] = o;
                    bits |= (o > 0.f ? 1u : 0u) << i;
                }
                if (mask_out && valid) mask_out[pix * 2 + (c >> 5)] = bits;
                // Two staging boxes alternate.  Box (box & 1) was last read by the store issued two boxes ago; the issuer waited
                // for that store's reads before arriving at the previous box's barrier, so it is free for everybody here.
                uint8_t* sbox = sBox + (box & 1) * C1_A_BYTES;
                uint8_t* srow = sbox + row * 128;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    *reinterpret_cast<float4*>(srow + ((k ^ (row & 7)) << 4)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                fence_proxy_async_smem();
                if (issuer) bulk_wait_group_read<0>();  // the previous box's store has drained its staging box (the next one to be written)
                named_bar_sync(1, 128);
                if (issuer) {
                    tma_store_4d(&tmOut, sbox, c, w0, h0, b);
                    bulk_commit_group();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
        if (issuer) bulk_wait_group<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 128);
    if (tv_out && threadIdx.x == 0) {
        // deterministic grid sum: per-CTA partials, the last CTA adds them in index order
        tv_partials[blockIdx.x] = (double)tv_warp[0] + (double)tv_warp[1] + (double)tv_warp[2] + (double)tv_warp[3];
        __threadfence();
        if (atomicAdd(tv_counter, 1u) == gridDim.x - 1) {
            __threadfence();
            double s = 0.0;
            for (unsigned int i = 0; i < gridDim.x; ++i) s += __ldcg(tv_partials + i);
            *tv_out = tv_strength * (float)s;
            *tv_counter = 0;
        }
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
                          int H, int W, int Cout, int round, cudaStream_t st, const ConvFirstTV* tv, int exact) {
    MAUA_REQUIRE(Cout == 64, "conv_first_fwd: the image layer must have 64 output channels (got %d)", Cout);
    const char* tc = getenv("MAUA_CONV1_TC");  // tensor-core kernel by default; 0: the FFMA kernel
    if (!exact && (!tc || atoi(tc) != 0)) {
        static unsigned long long attr_done = 0;
        MAUA_CUDA_CHECK((ensure_dynamic_smem(conv_first_tc_kernel, C1_SMEM, &attr_done)));
        CUtensorMap tmOut;
        int rc = make_tmap_nhwc(&tmOut, out, B, H, W, Cout, 16, 8);
        if (rc) return rc;
        const long nt = (long)B * ((W + 15) / 16) * ((H + 7) / 8);
        const long grid = nt > 148L * 2 ? 148L * 2 : nt;
        MAUA_REQUIRE(!tv || (tv->out && tv->partials && tv->counter && grid <= tv->max_blocks), "conv_first_fwd: bad TVLoss arguments");
        MAUA_CUDA_CHECK(launch_pdl<PDL_EDGE>(conv_first_tc_kernel, dim3((unsigned)grid), dim3(C1_THREADS), C1_SMEM, st, tmOut, img, w, bias,
                                             mask_out, B, H, W, round, tv ? tv->strength : 0.f, tv ? tv->out : (float*)nullptr,
                                             tv ? tv->partials : (double*)nullptr, tv ? tv->counter : (unsigned int*)nullptr));
        return MAUA_OK;
    }
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
