// Memory-bound kernels of the maua-style hot path: layout conversion, weight preparation,
// 2x2 pooling (forward / backward), TV loss, content MSE, small reductions and the Adam update.
// All are vectorised, coalesced, grid-stride kernels sized to a multiple of the SM count, with
// warp-shuffle reductions and a deterministic last-block final reduce (no float atomics).
#include "pointwise.cuh"
#include "conv_tc.cuh"
#include "reduce.cuh"

namespace maua {

namespace {

constexpr int kThreads = 256;
inline int grid_for(long n_items, int per_sm = 8) {
    long blocks = (n_items + kThreads - 1) / kThreads;
    long cap = 148L * per_sm;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

template <int NV, class Fin>
__device__ __forceinline__ void block_reduce_finish(double (&acc)[NV], double* partials, unsigned int* counter, Fin fin) {
    double tot[NV];
    if (grid_sum<NV>(acc, partials, counter, tot)) fin(tot);
}

// --------------------------------------------------------------------------------------------
// layout conversion
// --------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, long HW,
                                    int do_round) {
    // tile transpose through shared memory: 32 pixels x 32 channels
    __shared__ float tile[32][33];
    const long ptiles = (HW + 31) / 32;
    const int ctiles = (C + 31) / 32;
    const long ntiles = ptiles * ctiles * B;
    for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long pt = t % ptiles;
        const int ct = (t / ptiles) % ctiles;
        const int b = t / (ptiles * ctiles);
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
        for (int j = ty; j < 32; j += 8) {
            const int c = ct * 32 + j;
            const long p = pt * 32 + tx;
            tile[j][tx] = (c < C && p < HW) ? src[((long)b * C + c) * HW + p] : 0.f;
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            const long p = pt * 32 + j;
            const int c = ct * 32 + tx;
            if (c < C && p < HW) {
                float v = tile[tx][j];
                dst[((long)b * HW + p) * C + c] = do_round ? round_tf32(v) : v;
            }
        }
        __syncthreads();
    }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, long HW) {
    __shared__ float tile[32][33];
    const long ptiles = (HW + 31) / 32;
    const int ctiles = (C + 31) / 32;
    const long ntiles = ptiles * ctiles * B;
    for (long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long pt = t % ptiles;
        const int ct = (t / ptiles) % ctiles;
        const int b = t / (ptiles * ctiles);
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        for (int j = ty; j < 32; j += 8) {
            const long p = pt * 32 + j;
            const int c = ct * 32 + tx;
            tile[j][tx] = (c < C && p < HW) ? src[((long)b * HW + p) * C + c] : 0.f;
        }
        __syncthreads();
        for (int j = ty; j < 32; j += 8) {
            const int c = ct * 32 + j;
            const long p = pt * 32 + tx;
            if (c < C && p < HW) dst[((long)b * C + c) * HW + p] = tile[tx][j];
        }
        __syncthreads();
    }
}

// sign bitmap of an NHWC activation (one warp per 32-channel word group: lane = channel, ballot = word)
__global__ void relu_mask_bits_kernel(const float* __restrict__ x, uint32_t* __restrict__ bits, long nwords) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long wd = warp0; wd < nwords; wd += nwarps) {
        const uint32_t b = __ballot_sync(0xffffffffu, x[wd * 32 + lane] > 0.f);
        if (lane == 0) bits[wd] = b;
    }
}

// --------------------------------------------------------------------------------------------
// weight preparation (once per model load)
// --------------------------------------------------------------------------------------------
// fwd : wg[co][tap*Cin + ci]  = tf32(w[co][ci][ky][kx]),           tap = ky*3 + kx
// dgrad: wd[ci][tap*Cout + co] = tf32(w[co][ci][2-ky][2-kx])        (180-degree rotation, transposed)
__global__ void prep_weights_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin,
                                    int dgrad, int do_round, int T /* taps: 9 (3x3) or 1 (1x1) */) {
    const long total = (long)Cout * Cin * T;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        float v;
        if (!dgrad) {
            const int ci = i % Cin;
            const int tap = (i / Cin) % T;
            const int co = i / ((long)T * Cin);
            v = w[((long)co * Cin + ci) * T + tap];
        } else {
            const int co = i % Cout;
            const int tap = (i / Cout) % T;
            const int ci = i / ((long)T * Cout);
            v = w[((long)co * Cin + ci) * T + (T - 1 - tap)];
        }
        out[i] = do_round ? round_tf32(v) : v;
    }
}

// --------------------------------------------------------------------------------------------
// 2x2 stride-2 pooling on NHWC (models.py:119-122); floor semantics drop an odd last row/col.
// --------------------------------------------------------------------------------------------
__global__ void pool_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, int B, int H, int W, int C4,
                                int avg, int do_round, uint8_t* __restrict__ codes) {
    pdl_wait();
    pdl_trigger();
    const int PH = H / 2, PW = W / 2;
    const long total = (long)B * PH * PW * C4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = i % C4;
        long r = i / C4;
        const int pw = r % PW;
        r /= PW;
        const int ph = r % PH;
        const int b = r / PH;
        const long base = (((long)b * H + 2 * ph) * W + 2 * pw) * C4 + c;
        const float4 a0 = x[base], a1 = x[base + C4], a2 = x[base + (long)W * C4], a3 = x[base + (long)W * C4 + C4];
        float4 o;
        if (avg) {
            o.x = 0.25f * (a0.x + a1.x + a2.x + a3.x); o.y = 0.25f * (a0.y + a1.y + a2.y + a3.y);
            o.z = 0.25f * (a0.z + a1.z + a2.z + a3.z); o.w = 0.25f * (a0.w + a1.w + a2.w + a3.w);
            if (do_round) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        } else {
            o.x = fmaxf(fmaxf(a0.x, a1.x), fmaxf(a2.x, a3.x)); o.y = fmaxf(fmaxf(a0.y, a1.y), fmaxf(a2.y, a3.y));
            o.z = fmaxf(fmaxf(a0.z, a1.z), fmaxf(a2.z, a3.z)); o.w = fmaxf(fmaxf(a0.w, a1.w), fmaxf(a2.w, a3.w));
            if (codes)
                codes[i] = static_cast<uint8_t>(first_argmax(a0.x, a1.x, a2.x, a3.x) | (first_argmax(a0.y, a1.y, a2.y, a3.y) << 2) |
                                                (first_argmax(a0.z, a1.z, a2.z, a3.z) << 4) | (first_argmax(a0.w, a1.w, a2.w, a3.w) << 6));
        }
        y[i] = o;
    }
}


// gx[h][w][c] = (argmax of window == (h,w) ? gy[h/2][w/2][c] : 0) * (x > 0)  [+ addend]; the ReLU mask of
// the producing layer is folded in (x is the post-ReLU pre-pool activation).  Pixels in a dropped odd
// row / column receive zero (+ addend).
__global__ void pool_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ gy,
                                const float4* __restrict__ addend, float4* __restrict__ gx, int B, int H, int W,
                                int C4, int avg, int do_round) {
    pdl_wait();
    pdl_trigger();
    const int PH = H / 2, PW = W / 2;
    const int WH = (H + 1) / 2, WW = (W + 1) / 2;  // windows incl. partial ones
    const long total = (long)B * WH * WW * C4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = i % C4;
        long r = i / C4;
        const int pw = r % WW;
        r /= WW;
        const int ph = r % WH;
        const int b = r / WH;
        const bool full = (ph < PH) && (pw < PW);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (full) g = gy[(((long)b * PH + ph) * PW + pw) * C4 + c];
        float4 a[4];
        long offs[4];
        bool ok[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int h = 2 * ph + (k >> 1), w = 2 * pw + (k & 1);
            ok[k] = (h < H) && (w < W);
            offs[k] = (((long)b * H + h) * W + w) * C4 + c;
            a[k] = ok[k] ? x[offs[k]] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 o[4];
        if (avg) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                o[k].x = a[k].x > 0.f ? 0.25f * g.x : 0.f; o[k].y = a[k].y > 0.f ? 0.25f * g.y : 0.f;
                o[k].z = a[k].z > 0.f ? 0.25f * g.z : 0.f; o[k].w = a[k].w > 0.f ? 0.25f * g.w : 0.f;
            }
        } else {
            const int kx = first_argmax(a[0].x, a[1].x, a[2].x, a[3].x);
            const int ky = first_argmax(a[0].y, a[1].y, a[2].y, a[3].y);
            const int kz = first_argmax(a[0].z, a[1].z, a[2].z, a[3].z);
            const int kw = first_argmax(a[0].w, a[1].w, a[2].w, a[3].w);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                o[k].x = (k == kx && a[k].x > 0.f) ? g.x : 0.f; o[k].y = (k == ky && a[k].y > 0.f) ? g.y : 0.f;
                o[k].z = (k == kz && a[k].z > 0.f) ? g.z : 0.f; o[k].w = (k == kw && a[k].w > 0.f) ? g.w : 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!ok[k]) continue;
            float4 v = o[k];
            if (addend) {
                // gradient from a loss tapped at this (pre-pool) layer; masked by the same ReLU
                const float4 ad = addend[offs[k]];
                v.x += a[k].x > 0.f ? ad.x : 0.f; v.y += a[k].y > 0.f ? ad.y : 0.f;
                v.z += a[k].z > 0.f ? ad.z : 0.f; v.w += a[k].w > 0.f ? ad.w : 0.f;
            }
            if (do_round) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
            gx[offs[k]] = v;
        }
    }
}

// The same un-pooling without the pre-pool activation: the window's winner comes from the arg-max codes the forward pass wrote
// (one byte per 4 channels of a pooled pixel), x > 0 from the sign bitmap of the producing conv.  Same decisions, same
// arithmetic as pool_bwd_kernel -- bit-identical output -- for 5.2 instead of 9 bytes of traffic per element.
__global__ void pool_bwd_codes_kernel(const uint32_t* __restrict__ bits, const uint8_t* __restrict__ codes,
                                      const float4* __restrict__ gy, const float4* __restrict__ addend, float4* __restrict__ gx,
                                      int B, int H, int W, int C4, int avg, int do_round) {
    pdl_wait();
    pdl_trigger();
    const int PH = H / 2, PW = W / 2;
    const int WH = (H + 1) / 2, WW = (W + 1) / 2;  // windows incl. partial ones
    const int words = C4 >> 3;                     // 32-channel words of the bitmap per pixel
    const long total = (long)B * WH * WW * C4;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = i % C4;
        long r = i / C4;
        const int pw = r % WW;
        r /= WW;
        const int ph = r % WH;
        const int b = r / WH;
        const bool full = (ph < PH) && (pw < PW);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned int code = 0;
        if (full) {
            const long pi = (((long)b * PH + ph) * PW + pw) * C4 + c;
            g = gy[pi];
            if (!avg) code = codes[pi];
        }
        const int kx = code & 3, ky = (code >> 2) & 3, kz = (code >> 4) & 3, kw = (code >> 6) & 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int h = 2 * ph + (k >> 1), w = 2 * pw + (k & 1);
            if (h >= H || w >= W) continue;
            const long pix = ((long)b * H + h) * W + w;
            const long off = pix * C4 + c;
            const unsigned int s4 = (__ldg(bits + pix * words + (c >> 3)) >> ((c & 7) * 4)) & 0xFu;  // x > 0 of this thread's 4 channels
            const bool px = s4 & 1u, py = s4 & 2u, pz = s4 & 4u, pq = s4 & 8u;
            float4 v;
            if (avg) {
                v.x = px ? 0.25f * g.x : 0.f; v.y = py ? 0.25f * g.y : 0.f; v.z = pz ? 0.25f * g.z : 0.f; v.w = pq ? 0.25f * g.w : 0.f;
            } else {
                v.x = (k == kx && px) ? g.x : 0.f; v.y = (k == ky && py) ? g.y : 0.f;
                v.z = (k == kz && pz) ? g.z : 0.f; v.w = (k == kw && pq) ? g.w : 0.f;
            }
            if (addend) {
                const float4 ad = addend[off];
                v.x += px ? ad.x : 0.f; v.y += py ? ad.y : 0.f; v.z += pz ? ad.z : 0.f; v.w += pq ? ad.w : 0.f;
            }
            if (do_round) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
            gx[off] = v;
        }
    }
}

// --------------------------------------------------------------------------------------------
// TVLoss value (loss.py:224-233): strength * (sum|x[h]-x[h-1]| + sum|x[w]-x[w-1]|) on NCHW.
// The gradient is fused into the conv1_1 dgrad tail (conv_edge.cu).
// --------------------------------------------------------------------------------------------
__global__ void tv_value_kernel(const float* __restrict__ x, int planes, int H, int W, float strength,
                                float* __restrict__ loss_out, double* partials, unsigned int* counter) {
    pdl_wait();
    pdl_trigger();
    const long total = (long)planes * H * W;
    double acc[1] = {0.0};
    float local = 0.f;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int w = i % W;
        const int h = (i / W) % H;
        const float v = x[i];
        if (h > 0) local += fabsf(v - x[i - W]);
        if (w > 0) local += fabsf(v - x[i - 1]);
    }
    acc[0] = local;
    block_reduce_finish<1>(acc, partials, counter,
                           [=](double (&tot)[1]) { *loss_out = strength * (float)tot[0]; });
}

// --------------------------------------------------------------------------------------------
// ContentLoss value (loss.py:53-59): strength * mean((x * w? - target)^2).
// weights (optional) broadcast over channels: index = pixel (NHWC) or position in plane (NCHW).
// --------------------------------------------------------------------------------------------
__global__ void mse_value_kernel(const float4* __restrict__ x, const float4* __restrict__ t, long n4, float scale,
                                 float* __restrict__ loss_out, double* partials, unsigned int* counter) {
    pdl_wait();
    pdl_trigger();
    double acc[1] = {0.0};
    float local = 0.f;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 a = x[i], b = t[i];
        const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
        local += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        if ((i & 63) == 63) { acc[0] += local; local = 0.f; }
    }
    acc[0] += local;
    block_reduce_finish<1>(acc, partials, counter, [=](double (&tot)[1]) { *loss_out = scale * (float)tot[0]; });
}

__global__ void wmse_value_kernel(const float* __restrict__ x, const float* __restrict__ wts,
                                  const float* __restrict__ t, long n, long plane, float scale,
                                  float* __restrict__ loss_out, double* partials, unsigned int* counter) {
    pdl_wait();
    pdl_trigger();
    double acc[1] = {0.0};
    float local = 0.f;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float wv = wts ? wts[i % plane] : 1.f;
        const float d = x[i] * wv - t[i];
        local += d * d;
    }
    acc[0] = local;
    block_reduce_finish<1>(acc, partials, counter, [=](double (&tot)[1]) { *loss_out = scale * (float)tot[0]; });
}

// per-channel sums over pixels of an NHWC tensor (covariance mean, loss.py:87-89)
__global__ void channel_sum_kernel(const float* __restrict__ x, long P, int C, double* __restrict__ partial /*[grid][C]*/) {
    // block handles a contiguous pixel range; consecutive threads read consecutive channels (coalesced)
    const long per = (P + gridDim.x - 1) / gridDim.x;
    const long p0 = blockIdx.x * per, p1 = min(P, p0 + per);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double s = 0.0;
        float run = 0.f;
        int cnt = 0;
        for (long p = p0; p < p1; ++p) {
            run += x[p * C + c];
            if (++cnt == 64) { s += run; run = 0.f; cnt = 0; }
        }
        s += run;
        partial[(size_t)blockIdx.x * C + c] = s;
    }
}
__global__ void channel_sum_final_kernel(const double* __restrict__ partial, int nblocks, int C, long P,
                                         float* __restrict__ mean_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * C + c];
    mean_out[c] = (float)(s / (double)P);
}

// --------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam single-tensor math, reference optim.py:192-196): 28 bytes / element.
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g*g ; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// --------------------------------------------------------------------------------------------
// step_dev (optional): the step number lives on the device (graph-captured loops replay the same launch), and the
// bias corrections are derived from it here instead of on the host.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long n, float lr, float b1, float b2, float eps, float bc1,
                            float bc2_sqrt, const int* __restrict__ step_dev) {
    pdl_wait();
    pdl_trigger();
    const long n4 = n / 4;
    if (step_dev) {
        const int t = *step_dev;
        bc1 = (float)(1.0 - pow((double)b1, (double)t));
        bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, (double)t));
    }
    const float step = lr / bc1;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
#define MAUA_ADAM1(P, G, M, V)                                   \
    M = M + (G - M) * (1.f - b1);                                \
    V = V * b2 + (1.f - b2) * G * G;                             \
    P = P - step * (M / (sqrtf(V) / bc2_sqrt + eps));
        MAUA_ADAM1(pp.x, gg.x, mm.x, vv.x)
        MAUA_ADAM1(pp.y, gg.y, mm.y, vv.y)
        MAUA_ADAM1(pp.z, gg.z, mm.z, vv.z)
        MAUA_ADAM1(pp.w, gg.w, mm.w, vv.w)
        p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
    // tail
    for (long i = n4 * 4 + blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float P = p[i], G = g[i], M = m[i], V = v[i];
        MAUA_ADAM1(P, G, M, V)
        p[i] = P; m[i] = M; v[i] = V;
    }
#undef MAUA_ADAM1
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int nchw_to_nhwc_launch(const float* src, float* dst, int B, int C, int H, int W, int do_round, cudaStream_t st) {
    const long HW = (long)H * W;
    const long ntiles = ((HW + 31) / 32) * ((C + 31) / 32) * B;
    nchw_to_nhwc_kernel<<<(int)(ntiles > 148 * 16 ? 148 * 16 : ntiles), 256, 0, st>>>(src, dst, B, C, HW, do_round);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}
int nhwc_to_nchw_launch(const float* src, float* dst, int B, int C, int H, int W, cudaStream_t st) {
    const long HW = (long)H * W;
    const long ntiles = ((HW + 31) / 32) * ((C + 31) / 32) * B;
    nhwc_to_nchw_kernel<<<(int)(ntiles > 148 * 16 ? 148 * 16 : ntiles), 256, 0, st>>>(src, dst, B, C, HW);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}
int relu_mask_bits_launch(const float* x, uint32_t* bits, long npix, int C, cudaStream_t st) {
    MAUA_REQUIRE(C % 32 == 0, "mask bitmap: C %% 32 != 0");
    const long nwords = npix * (C / 32);
    relu_mask_bits_kernel<<<grid_for(nwords * 32, 8), kThreads, 0, st>>>(x, bits, nwords);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}
int prep_weights_launch(const float* w, float* out, int Cout, int Cin, int dgrad, int do_round, cudaStream_t st, int taps) {
    prep_weights_kernel<<<grid_for((long)Cout * Cin * taps), kThreads, 0, st>>>(w, out, Cout, Cin, dgrad, do_round, taps);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}
int pool_fwd_launch(const float* x, float* y, int B, int H, int W, int C, int avg, int do_round, cudaStream_t st, uint8_t* codes) {
    MAUA_REQUIRE(C % 4 == 0, "pool: C %% 4 != 0");
    if (H / 2 == 0 || W / 2 == 0) return MAUA_OK;
    const long total = (long)B * (H / 2) * (W / 2) * (C / 4);
    MAUA_CUDA_CHECK(launch_pdl(pool_fwd_kernel, dim3(grid_for(total, 16)), dim3(kThreads), 0, st, reinterpret_cast<const float4*>(x),
                               reinterpret_cast<float4*>(y), B, H, W, C / 4, avg, do_round, avg ? (uint8_t*)nullptr : codes));
    return MAUA_OK;
}
int pool_bwd_codes_launch(const uint32_t* bits, const uint8_t* codes, const float* gy, const float* addend, float* gx, int B,
                          int H, int W, int C, int avg, int do_round, cudaStream_t st) {
    MAUA_REQUIRE(C % 32 == 0 && bits && (avg || codes), "pool_bwd_codes: needs C %% 32 == 0, the sign bitmap and (max pooling) the codes");
    const long total = (long)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
    MAUA_CUDA_CHECK(launch_pdl(pool_bwd_codes_kernel, dim3(grid_for(total, 16)), dim3(kThreads), 0, st, bits, codes,
                               reinterpret_cast<const float4*>(gy), reinterpret_cast<const float4*>(addend),
                               reinterpret_cast<float4*>(gx), B, H, W, C / 4, avg, do_round));
    return MAUA_OK;
}
int pool_bwd_launch(const float* x, const float* gy, const float* addend, float* gx, int B, int H, int W, int C,
                    int avg, int do_round, cudaStream_t st) {
    MAUA_REQUIRE(C % 4 == 0, "pool: C %% 4 != 0");
    const long total = (long)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
    MAUA_CUDA_CHECK(launch_pdl(pool_bwd_kernel, dim3(grid_for(total, 16)), dim3(kThreads), 0, st, reinterpret_cast<const float4*>(x),
                               reinterpret_cast<const float4*>(gy), reinterpret_cast<const float4*>(addend),
                               reinterpret_cast<float4*>(gx), B, H, W, C / 4, avg, do_round));
    return MAUA_OK;
}
int tv_value_launch(const float* x, int planes, int H, int W, float strength, float* loss_out, ReduceScratch rs,
                    cudaStream_t st) {
    const int grid = grid_for((long)planes * H * W, 4);
    MAUA_REQUIRE(grid <= rs.max_blocks, "reduce scratch too small");
    MAUA_CUDA_CHECK(launch_pdl(tv_value_kernel, dim3(grid), dim3(kThreads), 0, st, x, planes, H, W, strength, loss_out, rs.partials,
                               rs.counter));
    return MAUA_OK;
}
int mse_value_launch(const float* x, const float* t, long n, float scale, float* loss_out, ReduceScratch rs,
                     cudaStream_t st) {
    MAUA_REQUIRE(n % 4 == 0, "mse: n %% 4 != 0");
    const int grid = grid_for(n / 4, 4);
    MAUA_REQUIRE(grid <= rs.max_blocks, "reduce scratch too small");
    MAUA_CUDA_CHECK(launch_pdl(mse_value_kernel, dim3(grid), dim3(kThreads), 0, st, reinterpret_cast<const float4*>(x),
                               reinterpret_cast<const float4*>(t), n / 4, scale, loss_out, rs.partials, rs.counter));
    return MAUA_OK;
}
int wmse_value_launch(const float* x, const float* wts, const float* t, long n, long plane, float scale,
                      float* loss_out, ReduceScratch rs, cudaStream_t st) {
    const int grid = grid_for(n, 4);
    MAUA_REQUIRE(grid <= rs.max_blocks, "reduce scratch too small");
    MAUA_CUDA_CHECK(launch_pdl(wmse_value_kernel, dim3(grid), dim3(kThreads), 0, st, x, wts, t, n, plane, scale, loss_out, rs.partials,
                               rs.counter));
    return MAUA_OK;
}
int channel_mean_launch(const float* x, long P, int C, float* mean_out, double* scratch, int scratch_blocks,
                        cudaStream_t st) {
    int blocks = (int)((P + 255) / 256);
    if (blocks > scratch_blocks) blocks = scratch_blocks;
    if (blocks < 1) blocks = 1;
    channel_sum_kernel<<<blocks, C < 256 ? (C < 32 ? 32 : C) : 256, 0, st>>>(x, P, C, scratch);
    MAUA_CUDA_CHECK(cudaGetLastError());
    channel_sum_final_kernel<<<(C + 127) / 128, 128, 0, st>>>(scratch, blocks, C, P, mean_out);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}
int adam_launch(float* p, const float* g, float* m, float* v, long n, float lr, float b1, float b2, float eps,
                int step, const int* step_dev, cudaStream_t st) {
    MAUA_REQUIRE(step >= 1 || step_dev, "adam step must be >= 1");
    if (step < 1) step = 1;
    MAUA_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                   reinterpret_cast<uintptr_t>(v)) & 15) == 0, "adam: pointers must be 16-byte aligned");
    const double bc1 = 1.0 - pow((double)b1, step);
    const double bc2 = 1.0 - pow((double)b2, step);
    MAUA_CUDA_CHECK(launch_pdl(adam_kernel, dim3(grid_for(n / 4 + 1, 8)), dim3(kThreads), 0, st, p, g, m, v, n, lr, b1, b2, eps,
                               (float)bc1, (float)sqrt(bc2), step_dev));
    return MAUA_OK;
}

}  // namespace maua
