// L-BFGS pixel update, fully device-resident: three launches per iteration, two streaming passes over the history.
//
// Replaces torch.optim.LBFGS as the reference drives it (optim.py:180-191: max_iter = num_iters,
// tolerance_grad = tolerance_change = -1, history 100, lr 1, NO line search), i.e. per iteration:
//     y = g - g_prev ; s = t d ; if y.s > 1e-10: push (s, y), ro = 1/y.s, H = y.s / y.y
//     q = -g ; for i newest..oldest: al_i = ro_i s_i.q ; q -= al_i y_i
//     r = H q ; for i oldest..newest: be_i = ro_i y_i.r ; r += (al_i - be_i) s_i ; d = r
//     t = (first iteration) ? min(1, 1/|g|_1) lr : lr ;  if g.d > -tolerance_change: stop ;  x += t d
//
// The two-loop recursion is 2k dependent (dot, axpy) steps over n-element vectors: the reference pays >= 4k+10 tiny
// launches and a host sync per dot.  Here the SAME recurrences are evaluated on scalars:  with the inner products
//     SG_i = s_i.g   YG_i = y_i.g   SY_ij = s_i.y_j   YY_ij = y_i.y_j   GG = g.g
// the loop-1 dots are  s_i.q_i = -SG_i - sum_{j>i} al_j SY_ij  and the loop-2 dots are
//     y_i.r_i = H (-YG_i - sum_j al_j YY_ij) + sum_{j<i} (al_j - be_j) SY_ji,
// so al / be follow from O(k^2) scalar work (one CTA, double precision) and the direction is one linear combination
//     d = -H g - H sum_j al_j y_j + sum_j (al_j - be_j) s_j.
// SY / YY are kept on the device and extended by one row/column per accepted pair.  Per iteration:
//   1. lbfgs_dots_kernel   : forms y, s into the ring, streams every history vector ONCE and takes all its dot products
//                            with g, y_new, s_new at the same time (multi-dot), per-CTA partials;
//   2. lbfgs_scalar_kernel : reduces the partials in a fixed order (deterministic), applies the y.s gate, updates
//                            SY / YY / ro / H, runs both recurrences, computes g.d, t and the halt flag;
//   3. lbfgs_update_kernel : streams the history a second time:  d = combination,  x += t d.
// HBM traffic = 2 reads of the 2k history vectors = 16 k n bytes per iteration, the same as the
// two-loop form, but as two bandwidth-bound sweeps instead of 2k latency-bound ones.
#include "maua_b200.h"
#include "pointwise.cuh"

namespace maua {

namespace {

constexpr int kMaxHist = 256;
constexpr int kRing = kMaxHist + 1;
constexpr int kThreads = 256;
constexpr int kChunk = 2048;  // floats of the vector handled by one CTA in the multi-dot pass (2 float4 per thread)
constexpr int kNV = 5;        // dot products per history slot

struct LbfgsState {  // device resident scalars
    int n_iter, hist_len, head, halted;
    int cand, accepted, pad0, pad1;
    float t, H_diag, lr, tol_change;
    float gtd, pad2, pad3, pad4;
    float cg;                 // coefficient of g in d
    float cy[kRing];          // coefficient of y_slot in d
    float cs[kRing];          // coefficient of s_slot in d
    double ro[kRing];
};

struct Args {
    int K;
    int first;
    long n, ld;
    int nchunks;
    float* param;
    const float* g;
    float* prev_g;
    float* d;
    float* S;  // [(K+1)][ld]
    float* Y;
    LbfgsState* st;
    float* partials;  // [nchunks][stride]  stride = kNV * ring + 8
    double* SY;       // [ring][ring]  s_i . y_j
    double* YY;       // [ring][ring]  y_i . y_j
};

__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// ---------------------------------------------------------------------------------------------------------------
// 1. multi-dot pass.  CTA c owns elements [c*kChunk, (c+1)*kChunk).  partial layout per CTA:
//    [slot*kNV + {0: s.g, 1: y.g, 2: s.y_new, 3: y.y_new, 4: y.s_new}] for every ring slot, then the globals
//    [G0 + {0: y_new.s_new, 1: y_new.y_new, 2: g.g, 3: |g|_1, 4: s_new.g, 5: y_new.g}].
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) lbfgs_dots_kernel(const Args a) {
    __shared__ float red[kThreads / 32][8];
    const LbfgsState* st = a.st;
    if (st->halted) return;
    const int ring = a.K + 1;
    const int len = st->hist_len, head = st->head;
    const int cand = a.first ? -1 : (head + len) % ring;
    const float t = st->t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long base = (long)blockIdx.x * kChunk;
    float* part = a.partials + (size_t)blockIdx.x * (kNV * ring + 8);

    // this thread's elements: 2 float4 (or scalar tail elements) of g, y_new, s_new stay in registers
    float4 g4[2], y4[2], s4[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const long e = base + (long)(u * kThreads + threadIdx.x) * 4;
        ok[u] = e < a.n;  // n is padded to ld (multiple of 4) with zeros in every vector, so float4 access is safe
        g4[u] = y4[u] = s4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!ok[u]) continue;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e + 3 < a.n) g = *reinterpret_cast<const float4*>(a.g + e);
        else { g.x = a.g[e]; if (e + 1 < a.n) g.y = a.g[e + 1]; if (e + 2 < a.n) g.z = a.g[e + 2]; }
        g4[u] = g;
        if (a.first) {
            *reinterpret_cast<float4*>(a.prev_g + e) = g;
        } else {
            const float4 pg = *reinterpret_cast<const float4*>(a.prev_g + e);
            const float4 dd = *reinterpret_cast<const float4*>(a.d + e);
            y4[u] = make_float4(g.x - pg.x, g.y - pg.y, g.z - pg.z, g.w - pg.w);
            s4[u] = make_float4(dd.x * t, dd.y * t, dd.z * t, dd.w * t);
            *reinterpret_cast<float4*>(a.Y + (size_t)cand * a.ld + e) = y4[u];
            *reinterpret_cast<float4*>(a.S + (size_t)cand * a.ld + e) = s4[u];
            *reinterpret_cast<float4*>(a.prev_g + e) = g;
        }
    }
    auto block_out = [&](float (&v)[8], int nv, float* dst) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (k < nv) v[k] = warp_sum(v[k]);
        __syncthreads();
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k < nv) red[warp][k] = v[k];
        __syncthreads();
        if (threadIdx.x < nv) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) s += red[w][threadIdx.x];
            dst[threadIdx.x] = s;
        }
    };
    {
        float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            v[0] += dot4(y4[u], s4[u]);
            v[1] += dot4(y4[u], y4[u]);
            v[2] += dot4(g4[u], g4[u]);
            v[3] += fabsf(g4[u].x) + fabsf(g4[u].y) + fabsf(g4[u].z) + fabsf(g4[u].w);
            v[4] += dot4(s4[u], g4[u]);
            v[5] += dot4(y4[u], g4[u]);
        }
        block_out(v, 6, part + kNV * ring);
    }
    // stream the history: every row's chunk is read exactly once
    for (int e = 0; e < len; ++e) {
        const int slot = (head + e) % ring;
        const float* sr = a.S + (size_t)slot * a.ld + base;
        const float* yr = a.Y + (size_t)slot * a.ld + base;
        float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!ok[u]) continue;
            const long o = (long)(u * kThreads + threadIdx.x) * 4;
            const float4 sv = __ldcs(reinterpret_cast<const float4*>(sr + o));
            const float4 yv = __ldcs(reinterpret_cast<const float4*>(yr + o));
            v[0] += dot4(sv, g4[u]);
            v[1] += dot4(yv, g4[u]);
            v[2] += dot4(sv, y4[u]);
            v[3] += dot4(yv, y4[u]);
            v[4] += dot4(yv, s4[u]);
        }
        block_out(v, kNV, part + slot * kNV);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. scalar phase (one CTA)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) s += sh[w];
    return s;
}

__global__ void __launch_bounds__(kThreads) lbfgs_scalar_kernel(const Args a) {
    __shared__ double sh[kThreads / 32];
    __shared__ double SG[kRing], YG[kRing], SYn[kRing], YYn[kRing], YSn[kRing], al[kRing], be[kRing], ro[kRing];
    __shared__ double glob[8];
    LbfgsState* st = a.st;
    if (st->halted) return;
    const int ring = a.K + 1;
    const int stride = kNV * ring + 8;
    int len = st->hist_len, head = st->head, n_iter = st->n_iter;
    double H = st->H_diag;
    const int tid = threadIdx.x;

    // deterministic reduction of the per-CTA partials (fixed order over CTAs), one (slot, value) per thread
    for (int idx = tid; idx < kNV * ring + 6; idx += kThreads) {
        double s = 0;
        for (int c = 0; c < a.nchunks; ++c) s += (double)a.partials[(size_t)c * stride + idx];
        if (idx >= kNV * ring) glob[idx - kNV * ring] = s;
        else {
            const int slot = idx / kNV, k = idx % kNV;
            (k == 0 ? SG : k == 1 ? YG : k == 2 ? SYn : k == 3 ? YYn : YSn)[slot] = s;
        }
    }
    for (int s = tid; s < ring; s += kThreads) ro[s] = st->ro[s];
    __syncthreads();

    int cand = -1, accepted = 0;
    if (a.first) {
        n_iter = 1; len = 0; head = 0; H = 1.0;
    } else {
        n_iter += 1;
        cand = (head + len) % ring;
        const double ys = glob[0], yy = glob[1];
        if ((float)ys > 1e-10f) {
            accepted = 1;
            // new row / column of the inner-product matrices (old entries in ring order)
            for (int e = tid; e < len; e += kThreads) {
                const int sl = (head + e) % ring;
                a.SY[(size_t)sl * ring + cand] = SYn[sl];    // s_old . y_new
                a.SY[(size_t)cand * ring + sl] = YSn[sl];    // s_new . y_old
                a.YY[(size_t)sl * ring + cand] = YYn[sl];
                a.YY[(size_t)cand * ring + sl] = YYn[sl];
            }
            if (tid == 0) {
                a.SY[(size_t)cand * ring + cand] = ys;
                a.YY[(size_t)cand * ring + cand] = yy;
                ro[cand] = 1.0 / (double)(float)ys;
                SG[cand] = glob[4];
                YG[cand] = glob[5];
            }
            if (len == a.K) head = (head + 1) % ring; else len += 1;
            H = (double)((float)ys / (float)yy);
        }
        __syncthreads();
        __threadfence_block();
    }
    __syncthreads();

    // loop 1 (newest -> oldest):  al_e = ro_e * ( -SG_e - sum_{f>e} al_f SY[e][f] )
    for (int e = len - 1; e >= 0; --e) {
        const int se = (head + e) % ring;
        double term = 0;
        for (int f = e + 1 + tid; f < len; f += kThreads) {
            const int sf = (head + f) % ring;
            term += al[sf] * a.SY[(size_t)se * ring + sf];
        }
        const double sum = block_sum(term, sh);
        if (tid == 0) al[se] = (double)(float)(ro[se] * (-SG[se] - sum));
        __syncthreads();
    }
    // loop 2 (oldest -> newest):  be_e = ro_e * ( H (-YG_e - sum_f al_f YY[e][f]) + sum_{f<e} (al_f - be_f) SY[f][e] )
    for (int e = 0; e < len; ++e) {
        const int se = (head + e) % ring;
        double term = 0;
        for (int f = tid; f < len; f += kThreads) {
            const int sf = (head + f) % ring;
            term += -H * al[sf] * a.YY[(size_t)se * ring + sf];
            if (f < e) term += (al[sf] - be[sf]) * a.SY[(size_t)sf * ring + se];
        }
        const double sum = block_sum(term, sh);
        if (tid == 0) be[se] = (double)(float)(ro[se] * (-H * YG[se] + sum));
        __syncthreads();
    }
    // coefficients of the direction and the directional derivative g.d
    double term = 0;
    for (int e = tid; e < len; e += kThreads) {
        const int se = (head + e) % ring;
        const double cy = -H * al[se], cs = al[se] - be[se];
        st->cy[se] = (float)cy;
        st->cs[se] = (float)cs;
        term += cy * YG[se] + cs * SG[se];
    }
    const double sum = block_sum(term, sh);
    if (tid == 0) {
        const double gtd = -H * glob[2] + sum;
        const float lr = st->lr;
        const float t = (n_iter == 1) ? fminf(1.f, 1.f / (float)glob[3]) * lr : lr;
        st->cg = (float)(-H);
        st->n_iter = n_iter; st->hist_len = len; st->head = head; st->H_diag = (float)H; st->t = t;
        st->gtd = (float)gtd; st->cand = cand; st->accepted = accepted;
        if (accepted) st->ro[cand] = ro[cand];
        if ((float)gtd > -st->tol_change) st->halted = 1;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. direction + parameter update:  d = cg g + sum cy_j y_j + sum cs_j s_j ;  x += t d
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) lbfgs_update_kernel(const Args a) {
    __shared__ float cy[kRing], cs[kRing];
    __shared__ int slots[kRing];
    const LbfgsState* st = a.st;
    if (st->halted) return;
    const int ring = a.K + 1;
    const int len = st->hist_len, head = st->head;
    for (int e = threadIdx.x; e < len; e += kThreads) {
        const int sl = (head + e) % ring;
        slots[e] = sl; cy[e] = st->cy[sl]; cs[e] = st->cs[sl];
    }
    __syncthreads();
    const float cg = st->cg, t = st->t;
    const long n4 = a.ld >> 2;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const long e0 = i * 4;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e0 + 3 < a.n) g = *reinterpret_cast<const float4*>(a.g + e0);
        else { if (e0 < a.n) g.x = a.g[e0]; if (e0 + 1 < a.n) g.y = a.g[e0 + 1]; if (e0 + 2 < a.n) g.z = a.g[e0 + 2]; }
        float4 d = make_float4(cg * g.x, cg * g.y, cg * g.z, cg * g.w);
#pragma unroll 4
        for (int e = 0; e < len; ++e) {
            const size_t off = (size_t)slots[e] * a.ld + e0;
            const float4 yv = __ldcs(reinterpret_cast<const float4*>(a.Y + off));
            const float4 sv = __ldcs(reinterpret_cast<const float4*>(a.S + off));
            const float c1 = cy[e], c2 = cs[e];
            d.x = fmaf(c1, yv.x, d.x); d.y = fmaf(c1, yv.y, d.y); d.z = fmaf(c1, yv.z, d.z); d.w = fmaf(c1, yv.w, d.w);
            d.x = fmaf(c2, sv.x, d.x); d.y = fmaf(c2, sv.y, d.y); d.z = fmaf(c2, sv.z, d.z); d.w = fmaf(c2, sv.w, d.w);
        }
        *reinterpret_cast<float4*>(a.d + e0) = d;
        if (e0 + 3 < a.n) {
            float4 p = *reinterpret_cast<const float4*>(a.param + e0);
            p.x = fmaf(t, d.x, p.x); p.y = fmaf(t, d.y, p.y); p.z = fmaf(t, d.z, p.z); p.w = fmaf(t, d.w, p.w);
            *reinterpret_cast<float4*>(a.param + e0) = p;
        } else {
            if (e0 < a.n) a.param[e0] = fmaf(t, d.x, a.param[e0]);
            if (e0 + 1 < a.n) a.param[e0 + 1] = fmaf(t, d.y, a.param[e0 + 1]);
            if (e0 + 2 < a.n) a.param[e0 + 2] = fmaf(t, d.z, a.param[e0 + 2]);
        }
    }
}

}  // namespace
}  // namespace maua

using namespace maua;

struct maua_lbfgs {
    long n = 0, ld = 0;
    int K = 0;
    long calls = 0;
    int device = 0;
    int nchunks = 0;
    float *prev_g = nullptr, *d = nullptr, *S = nullptr, *Y = nullptr, *partials = nullptr;
    double *SY = nullptr, *YY = nullptr;
    LbfgsState* st = nullptr;
};

extern "C" {

MAUA_API int maua_lbfgs_create(long n, int history, float lr, float tolerance_change, maua_lbfgs_t** out) {
    MAUA_REQUIRE(out && n > 0 && history >= 1 && history <= kMaxHist, "maua_lbfgs_create: bad arguments (history <= %d)",
                 kMaxHist);
    maua_lbfgs* s = new maua_lbfgs();
    s->n = n; s->K = history;
    s->ld = (n + 3) & ~3L;
    s->nchunks = (int)((s->ld + kChunk - 1) / kChunk);
    cudaGetDevice(&s->device);
    const int ring = history + 1;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes);
    };
    // every vector is padded to a multiple of kChunk floats so the multi-dot pass can use unguarded float4 access
    const size_t padded = (size_t)s->nchunks * kChunk;
    alloc((void**)&s->prev_g, padded * sizeof(float));
    alloc((void**)&s->d, padded * sizeof(float));
    alloc((void**)&s->S, ((size_t)ring * s->ld + kChunk) * sizeof(float));
    alloc((void**)&s->Y, ((size_t)ring * s->ld + kChunk) * sizeof(float));
    alloc((void**)&s->partials, (size_t)s->nchunks * (kNV * ring + 8) * sizeof(float));
    alloc((void**)&s->SY, (size_t)ring * ring * sizeof(double));
    alloc((void**)&s->YY, (size_t)ring * ring * sizeof(double));
    alloc((void**)&s->st, sizeof(LbfgsState));
    if (e != cudaSuccess) {
        set_last_error("maua_lbfgs_create: cudaMalloc failed (%s) for n=%ld history=%d", cudaGetErrorString(e), n, history);
        cudaGetLastError();
        maua_lbfgs_destroy(s);
        return MAUA_ERR_OOM;
    }
    LbfgsState h;
    memset(&h, 0, sizeof(h));
    h.lr = lr; h.tol_change = tolerance_change; h.H_diag = 1.f; h.t = lr; h.cand = -1;
    MAUA_CUDA_CHECK(cudaMemcpy(s->st, &h, sizeof(h), cudaMemcpyHostToDevice));
    *out = s;
    return MAUA_OK;
}

MAUA_API void maua_lbfgs_destroy(maua_lbfgs_t* s) {
    if (!s) return;
    cudaFree(s->prev_g); cudaFree(s->d); cudaFree(s->S); cudaFree(s->Y); cudaFree(s->partials);
    cudaFree(s->SY); cudaFree(s->YY); cudaFree(s->st);
    delete s;
}

MAUA_API int maua_lbfgs_step(maua_lbfgs_t* s, float* param, const float* grad, maua_stream_t stream) {
    MAUA_REQUIRE(s && param && grad, "maua_lbfgs_step: null pointer");
    MAUA_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0,
                 "maua_lbfgs_step: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    Args a;
    a.K = s->K; a.first = s->calls == 0; a.n = s->n; a.ld = s->ld; a.nchunks = s->nchunks;
    a.param = param; a.g = grad; a.prev_g = s->prev_g; a.d = s->d; a.S = s->S; a.Y = s->Y;
    a.st = s->st; a.partials = s->partials; a.SY = s->SY; a.YY = s->YY;
    lbfgs_dots_kernel<<<s->nchunks, kThreads, 0, st>>>(a);
    MAUA_CUDA_CHECK(cudaGetLastError());
    lbfgs_scalar_kernel<<<1, kThreads, 0, st>>>(a);
    MAUA_CUDA_CHECK(cudaGetLastError());
    long blocks = ((s->ld >> 2) + kThreads - 1) / kThreads;
    if (blocks > 148 * 8) blocks = 148 * 8;
    lbfgs_update_kernel<<<(int)blocks, kThreads, 0, st>>>(a);
    MAUA_CUDA_CHECK(cudaGetLastError());
    s->calls += 1;
    return MAUA_OK;
}

MAUA_API int maua_lbfgs_query(maua_lbfgs_t* s, int* n_iter, int* hist_len, int* halted, maua_stream_t stream) {
    MAUA_REQUIRE(s, "maua_lbfgs_query: null state");
    LbfgsState h;
    MAUA_CUDA_CHECK(cudaMemcpyAsync(&h, s->st, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    MAUA_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    if (n_iter) *n_iter = h.n_iter;
    if (hist_len) *hist_len = h.hist_len;
    if (halted) *halted = h.halted;
    return MAUA_OK;
}

}  // extern "C"
