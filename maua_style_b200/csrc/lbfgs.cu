// L-BFGS pixel update, fully device-resident: ONE cooperative kernel launch per iteration, no host sync.
//
// Replaces torch.optim.LBFGS as the reference drives it (optim.py:180-191: max_iter = num_iters,
// tolerance_grad = tolerance_change = -1, history 100, lr 1, NO line search), i.e. per iteration:
//     y = g - g_prev ; s = t d ; if y.s > 1e-10: push (s, y), ro = 1/y.s, H = y.s / y.y
//     q = -g ; for i newest..oldest: al_i = ro_i s_i.q ; q -= al_i y_i
//     r = H q ; for i oldest..newest: be_i = ro_i y_i.r ; r += (al_i - be_i) s_i ; d = r
//     t = (first iteration) ? min(1, 1/|g|_1) lr : lr ;  if g.d > -tolerance_change: stop ;  x += t d
// The reference pays >= 4k+10 tiny launches plus a host sync per dot product.  Here the whole two-loop recursion runs
// inside one persistent kernel (2 CTAs per SM, launched cooperatively so all CTAs are co-resident): the 2k+3 dependent
// passes are separated by a software grid barrier instead of kernel boundaries.  Each pass fuses "apply the axpy whose
// coefficient the previous pass produced" with "partial dot product for the next coefficient"; after the barrier every
// CTA re-reduces the per-CTA partials in the same fixed order, so all CTAs hold bit-identical scalars (deterministic,
// no float atomics, no broadcast step).  q / r live in the d buffer (L2 resident: 12.6 MB at 1024^2), so HBM traffic
// is the two reads of every history vector: 16 k n bytes per iteration.
#include <cooperative_groups.h>

#include "maua_b200.h"
#include "pointwise.cuh"

namespace maua {

namespace {

constexpr int kMaxHist = 256;
constexpr int kLThreads = 256;
constexpr int kCtasPerSm = 2;

struct LbfgsState {  // device resident
    int n_iter, hist_len, head, halted;
    float t, H_diag, lr, tol_change;
    float gtd, ys, yy, pad0;
    unsigned long long barrier;  // monotonically increasing arrival counter of the grid barrier
    float ro[kMaxHist + 1];
};

struct StepArgs {
    int K;      // history capacity (ring has K + 1 slots)
    int hb;     // host-side upper bound of the history length for this call
    int first;  // first call: no (s, y) pair yet
    long n;
    long ld;    // history row stride in floats (n rounded up to a multiple of 4 so rows stay float4 aligned)
    unsigned long long bar_base;  // barriers executed by all previous launches
    float* param;
    const float* g;
    float* prev_g;
    float* d;
    float* S;  // [(K+1)][n]
    float* Y;
    LbfgsState* st;
    double* partials;  // [2][grid][2]
};

__device__ __forceinline__ float4 ld4(const float* p, long i) { return reinterpret_cast<const float4*>(p)[i]; }
__device__ __forceinline__ float4 ld4_stream(const float* p, long i) { return __ldcs(reinterpret_cast<const float4*>(p) + i); }
__device__ __forceinline__ void st4(float* p, long i, float4 v) { reinterpret_cast<float4*>(p)[i] = v; }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

template <class Body>
__device__ __forceinline__ void for_each4(long n, Body body) {
    const long n4 = n >> 2;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) body(i, true);
    for (long i = (n4 << 2) + blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        body(i, false);
}

// Software grid barrier + deterministic all-reduce of two doubles.  Every CTA publishes its partial sums, arrives on
// the monotonic counter, waits until all CTAs of this barrier generation have arrived, then reduces all partials in
// index order.  `gen` counts barriers since the state was created (never reset, so no ABA problem).
__device__ __forceinline__ void grid_allreduce2(double a0, double a1, const StepArgs& a, unsigned long long gen,
                                                double (&tot)[2]) {
    __shared__ double sh[2][kLThreads / 32];
    __shared__ double bc[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* part = a.partials + (gen & 1) * (size_t)gridDim.x * 2;
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) { sh[0][warp] = a0; sh[1][warp] = a1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int i = 0; i < kLThreads / 32; ++i) { s0 += sh[0][i]; s1 += sh[1][i]; }
        part[2 * blockIdx.x] = s0;
        part[2 * blockIdx.x + 1] = s1;
        __threadfence();
        atomicAdd(&a.st->barrier, 1ULL);
        const unsigned long long target = (gen + 1) * (unsigned long long)gridDim.x;
        unsigned long long seen;
        do {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(&a.st->barrier) : "memory");
        } while (seen < target);
        __threadfence();
    }
    __syncthreads();
    double p0 = 0, p1 = 0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += kLThreads) {
        p0 += __ldcg(part + 2 * b);
        p1 += __ldcg(part + 2 * b + 1);
    }
    p0 = warp_sum(p0);
    p1 = warp_sum(p1);
    if (lane == 0) { sh[0][warp] = p0; sh[1][warp] = p1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int i = 0; i < kLThreads / 32; ++i) { s0 += sh[0][i]; s1 += sh[1][i]; }
        bc[0] = s0; bc[1] = s1;
    }
    __syncthreads();
    tot[0] = bc[0];
    tot[1] = bc[1];
}

// One fused pass over the vectors: q (+)= coef * pv ; [q *= scale] ; partial dot of dv with the updated q.
template <bool FINAL>
__device__ __forceinline__ void fused_pass(const StepArgs& a, const float* pv, float coef, bool scale, float Hd,
                                           const float* dv, float& l0, float& l1) {
    l0 = 0.f;
    l1 = 0.f;
    if (!pv && !dv && !scale) return;
    for_each4(a.n, [&](long i, bool v4) {
        if (v4) {
            float4 q = ld4(a.d, i);
            if (pv) {
                const float4 h = ld4_stream(pv, i);
                q.x = fmaf(coef, h.x, q.x); q.y = fmaf(coef, h.y, q.y); q.z = fmaf(coef, h.z, q.z); q.w = fmaf(coef, h.w, q.w);
            }
            if (scale) { q.x *= Hd; q.y *= Hd; q.z *= Hd; q.w *= Hd; }
            if (pv || scale) st4(a.d, i, q);
            if (dv) {
                const float4 h = ld4_stream(dv, i);
                l0 += dot4(h, q);
                if (FINAL) l1 += fabsf(h.x) + fabsf(h.y) + fabsf(h.z) + fabsf(h.w);
            }
        } else {
            float q = a.d[i];
            if (pv) q = fmaf(coef, pv[i], q);
            if (scale) q *= Hd;
            if (pv || scale) a.d[i] = q;
            if (dv) { l0 += dv[i] * q; if (FINAL) l1 += fabsf(dv[i]); }
        }
    });
}

__global__ void __launch_bounds__(kLThreads, kCtasPerSm) lbfgs_step_kernel(const StepArgs a) {
    __shared__ float al_s[kMaxHist + 1];
    LbfgsState* st = a.st;
    // scalar state: identical in every CTA (read before anyone writes; written back by CTA 0 at the very end)
    const int halted = st->halted;
    int len = st->hist_len, head = st->head, n_iter = st->n_iter;
    float H_diag = st->H_diag, t = st->t;
    const float lr = st->lr, tol_change = st->tol_change;
    const int ring = a.K + 1;
    unsigned long long gen = a.bar_base;
    double tot[2];
    float l0, l1;
    int cand = -1;
    float ro_cand = 0.f;

    // ---- memory update: y = g - g_prev, s = t d, q = -g -------------------------------------------------------
    if (a.first) {
        if (!halted)
            for_each4(a.n, [&](long i, bool v4) {
                if (v4) {
                    const float4 g = ld4(a.g, i);
                    st4(a.prev_g, i, g);
                    st4(a.d, i, make_float4(-g.x, -g.y, -g.z, -g.w));
                } else {
                    a.prev_g[i] = a.g[i];
                    a.d[i] = -a.g[i];
                }
            });
        grid_allreduce2(0, 0, a, gen++, tot);
        n_iter = 1; len = 0; head = 0; H_diag = 1.f;
    } else {
        cand = (head + len) % ring;
        float* Yc = a.Y + (size_t)cand * a.ld;
        float* Sc = a.S + (size_t)cand * a.ld;
        l0 = l1 = 0.f;
        if (!halted)
            for_each4(a.n, [&](long i, bool v4) {
                if (v4) {
                    const float4 g = ld4(a.g, i), pg = ld4(a.prev_g, i), dd = ld4(a.d, i);
                    const float4 y = make_float4(g.x - pg.x, g.y - pg.y, g.z - pg.z, g.w - pg.w);
                    const float4 s = make_float4(dd.x * t, dd.y * t, dd.z * t, dd.w * t);
                    st4(Yc, i, y); st4(Sc, i, s); st4(a.prev_g, i, g);
                    st4(a.d, i, make_float4(-g.x, -g.y, -g.z, -g.w));
                    l0 += dot4(y, s); l1 += dot4(y, y);
                } else {
                    const float g = a.g[i], y = g - a.prev_g[i], s = a.d[i] * t;
                    Yc[i] = y; Sc[i] = s; a.prev_g[i] = g; a.d[i] = -g;
                    l0 += y * s; l1 += y * y;
                }
            });
        grid_allreduce2(l0, l1, a, gen++, tot);
        if (!halted) {
            n_iter += 1;
            const float ys = (float)tot[0], yy = (float)tot[1];
            if (ys > 1e-10f) {
                if (len == a.K) head = (head + 1) % ring; else len += 1;
                ro_cand = 1.f / ys;
                H_diag = ys / yy;
            } else {
                cand = -1;
            }
        }
    }
    auto ro_of = [&](int slot) { return slot == cand ? ro_cand : st->ro[slot]; };

    // ---- two-loop recursion -------------------------------------------------------------------------------------
    const float* pend_v = nullptr;  // vector of the pending axpy
    float pend_c = 0.f;
    for (int j = 0; j < a.hb; ++j) {  // loop 1: newest -> oldest
        const float* dv = nullptr;
        int slot = -1;
        if (!halted && j < len) { slot = (head + (len - 1 - j)) % ring; dv = a.S + (size_t)slot * a.ld; }
        fused_pass<false>(a, halted ? nullptr : pend_v, pend_c, false, 1.f, dv, l0, l1);
        grid_allreduce2(l0, l1, a, gen++, tot);
        if (slot >= 0) {
            const float al = (float)tot[0] * ro_of(slot);
            if (threadIdx.x == 0) al_s[slot] = al;
            pend_v = a.Y + (size_t)slot * a.ld;
            pend_c = -al;
        } else {
            pend_v = nullptr;
        }
    }
    __syncthreads();
    {   // r = H_diag * q, first dot of loop 2
        const float* dv = nullptr;
        int slot = -1;
        if (!halted && len > 0) { slot = head; dv = a.Y + (size_t)slot * a.ld; }
        fused_pass<false>(a, halted ? nullptr : pend_v, pend_c, !halted, H_diag, dv, l0, l1);
        grid_allreduce2(l0, l1, a, gen++, tot);
        if (slot >= 0) {
            pend_v = a.S + (size_t)slot * a.ld;
            pend_c = al_s[slot] - (float)tot[0] * ro_of(slot);
        } else {
            pend_v = nullptr;
        }
    }
    for (int j = 1; j < a.hb; ++j) {  // loop 2: oldest -> newest
        const float* dv = nullptr;
        int slot = -1;
        if (!halted && j < len) { slot = (head + j) % ring; dv = a.Y + (size_t)slot * a.ld; }
        fused_pass<false>(a, halted ? nullptr : pend_v, pend_c, false, 1.f, dv, l0, l1);
        grid_allreduce2(l0, l1, a, gen++, tot);
        if (slot >= 0) {
            pend_v = a.S + (size_t)slot * a.ld;
            pend_c = al_s[slot] - (float)tot[0] * ro_of(slot);
        } else {
            pend_v = nullptr;
        }
    }
    // last axpy, directional derivative g.d and |g|_1
    fused_pass<true>(a, halted ? nullptr : pend_v, pend_c, false, 1.f, halted ? nullptr : a.g, l0, l1);
    grid_allreduce2(l0, l1, a, gen++, tot);
    int new_halted = halted;
    float gtd = 0.f;
    if (!halted) {
        gtd = (float)tot[0];
        t = (n_iter == 1) ? fminf(1.f, 1.f / (float)tot[1]) * lr : lr;
        if (gtd > -tol_change) new_halted = 1;
    }
    // ---- x += t d ----------------------------------------------------------------------------------------------
    if (!new_halted)
        for_each4(a.n, [&](long i, bool v4) {
            if (v4) {
                float4 p = ld4(a.param, i);
                const float4 dd = ld4(a.d, i);
                p.x = fmaf(t, dd.x, p.x); p.y = fmaf(t, dd.y, p.y); p.z = fmaf(t, dd.z, p.z); p.w = fmaf(t, dd.w, p.w);
                st4(a.param, i, p);
            } else {
                a.param[i] = fmaf(t, a.d[i], a.param[i]);
            }
        });
    if (blockIdx.x == 0 && threadIdx.x == 0 && !halted) {
        st->n_iter = n_iter; st->hist_len = len; st->head = head; st->H_diag = H_diag; st->t = t;
        st->halted = new_halted; st->gtd = gtd;
        if (cand >= 0) st->ro[cand] = ro_cand;
    }
}

int barriers_per_step(int hb) { return 1 + hb + 1 + (hb > 1 ? hb - 1 : 0) + 1; }

}  // namespace
}  // namespace maua

using namespace maua;

struct maua_lbfgs {
    long n = 0;
    int K = 0;
    long calls = 0;
    unsigned long long barriers = 0;
    int device = 0;
    int grid = 0;
    float *prev_g = nullptr, *d = nullptr, *S = nullptr, *Y = nullptr;
    LbfgsState* st = nullptr;
    double* partials = nullptr;
};

extern "C" {

MAUA_API int maua_lbfgs_create(long n, int history, float lr, float tolerance_change, maua_lbfgs_t** out) {
    MAUA_REQUIRE(out && n > 0 && history >= 1 && history <= kMaxHist, "maua_lbfgs_create: bad arguments (history <= %d)",
                 kMaxHist);
    maua_lbfgs* s = new maua_lbfgs();
    s->n = n; s->K = history;
    cudaGetDevice(&s->device);
    int sms = 0, coop = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, s->device);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lbfgs_step_kernel, kLThreads, 0);
    if (!coop || sms <= 0 || per_sm <= 0) {
        set_last_error("maua_lbfgs_create: device %d cannot launch the cooperative L-BFGS kernel", s->device);
        delete s;
        return MAUA_ERR_CUDA;
    }
    if (per_sm > kCtasPerSm) per_sm = kCtasPerSm;
    s->grid = sms * per_sm;
    const size_t vec = ((size_t)n * sizeof(float) + 255) & ~size_t(255);
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    alloc((void**)&s->prev_g, vec);
    alloc((void**)&s->d, vec);
    const size_t ld = ((size_t)n + 3) & ~size_t(3);
    alloc((void**)&s->S, (size_t)(history + 1) * ld * sizeof(float) + 256);
    alloc((void**)&s->Y, (size_t)(history + 1) * ld * sizeof(float) + 256);
    alloc((void**)&s->st, sizeof(LbfgsState));
    alloc((void**)&s->partials, sizeof(double) * 2 * 2 * s->grid);
    if (e != cudaSuccess) {
        set_last_error("maua_lbfgs_create: cudaMalloc failed (%s) for n=%ld history=%d", cudaGetErrorString(e), n, history);
        cudaGetLastError();
        maua_lbfgs_destroy(s);
        return MAUA_ERR_OOM;
    }
    LbfgsState h;
    memset(&h, 0, sizeof(h));
    h.lr = lr; h.tol_change = tolerance_change; h.H_diag = 1.f; h.t = lr;
    MAUA_CUDA_CHECK(cudaMemcpy(s->st, &h, sizeof(h), cudaMemcpyHostToDevice));
    *out = s;
    return MAUA_OK;
}

MAUA_API void maua_lbfgs_destroy(maua_lbfgs_t* s) {
    if (!s) return;
    cudaFree(s->prev_g); cudaFree(s->d); cudaFree(s->S); cudaFree(s->Y); cudaFree(s->st); cudaFree(s->partials);
    delete s;
}

MAUA_API int maua_lbfgs_step(maua_lbfgs_t* s, float* param, const float* grad, maua_stream_t stream) {
    MAUA_REQUIRE(s && param && grad, "maua_lbfgs_step: null pointer");
    MAUA_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0,
                 "maua_lbfgs_step: pointers must be 16-byte aligned");
    StepArgs a;
    a.K = s->K; a.n = s->n; a.ld = (s->n + 3) & ~3L; a.param = param; a.g = grad; a.prev_g = s->prev_g; a.d = s->d; a.S = s->S; a.Y = s->Y;
    a.st = s->st; a.partials = s->partials;
    a.hb = (int)(s->calls < s->K ? s->calls : s->K);  // host upper bound of the history length
    a.first = s->calls == 0;
    a.bar_base = s->barriers;
    void* kargs[] = {(void*)&a};
    MAUA_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)lbfgs_step_kernel, dim3(s->grid), dim3(kLThreads), kargs, 0,
                                                (cudaStream_t)stream));
    s->barriers += barriers_per_step(a.hb);
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
