// L-BFGS pixel update, fully device-resident: four launches per iteration, two streaming passes over the history.
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
// so al / be follow from O(k^2) scalar work and the direction is one linear combination
//     d = -H g - H sum_j al_j y_j + sum_j (al_j - be_j) s_j.
// SY / YY are kept on the device and extended by one row/column per accepted pair.  Per iteration:
//   1. lbfgs_dots_kernel   : forms y, s into the ring, streams every history vector ONCE and takes all its dot products
//                            with g, y_new, s_new at the same time (multi-dot).  Persistent grid sized to the SM count,
//                            every CTA owns an equal contiguous span of the vectors (no wave quantisation), partial
//                            sums accumulate in shared memory and are written once per CTA;
//   2. lbfgs_reduce_kernel : one warp per dot product sums the per-CTA partials in a fixed order (deterministic);
//   3. lbfgs_scalar_kernel : applies the y.s gate, updates SY / YY / ro / H, runs both recurrences (triangular solves,
//                            column-oriented inside ONE warp so a step costs a shuffle, not a block barrier),
//                            computes g.d, t and the halt flag;
//   4. lbfgs_update_kernel : streams the history a second time:  d = combination,  x += t d  (same balanced spans).
// HBM traffic = 2 reads of the 2k history vectors = 16 k n bytes per iteration: the floor of the
// algorithm (the first pass needs the new gradient, the second needs the coefficients the first one produces).
#include <cstdlib>

#include "maua_b200.h"
#include "pointwise.cuh"

namespace maua {

namespace {

constexpr int kMaxHist = 256;
constexpr int kRing = kMaxHist + 1;
constexpr int kThreads = 256;
constexpr int kSub4 = 2 * kThreads;  // float4 elements of one sub-chunk (2 per thread)
constexpr int kNV = 5;               // dot products per history slot
constexpr int kNG = 8;               // global dot products (6 used)

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
    int grid;         // CTAs of the two streaming kernels
    int small;        // at most one float4 per thread of the streaming kernels: the passes are load-latency chains (see lbfgs_dots_kernel)
    int mats_in_smem; // 2: SY and YY cached in shared memory by the scalar kernel, 1: SY only, 0: neither
    float* param;
    const float* g;
    float* prev_g;
    float* d;
    float* S;  // [(K+1)][ld]
    float* Y;
    LbfgsState* st;
    float* partials;  // [nacc][grid]  (transposed: the reduce kernel reads one row per dot product, coalesced)
    double* reduced;  // [nacc]        nacc = kNV * ring + kNG
    double* SY;       // [ring][ring]  s_i . y_j
    double* YY;       // [ring][ring]  y_i . y_j
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// ---------------------------------------------------------------------------------------------------------------
// 1. multi-dot pass.  CTA b owns the float4 elements [n4 b / G, n4 (b+1) / G).  Accumulator layout:
//    [slot*kNV + {0: s.g, 1: y.g, 2: s.y_new, 3: y.y_new, 4: y.s_new}] for every ring slot, then the globals
//    [G0 + {0: y_new.s_new, 1: y_new.y_new, 2: g.g, 3: |g|_1, 4: s_new.g, 5: y_new.g}].
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) lbfgs_dots_kernel(const Args a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float acc[];  // [nacc]
    __shared__ float red[kThreads / 32][2 * kNV];
    const LbfgsState* st = a.st;
    if (st->halted) return;
    const int ring = a.K + 1;
    const int nacc = kNV * ring + kNG;
    const int len = st->hist_len, head = st->head;
    const int cand = a.first ? -1 : (head + len) % ring;
    const float t = st->t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < nacc; i += kThreads) acc[i] = 0.f;

    // warp shuffle -> per-warp slots -> one thread per value adds the block total to its shared accumulator
    auto block_acc = [&](float (&v)[2 * kNV], int nv, int dst) {
#pragma unroll
        for (int k = 0; k < 2 * kNV; ++k)
            if (k < nv) v[k] = warp_sum(v[k]);
        __syncthreads();  // previous use of red[] is complete (also orders the zero-fill of acc[] the first time)
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < 2 * kNV; ++k)
                if (k < nv) red[warp][k] = v[k];
        __syncthreads();
        if (threadIdx.x < nv) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) s += red[w][threadIdx.x];
            acc[dst + threadIdx.x] += s;
        }
    };

    const long n4 = a.ld >> 2;
    const long b0 = n4 * blockIdx.x / gridDim.x, b1 = n4 * (blockIdx.x + 1) / gridDim.x;
    for (long c0 = b0; c0 < b1; c0 += kSub4) {
        // this thread's elements of the sub-chunk: 2 float4 of g, y_new, s_new stay in registers
        float4 g4[2], y4[2], s4[2];
        bool ok[2];
        long off[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long i4 = c0 + u * kThreads + threadIdx.x;
            const long e = i4 * 4;
            off[u] = e;
            ok[u] = i4 < b1;
            g4[u] = y4[u] = s4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!ok[u]) continue;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);  // the gradient is n long, every other vector is padded to ld
            if (e + 3 < a.n) g = *reinterpret_cast<const float4*>(a.g + e);
            else { g.x = a.g[e]; if (e + 1 < a.n) g.y = a.g[e + 1]; if (e + 2 < a.n) g.z = a.g[e + 2]; }
            g4[u] = g;
            if (!a.first) {
                const float4 pg = *reinterpret_cast<const float4*>(a.prev_g + e);
                const float4 dd = *reinterpret_cast<const float4*>(a.d + e);
                y4[u] = make_float4(g.x - pg.x, g.y - pg.y, g.z - pg.z, g.w - pg.w);
                s4[u] = make_float4(dd.x * t, dd.y * t, dd.z * t, dd.w * t);
                *reinterpret_cast<float4*>(a.Y + (size_t)cand * a.ld + e) = y4[u];
                *reinterpret_cast<float4*>(a.S + (size_t)cand * a.ld + e) = s4[u];
            }
            *reinterpret_cast<float4*>(a.prev_g + e) = g;
        }
        {
            float v[2 * kNV] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                v[0] += dot4(y4[u], s4[u]);
                v[1] += dot4(y4[u], y4[u]);
                v[2] += dot4(g4[u], g4[u]);
                v[3] += fabsf(g4[u].x) + fabsf(g4[u].y) + fabsf(g4[u].z) + fabsf(g4[u].w);
                v[4] += dot4(s4[u], g4[u]);
                v[5] += dot4(y4[u], g4[u]);
            }
            block_acc(v, 6, kNV * ring);
        }
        // stream the history, two slots (8 independent 16-byte loads per thread) per round
        for (int e = 0; e < len; e += 2) {
            const int slot0 = (head + e) % ring;
            const bool two = e + 1 < len;
            const int slot1 = two ? (head + e + 1) % ring : slot0;
            float4 sv[2][2], yv[2][2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                sv[0][u] = sv[1][u] = yv[0][u] = yv[1][u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (!ok[u]) continue;
                sv[0][u] = __ldcs(reinterpret_cast<const float4*>(a.S + (size_t)slot0 * a.ld + off[u]));
                yv[0][u] = __ldcs(reinterpret_cast<const float4*>(a.Y + (size_t)slot0 * a.ld + off[u]));
                if (two) {
                    sv[1][u] = __ldcs(reinterpret_cast<const float4*>(a.S + (size_t)slot1 * a.ld + off[u]));
                    yv[1][u] = __ldcs(reinterpret_cast<const float4*>(a.Y + (size_t)slot1 * a.ld + off[u]));
                }
            }
            float v[2 * kNV] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    v[q * kNV + 0] += dot4(sv[q][u], g4[u]);
                    v[q * kNV + 1] += dot4(yv[q][u], g4[u]);
                    v[q * kNV + 2] += dot4(sv[q][u], y4[u]);
                    v[q * kNV + 3] += dot4(yv[q][u], y4[u]);
                    v[q * kNV + 4] += dot4(yv[q][u], s4[u]);
                }
            if (two && slot1 == slot0 + 1) {
                block_acc(v, 2 * kNV, slot0 * kNV);  // adjacent ring slots: one reduction for both
            } else {
                block_acc(v, kNV, slot0 * kNV);
                if (two) {
                    float w[2 * kNV];
#pragma unroll
                    for (int k = 0; k < kNV; ++k) { w[k] = v[kNV + k]; w[kNV + k] = 0.f; }
                    block_acc(w, kNV, slot1 * kNV);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nacc; i += kThreads) a.partials[(size_t)i * gridDim.x + blockIdx.x] = acc[i];
}

// Same pass for SMALL vectors (at most one float4 per thread, <= ~300^2 images): there the block-wide reduction after every two
// slots made the pass a chain of ~50 load-latency + barrier rounds (2 TB/s at 256^2).
__global__ void __launch_bounds__(kThreads) lbfgs_dots_small_kernel(const Args a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float acc[];  // [nacc], then the staged sub-chunk of g, y_new, s_new: 3 x kSub4 float4
    __shared__ float red[kThreads / 32][2 * kNV];
    const LbfgsState* st = a.st;
    if (st->halted) return;
    const int ring = a.K + 1;
    const int nacc = kNV * ring + kNG;
    float4* sg = reinterpret_cast<float4*>(acc + ((nacc + 3) & ~3));
    float4* sy = sg + kSub4;
    float4* ss = sy + kSub4;
    const int len = st->hist_len, head = st->head;
    const int cand = a.first ? -1 : (head + len) % ring;
    const float t = st->t;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < nacc; i += kThreads) acc[i] = 0.f;

    // warp shuffle -> per-warp slots -> one thread per value adds the block total to its shared accumulator
    auto block_acc = [&](float (&v)[2 * kNV], int nv, int dst) {
#pragma unroll
        for (int k = 0; k < 2 * kNV; ++k)
            if (k < nv) v[k] = warp_sum(v[k]);
        __syncthreads();  // previous use of red[] is complete (also orders the zero-fill of acc[] the first time)
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < 2 * kNV; ++k)
                if (k < nv) red[warp][k] = v[k];
        __syncthreads();
        if (threadIdx.x < nv) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) s += red[w][threadIdx.x];
            acc[dst + threadIdx.x] += s;
        }
    };

    const long n4 = a.ld >> 2;
    const long b0 = n4 * blockIdx.x / gridDim.x, b1 = n4 * (blockIdx.x + 1) / gridDim.x;
    for (long c0 = b0; c0 < b1; c0 += kSub4) {
        // this thread's elements of the sub-chunk: 2 float4 of g, y_new, s_new stay in registers
        float4 g4[2], y4[2], s4[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long i4 = c0 + u * kThreads + threadIdx.x;
            const long e = i4 * 4;
            ok[u] = i4 < b1;
            g4[u] = y4[u] = s4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!ok[u]) continue;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);  // the gradient is n long, every other vector is padded to ld
            if (e + 3 < a.n) g = *reinterpret_cast<const float4*>(a.g + e);
            else { g.x = a.g[e]; if (e + 1 < a.n) g.y = a.g[e + 1]; if (e + 2 < a.n) g.z = a.g[e + 2]; }
            g4[u] = g;
            if (!a.first) {
                const float4 pg = *reinterpret_cast<const float4*>(a.prev_g + e);
                const float4 dd = *reinterpret_cast<const float4*>(a.d + e);
                y4[u] = make_float4(g.x - pg.x, g.y - pg.y, g.z - pg.z, g.w - pg.w);
                s4[u] = make_float4(dd.x * t, dd.y * t, dd.z * t, dd.w * t);
                *reinterpret_cast<float4*>(a.Y + (size_t)cand * a.ld + e) = y4[u];
                *reinterpret_cast<float4*>(a.S + (size_t)cand * a.ld + e) = s4[u];
            }
            *reinterpret_cast<float4*>(a.prev_g + e) = g;
        }
        {
            float v[2 * kNV] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                v[0] += dot4(y4[u], s4[u]);
                v[1] += dot4(y4[u], y4[u]);
                v[2] += dot4(g4[u], g4[u]);
                v[3] += fabsf(g4[u].x) + fabsf(g4[u].y) + fabsf(g4[u].z) + fabsf(g4[u].w);
                v[4] += dot4(s4[u], g4[u]);
                v[5] += dot4(y4[u], g4[u]);
            }
            block_acc(v, 6, kNV * ring);
        }
        // Stream the history: one WARP per slot.  The sub-chunk's g / y_new / s_new sit in shared memory; warp w takes the slots
        // w, w + 8, ... and walks the sub-chunk with 32 lanes, five running dot products per lane in registers, one shuffle
        // reduction per slot, no block barrier inside the slot loop -- so a lane keeps up to 32 independent 16-byte loads in
        // flight.  (The first version reduced across the block after every two slots: at 256^2, where a CTA owns half a
        // sub-chunk, the pass was a chain of 50 load-latency + barrier rounds and ran at 2 TB/s.)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = u * kThreads + threadIdx.x;
            sg[i] = g4[u]; sy[i] = y4[u]; ss[i] = s4[u];   // zeros beyond the span
        }
        __syncthreads();
        const int nvalid = (int)((b1 - c0) < (long)kSub4 ? (b1 - c0) : (long)kSub4);
        // L2 prefetch of a slot's piece of this sub-chunk: 128-byte lines, lane l takes lines l and l + 32.  Issued two slots
        // ahead of the loads, so that the memory-level parallelism does not depend on how far ptxas hoists the loads themselves.
        auto prefetch_slot = [&](int e) {
            if (e >= len) return;
            const int slot = (head + e) % ring;
            const float4* Sp = reinterpret_cast<const float4*>(a.S + (size_t)slot * a.ld) + c0;
            const float4* Yp = reinterpret_cast<const float4*>(a.Y + (size_t)slot * a.ld) + c0;
            for (int i = lane * 8; i < nvalid; i += 256) { prefetch_l2(Sp + i); prefetch_l2(Yp + i); }
        };
        prefetch_slot(warp);
        prefetch_slot(warp + kThreads / 32);
        for (int e = warp; e < len; e += kThreads / 32) {
            prefetch_slot(e + 2 * (kThreads / 32));
            const int slot = (head + e) % ring;
            const float4* Sp = reinterpret_cast<const float4*>(a.S + (size_t)slot * a.ld) + c0;
            const float4* Yp = reinterpret_cast<const float4*>(a.Y + (size_t)slot * a.ld) + c0;
            float v[kNV] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
            for (int i = lane; i < nvalid; i += 32) {
                const float4 sv = __ldcs(Sp + i), yv = __ldcs(Yp + i);
                const float4 gg = sg[i], yn = sy[i], sn = ss[i];
                v[0] += dot4(sv, gg);
                v[1] += dot4(yv, gg);
                v[2] += dot4(sv, yn);
                v[3] += dot4(yv, yn);
                v[4] += dot4(yv, sn);
            }
#pragma unroll
            for (int k = 0; k < kNV; ++k) v[k] = warp_sum(v[k]);
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < kNV; ++k) acc[slot * kNV + k] += v[k];  // this warp is the only writer of this slot
            }
        }
        __syncthreads();  // the staged sub-chunk is overwritten by the next one
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nacc; i += kThreads) a.partials[(size_t)i * gridDim.x + blockIdx.x] = acc[i];
}

// ---------------------------------------------------------------------------------------------------------------
// 2. deterministic reduction of the per-CTA partials: one warp per dot product, fixed lane / shuffle order
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) lbfgs_reduce_kernel(const Args a) {
    pdl_wait();
    pdl_trigger();
    if (a.st->halted) return;
    const int nacc = kNV * (a.K + 1) + kNG;
    const int idx = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (idx >= nacc) return;
    const int lane = threadIdx.x & 31;
    const float* row = a.partials + (size_t)idx * a.grid;
    double s = 0.0;
    for (int c = lane; c < a.grid; c += 32) s += (double)__ldcg(row + c);
    s = warp_sum(s);
    if (lane == 0) a.reduced[idx] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// 3. scalar phase (one CTA; the two triangular solves run inside warp 0)
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

inline size_t dots_smem_bytes(int nacc) { return (size_t)((nacc + 3) & ~3) * sizeof(float) + 3 * (size_t)kSub4 * sizeof(float4); }


__global__ void __launch_bounds__(kThreads) lbfgs_scalar_kernel(const Args a) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ double mats[];  // optional cache of SY (and YY): [ring][ring] each
    __shared__ double sh[kThreads / 32];
    __shared__ double SG[kRing], YG[kRing], SYn[kRing], YYn[kRing], YSn[kRing], al[kRing], be[kRing], ro[kRing], wv[kRing];
    double* rhs = SYn;  // running right-hand sides of the two solves, chronological index (SYn / YYn / YSn are consumed by then)
    __shared__ double glob[kNG];
    __shared__ int slot_of[kRing];
    LbfgsState* st = a.st;
    if (st->halted) return;
    const int ring = a.K + 1;
    int len = st->hist_len, head = st->head, n_iter = st->n_iter;
    double H = st->H_diag;
    const int tid = threadIdx.x;

    for (int idx = tid; idx < kNV * ring + 6; idx += kThreads) {
        const double s = a.reduced[idx];
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
    }
    __syncthreads();  // block-scope ordering of the global SY / YY writes above with the reads below

    // chronological index -> ring slot, and the matrices staged in shared memory when they fit
    for (int e = tid; e < len; e += kThreads) slot_of[e] = (head + e) % ring;
    const double* SYm = a.SY;
    const double* YYm = a.YY;
    if (a.mats_in_smem >= 1) {
        for (int i = tid; i < ring * ring; i += kThreads) mats[i] = a.SY[i];
        SYm = mats;
        if (a.mats_in_smem >= 2) {
            for (int i = tid; i < ring * ring; i += kThreads) mats[ring * ring + i] = a.YY[i];
            YYm = mats + ring * ring;
        }
    }
    __syncthreads();

    // loop 1 (newest -> oldest):  al_e = ro_e * ( -SG_e - sum_{f>e} al_f SY[e][f] ).  Column-oriented inside warp 0: lane l keeps
    // the running right-hand sides of its entries e = l, l+32, ...  The entries are solved in blocks of 32: the lane first
    // loads its row of the 32 x 32 diagonal block into registers, so that a step of the dependent chain is multiply ->
    // round-to-float -> broadcast -> one fp64 FMA with no memory access in it (the first version re-read shared memory behind
    // every broadcast: ~500 cycles per step, 53 us at history 100); the finished block is then subtracted from the older
    // entries' right-hand sides.  Every r_f still receives its terms in the order e = len-1 .. f+1: the sums are bit-identical.
    // (The right-hand sides live in shared memory between blocks -- every lane only touches its own entries -- so the block
    // loops are runtime loops and only the 32 steps of a diagonal block are unrolled: a first version with everything unrolled
    // over 8 register-resident blocks was 41 k instructions of straight-line code and ran slower than the loop it replaced.)
    if (tid < 32) {
        const int lane = tid;
        for (int e = lane; e < len; e += 32) rhs[e] = -SG[slot_of[e]];
        __syncwarp();
        const int nblk = (len + 31) / 32;
#pragma unroll 1
        for (int jb = nblk - 1; jb >= 0; --jb) {
            const int top = min(len - 1 - jb * 32, 31);
            const int mine = jb * 32 + lane;                       // this lane's entry of the block
            const bool live = mine < len;
            const int s_mine = live ? slot_of[mine] : 0;
            const double ro_mine = live ? ro[s_mine] : 0.0;
            double m[32];                                          // SY[mine][jb*32 + l] for the newer entries of the block
#pragma unroll
            for (int l = 0; l < 32; ++l) m[l] = (live && l > lane && l <= top) ? SYm[(size_t)s_mine * ring + slot_of[jb * 32 + l]] : 0.0;
            double cur = live ? rhs[mine] : 0.0, al_mine = 0.0;
#pragma unroll
            for (int l = 31; l >= 0; --l) {
                if (l > top) continue;                             // (warp-uniform)
                double ae = (double)(float)(ro_mine * cur);        // meaningful in lane l, whose right-hand side is complete
                ae = __shfl_sync(0xffffffffu, ae, l);
                if (lane == l) al_mine = ae;
                cur -= ae * m[l];                                  // m[l] = 0 for the lanes at or above l
            }
            if (live) al[s_mine] = al_mine;
            __syncwarp();
            // the finished block against the older blocks (same order of terms as before: e descending)
#pragma unroll 1
            for (int j = 0; j < jb; ++j) {
                const int f = lane + 32 * j;
                const size_t rowf = (size_t)slot_of[f] * ring;
                double acc = rhs[f];
#pragma unroll 4
                for (int l = top; l >= 0; --l) {
                    const int se = slot_of[jb * 32 + l];
                    acc -= al[se] * SYm[rowf + se];
                }
                rhs[f] = acc;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // w_e = -H ( YG_e + sum_f al_f YY[e][f] )   (independent of be: all threads)
    for (int e = tid; e < len; e += kThreads) {
        const int se = slot_of[e];
        double s = YG[se];
        for (int f = 0; f < len; ++f) s += al[slot_of[f]] * YYm[(size_t)se * ring + slot_of[f]];
        wv[se] = -H * s;
    }
    __syncthreads();
    // loop 2 (oldest -> newest):  be_e = ro_e * ( w_e + sum_{f<e} (al_f - be_f) SY[f][e] ), same blocking: the lane holds its
    // COLUMN of the diagonal block (SY[older entry of the block][mine]) in registers
    if (tid < 32) {
        const int lane = tid;
        for (int e = lane; e < len; e += 32) rhs[e] = wv[slot_of[e]];
        __syncwarp();
        const int nblk = (len + 31) / 32;
#pragma unroll 1
        for (int jb = 0; jb < nblk; ++jb) {
            const int top = min(len - jb * 32, 32);                // entries in this block
            const int mine = jb * 32 + lane;
            const bool live = mine < len;
            const int s_mine = live ? slot_of[mine] : 0;
            const double ro_mine = live ? ro[s_mine] : 0.0;
            const double al_mine = live ? al[s_mine] : 0.0;
            double m[32];
#pragma unroll
            for (int l = 0; l < 32; ++l) m[l] = (live && l < lane && l < top) ? SYm[(size_t)slot_of[jb * 32 + l] * ring + s_mine] : 0.0;
            double cur = live ? rhs[mine] : 0.0, be_mine = 0.0;
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                if (l >= top) continue;                            // (warp-uniform)
                double b = (double)(float)(ro_mine * cur);
                b = __shfl_sync(0xffffffffu, b, l);
                if (lane == l) be_mine = b;
                const double c = __shfl_sync(0xffffffffu, al_mine, l) - b;
                cur += c * m[l];                                   // m[l] = 0 for the lanes at or below l
            }
            if (live) { be[s_mine] = be_mine; wv[s_mine] = al_mine - be_mine; }  // wv is free now: al - be of the finished block
            __syncwarp();
#pragma unroll 1
            for (int j = jb + 1; j < nblk; ++j) {
                const int f = lane + 32 * j;                       // newer entries
                if (f >= len) continue;
                const int sf = slot_of[f];
                double acc = rhs[f];
#pragma unroll 4
                for (int l = 0; l < top; ++l) {
                    const int se = slot_of[jb * 32 + l];
                    acc += wv[se] * SYm[(size_t)se * ring + sf];
                }
                rhs[f] = acc;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // coefficients of the direction and the directional derivative g.d
    double term = 0;
    for (int e = tid; e < len; e += kThreads) {
        const int se = slot_of[e];
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
// 4. direction + parameter update:  d = cg g + sum cy_j y_j + sum cs_j s_j ;  x += t d
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) lbfgs_update_kernel(const Args a) {
    pdl_wait();
    pdl_trigger();
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
    const long b0 = n4 * blockIdx.x / gridDim.x, b1 = n4 * (blockIdx.x + 1) / gridDim.x;  // equal contiguous spans
    for (long i = b0 + threadIdx.x; i < b1; i += kThreads) {
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

// The same update for SMALL vectors (one element per thread: pure load latency): batches of 8 slots, L2 prefetch two batches ahead.
__global__ void __launch_bounds__(kThreads) lbfgs_update_small_kernel(const Args a) {
    pdl_wait();
    pdl_trigger();
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
    const long b0 = n4 * blockIdx.x / gridDim.x, b1 = n4 * (blockIdx.x + 1) / gridDim.x;  // equal contiguous spans
    for (long i = b0 + threadIdx.x; i < b1; i += kThreads) {
        const long e0 = i * 4;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e0 + 3 < a.n) g = *reinterpret_cast<const float4*>(a.g + e0);
        else { if (e0 < a.n) g.x = a.g[e0]; if (e0 + 1 < a.n) g.y = a.g[e0 + 1]; if (e0 + 2 < a.n) g.z = a.g[e0 + 2]; }
        float4 d = make_float4(cg * g.x, cg * g.y, cg * g.z, cg * g.w);
        // batches of 8 slots: 16 independent 16-byte loads in flight per thread, then the FMAs in slot order (the order of the
        // terms is the old one: bit-identical sums).  At <= 512^2 a thread owns a single element and the pass is pure latency.
        constexpr int kBatch = 8;
        int e = 0;
        // L2 prefetch two batches ahead: one thread per 128-byte line (8 consecutive float4 elements share it)
        const bool pf = ((i & 7) == 0);
        auto prefetch_batch = [&](int e_first) {
            if (!pf) return;
#pragma unroll
            for (int k = 0; k < kBatch; ++k)
                if (e_first + k < len) {
                    const size_t off = (size_t)slots[e_first + k] * a.ld + e0;
                    prefetch_l2(a.Y + off); prefetch_l2(a.S + off);
                }
        };
        prefetch_batch(0);
        prefetch_batch(kBatch);
        for (; e + kBatch <= len; e += kBatch) {
            prefetch_batch(e + 2 * kBatch);
            float4 yv[kBatch], sv[kBatch];
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const size_t off = (size_t)slots[e + k] * a.ld + e0;
                yv[k] = __ldcs(reinterpret_cast<const float4*>(a.Y + off));
                sv[k] = __ldcs(reinterpret_cast<const float4*>(a.S + off));
            }
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
                const float c1 = cy[e + k], c2 = cs[e + k];
                d.x = fmaf(c1, yv[k].x, d.x); d.y = fmaf(c1, yv[k].y, d.y); d.z = fmaf(c1, yv[k].z, d.z); d.w = fmaf(c1, yv[k].w, d.w);
                d.x = fmaf(c2, sv[k].x, d.x); d.y = fmaf(c2, sv[k].y, d.y); d.z = fmaf(c2, sv[k].z, d.z); d.w = fmaf(c2, sv[k].w, d.w);
            }
        }
        for (; e < len; ++e) {
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
    int grid_dots = 0, grid_update = 0, mats_in_smem = 0;
    size_t scalar_smem = 0;
    float *prev_g = nullptr, *d = nullptr, *S = nullptr, *Y = nullptr, *partials = nullptr;
    double *SY = nullptr, *YY = nullptr, *reduced = nullptr;
    LbfgsState* st = nullptr;
    LbfgsState* st_init = nullptr;  // pristine copy of the device state (maua_lbfgs_reset)
};

extern "C" {

MAUA_API int maua_lbfgs_create(long n, int history, float lr, float tolerance_change, maua_lbfgs_t** out) {
    MAUA_REQUIRE(out && n > 0 && history >= 1 && history <= kMaxHist, "maua_lbfgs_create: bad arguments (history <= %d)",
                 kMaxHist);
    maua_lbfgs* s = new maua_lbfgs();
    s->n = n; s->K = history;
    s->ld = (n + 3) & ~3L;
    cudaGetDevice(&s->device);
    const int ring = history + 1;
    const int nacc = kNV * ring + kNG;
    cudaError_t e = cudaSuccess;
    // persistent grids: (CTAs that fit on one SM) x (SM count), every CTA gets an equal contiguous span
    int sms = 148, occ_dots = 1, occ_upd = 1;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
    if (sms <= 0) sms = 148;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_dots, lbfgs_dots_kernel, kThreads, nacc * sizeof(float));
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_upd, lbfgs_update_kernel, kThreads, 0);
    if (occ_dots < 1) occ_dots = 1;
    if (occ_upd < 1) occ_upd = 1;
    const long n4 = s->ld >> 2;
    const long max_useful = (n4 + kThreads - 1) / kThreads;  // at least one float4 per thread
    s->grid_dots = (int)(sms * (long)occ_dots < max_useful ? sms * (long)occ_dots : max_useful);
    s->grid_update = (int)(sms * (long)occ_upd < max_useful ? sms * (long)occ_upd : max_useful);
    const size_t mat = (size_t)ring * ring * sizeof(double);
    s->mats_in_smem = 2 * mat <= 180 * 1024 ? 2 : (mat <= 180 * 1024 ? 1 : 0);
    s->scalar_smem = s->mats_in_smem * mat;
    if (e == cudaSuccess && s->scalar_smem > 0)
        e = cudaFuncSetAttribute(lbfgs_scalar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->scalar_smem);
    auto alloc = [&](void** p, size_t bytes) {
        if (e == cudaSuccess) e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes);
    };
    // every vector is padded to ld (a multiple of 4 floats) with zeros so the streaming passes use float4 access
    alloc((void**)&s->prev_g, (size_t)s->ld * sizeof(float));
    alloc((void**)&s->d, (size_t)s->ld * sizeof(float));
    alloc((void**)&s->S, (size_t)ring * s->ld * sizeof(float));
    alloc((void**)&s->Y, (size_t)ring * s->ld * sizeof(float));
    alloc((void**)&s->partials, (size_t)nacc * s->grid_dots * sizeof(float));
    alloc((void**)&s->reduced, (size_t)nacc * sizeof(double));
    alloc((void**)&s->SY, mat);
    alloc((void**)&s->YY, mat);
    alloc((void**)&s->st, sizeof(LbfgsState));
    if (e != cudaSuccess) {
        set_last_error("maua_lbfgs_create: CUDA failure (%s) for n=%ld history=%d", cudaGetErrorString(e), n, history);
        cudaGetLastError();
        maua_lbfgs_destroy(s);
        return e == cudaErrorMemoryAllocation ? MAUA_ERR_OOM : MAUA_ERR_CUDA;
    }
    LbfgsState h;
    memset(&h, 0, sizeof(h));
    h.lr = lr; h.tol_change = tolerance_change; h.H_diag = 1.f; h.t = lr; h.cand = -1;
    MAUA_CUDA_CHECK(cudaMemcpy(s->st, &h, sizeof(h), cudaMemcpyHostToDevice));
    MAUA_CUDA_CHECK(cudaMalloc((void**)&s->st_init, sizeof(LbfgsState)));
    MAUA_CUDA_CHECK(cudaMemcpy(s->st_init, &h, sizeof(h), cudaMemcpyHostToDevice));
    *out = s;
    return MAUA_OK;
}

MAUA_API void maua_lbfgs_destroy(maua_lbfgs_t* s) {
    if (!s) return;
    cudaFree(s->prev_g); cudaFree(s->d); cudaFree(s->S); cudaFree(s->Y); cudaFree(s->partials); cudaFree(s->reduced);
    cudaFree(s->SY); cudaFree(s->YY); cudaFree(s->st); cudaFree(s->st_init);
    delete s;
}

MAUA_API int maua_lbfgs_step(maua_lbfgs_t* s, float* param, const float* grad, maua_stream_t stream) {
    MAUA_REQUIRE(s && param && grad, "maua_lbfgs_step: null pointer");
    MAUA_REQUIRE(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0,
                 "maua_lbfgs_step: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int nacc = kNV * (s->K + 1) + kNG;
    Args a;
    a.K = s->K; a.first = s->calls == 0; a.n = s->n; a.ld = s->ld; a.grid = s->grid_dots; a.mats_in_smem = s->mats_in_smem;
    a.small = (s->ld >> 2) <= (long)s->grid_dots * kThreads ? 1 : 0;
    if (const char* f = getenv("MAUA_LBFGS_SMALL")) a.small = atoi(f) != 0;  // developer A/B switch
    a.param = param; a.g = grad; a.prev_g = s->prev_g; a.d = s->d; a.S = s->S; a.Y = s->Y;
    a.st = s->st; a.partials = s->partials; a.reduced = s->reduced; a.SY = s->SY; a.YY = s->YY;
    if (a.small) MAUA_CUDA_CHECK(launch_pdl<PDL_LBFGS>(lbfgs_dots_small_kernel, dim3(s->grid_dots), dim3(kThreads), dots_smem_bytes(nacc), st, a));
    else MAUA_CUDA_CHECK(launch_pdl<PDL_LBFGS>(lbfgs_dots_kernel, dim3(s->grid_dots), dim3(kThreads), nacc * sizeof(float), st, a));
    MAUA_CUDA_CHECK(launch_pdl<PDL_LBFGS>(lbfgs_reduce_kernel, dim3((nacc + kThreads / 32 - 1) / (kThreads / 32)), dim3(kThreads), 0, st, a));
    MAUA_CUDA_CHECK(launch_pdl<PDL_LBFGS>(lbfgs_scalar_kernel, dim3(1), dim3(kThreads), s->scalar_smem, st, a));
    if (a.small) MAUA_CUDA_CHECK(launch_pdl<PDL_LBFGS>(lbfgs_update_small_kernel, dim3(s->grid_update), dim3(kThreads), 0, st, a));
    else MAUA_CUDA_CHECK(launch_pdl<PDL_LBFGS>(lbfgs_update_kernel, dim3(s->grid_update), dim3(kThreads), 0, st, a));
    s->calls += 1;
    return MAUA_OK;
}

MAUA_API int maua_lbfgs_reset(maua_lbfgs_t* s, maua_stream_t stream) {
    MAUA_REQUIRE(s, "maua_lbfgs_reset: null state");
    cudaStream_t st = (cudaStream_t)stream;
    const int ring = s->K + 1;
    const size_t mat = (size_t)ring * ring * sizeof(double);
    // history slots are only read up to hist_len and every vector's zero padding is preserved by the kernels, so the
    // (large) S / Y rings need no clearing: a reset costs a few hundred KB of memsets instead of re-allocating 2.5 GB
    MAUA_CUDA_CHECK(cudaMemsetAsync(s->prev_g, 0, (size_t)s->ld * sizeof(float), st));
    MAUA_CUDA_CHECK(cudaMemsetAsync(s->d, 0, (size_t)s->ld * sizeof(float), st));
    MAUA_CUDA_CHECK(cudaMemsetAsync(s->SY, 0, mat, st));
    MAUA_CUDA_CHECK(cudaMemsetAsync(s->YY, 0, mat, st));
    MAUA_CUDA_CHECK(cudaMemsetAsync(s->reduced, 0, (size_t)(kNV * ring + kNG) * sizeof(double), st));
    MAUA_CUDA_CHECK(cudaMemcpyAsync(s->st, s->st_init, sizeof(LbfgsState), cudaMemcpyDeviceToDevice, st));
    s->calls = 0;
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
