// Gram / covariance matrix as a symmetric tensor-core SYRK (tcgen05, TF32 operands, FP32 accumulate),
// plus the small StyleLoss reductions around it.
//
// Replaces reference loss.py:67-91 (GramMatrix.forward: torch.mm(x_flat, x_flat.t()) on the
// [C, H*W] view) and loss.py:141-181 (StyleLoss static + dynamic terms; for B = 1 the dynamic Gram is
// identical to the static one, so ONE SYRK serves both -- SURVEY.md section 8a "net effect of R4-R6").
//
// Layout: features are NHWC, i.e. the flattened matrix is F[P][C] with channels contiguous, and
//   G[c][d] = sum_p F[p][c] F[p][d]  is a GEMM whose M, N are channels and whose K is pixels.
// Both operands are therefore "MN-major" for the tensor core (the M/N index is the contiguous one).
// For TF32 the only MN-major shared-memory layout the tensor core accepts is the 128B swizzle with 32-byte
// atoms (UMMA layout type SWIZZLE_128B_BASE32B, TMA mode CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): a TMA box
// {32 channels, 64 pixels} lands as 64 rows of 128 bytes whose 32-byte chunks are XOR-swizzled with (row % 4),
// i.e. 4 K-rows x 128 B atoms (SBO = 512 B between 4-pixel groups), 32-channel groups LBO = 8192 B apart.  Only upper-triangular 128x128 tiles are computed; for
// diagonal tiles the B operand aliases the A tile in shared memory (half the smem traffic).
// K (pixels) is split across CTAs so that all 148 SMs stream F once: the C = 64 / 128 layers are
// HBM-bound (intensity C/2 flop/B).  Partials are written to a workspace and summed in a fixed order by
// the finalize kernel (deterministic, no float atomics), which also symmetrises, applies the covariance
// rank-1 correction  sum (x-mu)(y-mu) = sum xy - P mu mu^T  and the 1/(C*P) normalisation.
#include <cstdlib>

#include "gram.cuh"
#include "reduce.cuh"

namespace maua {

namespace {

constexpr int GK = 64;                // pixels per pipeline stage
constexpr int BOX_BYTES = GK * 128;   // one {32 ch, 64 px} box
constexpr int kMaxSplit = 148;

struct GramParams {
    int C;
    long P;
    int T;        // 128-wide tiles per side
    int nsplit;
    float* partial;  // [nsplit][C][C]
};

template <int BN, bool OFFDIAG>
struct GramCfg {
    static constexpr int A_BYTES = 4 * BOX_BYTES;                       // 128 channels
    static constexpr int B_BYTES = OFFDIAG ? (BN / 32) * BOX_BYTES : 0;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int NSTAGES = (192 * 1024) / STAGE_BYTES;          // 6 or 3
    static constexpr int SMEM_BYTES = NSTAGES * STAGE_BYTES + 1024 + 256;
};

// NACC > 1 (default 4): pipeline stage ks accumulates into TMEM accumulator ks % NACC and the epilogue adds
// the NACC accumulators.  The tensor core's fp32 accumulate truncates, so the error of a sum grows linearly with the
// length of the accumulation chain (measured: 2.6e-5 relative at relu1_1 / 1024^2, ~890 dependent MMAs per CTA);
// NACC interleaved chains are NACC times shorter.  This matters for the covariance, where sum(xy) - P mu mu^T cancels.
template <int BN, bool OFFDIAG, int NACC = 1>
__global__ void __launch_bounds__(256, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmF, const GramParams p) {
    static_assert(BN * NACC <= 512, "TMEM has 512 columns");
    using Cfg = GramCfg<BN, OFFDIAG>;
    constexpr int NST = Cfg::NSTAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + NST * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + NST;
    uint64_t* tmem_full_bar = empty_bar + NST;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // upper-triangular tile id -> (ti, tj), ti <= tj
    int ti = 0, tj = 0;
    {
        int id = blockIdx.x;
        for (ti = 0; ti < p.T; ++ti) {
            const int rowlen = p.T - ti;
            if (id < rowlen) { tj = ti + id; break; }
            id -= rowlen;
        }
    }
    const bool diag = (ti == tj);
    const int m0 = ti * 128, n0 = tj * 128;
    const int boxesA = min(4, (p.C - m0) / 32);
    const int boxesB = diag ? 0 : min(BN / 32, (p.C - n0) / 32);

    const long total_st = (p.P + GK - 1) / GK;
    const long st0 = total_st * blockIdx.y / p.nsplit;
    const long st1 = total_st * (blockIdx.y + 1) / p.nsplit;
    const int nk = (int)(st1 - st0);

    if (warp == 0 && lane == 0) tma_prefetch_desc(&tmF);
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 2) { tmem_alloc(tmem_ptr_smem, BN * NACC); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    pdl_wait();
    pdl_trigger();

    // producer / MMA roles run warp-converged; one lane chosen by elect.sync issues (see conv_tc.cu)
    if (warp == 0) {
        for (int ks = 0; ks < nk; ++ks) {
            const int stage = ks % NST;
            const uint32_t phase = (ks / NST) & 1;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
                uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
                uint8_t* sB = sA + Cfg::A_BYTES;
                mbar_arrive_expect_tx(&full_bar[stage], (boxesA + boxesB) * BOX_BYTES);
                const int prow = (int)((st0 + ks) * GK);
                for (int bx = 0; bx < boxesA; ++bx) tma_load_2d(sA + bx * BOX_BYTES, &tmF, &full_bar[stage], m0 + bx * 32, prow);
                for (int bx = 0; bx < boxesB; ++bx) tma_load_2d(sB + bx * BOX_BYTES, &tmF, &full_bar[stage], n0 + bx * 32, prow);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_tf32(128, BN, 1, 1);  // both operands MN-major
        for (int ks = 0; ks < nk; ++ks) {
            const int stage = ks % NST;
            const uint32_t phase = (ks / NST) & 1;
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                const uint32_t sB = (OFFDIAG && !diag) ? sA + Cfg::A_BYTES : sA;
#pragma unroll
                for (int kk = 0; kk < GK / 8; ++kk) {
                    // MN-major TF32: layout type 1 (128B swizzle, 32-byte atoms): 4-row x 128 B atoms, SBO = 512 B
                    const uint64_t adesc = make_smem_desc(sA + kk * 1024, BOX_BYTES, 512, 1);
                    const uint64_t bdesc = make_smem_desc(sB + kk * 1024, BOX_BYTES, 512, 1);
                    umma_tf32(tmem_base + (NACC > 1 ? (ks % NACC) * BN : 0), adesc, bdesc, idesc,
                              (ks >= NACC || kk > 0) ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
                if (ks == nk - 1) umma_commit(tmem_full_bar);
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int c = m0 + row;
        float* dst = p.partial + ((size_t)blockIdx.y * p.C + c) * p.C + n0;
        if (nk > 0) {
            mbar_wait(tmem_full_bar, 0);
            tc_fence_after();
#pragma unroll 1
            for (int col = 0; col < BN; col += 16) {
                float v[16];
                tmem_ld_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + col, v);
                if (NACC > 1) {
                    const int used = nk < NACC ? nk : NACC;  // accumulators that received at least one stage
#pragma unroll 1
                    for (int a = 1; a < used; ++a) {
                        float u[16];
                        tmem_ld_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + a * BN + col, u);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += u[i];
                    }
                }
                if (c < p.C && n0 + col < p.C) {  // (C % 128 == 64: the last tile row / column is half a tile)
#pragma unroll
                    for (int i = 0; i < 16; i += 4)
                        *reinterpret_cast<float4*>(dst + col + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
        } else if (c < p.C) {
            for (int col = 0; col < BN && n0 + col < p.C; col += 4) *reinterpret_cast<float4*>(dst + col) = make_float4(0, 0, 0, 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, BN * NACC);
}

// naive SIMT cross-check: partial slot 0 = full (upper + lower) raw Gram
__global__ void gram_ref_kernel(const float* __restrict__ f, long P, int C, float* __restrict__ out) {
    __shared__ float sa[16][17], sb[16][17];
    const int c = blockIdx.y * 16 + threadIdx.y, d = blockIdx.x * 16 + threadIdx.x;
    float acc = 0.f;
    for (long p0 = 0; p0 < P; p0 += 16) {
        const long pa = p0 + threadIdx.x;
        sa[threadIdx.y][threadIdx.x] = (pa < P && c < C) ? f[pa * C + c] : 0.f;                          // [c][p]
        const long pb = p0 + threadIdx.y;
        sb[threadIdx.y][threadIdx.x] = (pb < P && d < C) ? f[pb * C + d] : 0.f;                          // [p][d]
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = fmaf(sa[threadIdx.y][k], sb[k][threadIdx.x], acc);
        __syncthreads();
    }
    if (c < C && d < C) out[(size_t)c * C + d] = acc;
}

// Exact-arithmetic Gram / covariance (MAUA_IMPL_FP32): operands are the stored fp32 features, centred on load like the
// reference does (loss.py:87-89: x - x.mean) when `mean` is given; products of 16 pixels are summed with FFMA in fp32 and
// those chunk sums are added in fp64.  Pixels are split over blockIdx.z; the fp64 partials are added in a fixed order.
__global__ void gram_exact_kernel(const float* __restrict__ f, long P, int C, const float* __restrict__ mean,
                                  double* __restrict__ dpartial) {
    __shared__ float sa[16][17], sb[16][17];
    const int c = blockIdx.y * 16 + threadIdx.y, d = blockIdx.x * 16 + threadIdx.x;
    const int ca = blockIdx.y * 16 + threadIdx.x;  // channel this thread loads for the A tile ([p][c], coalesced over c)
    const float mu_a = (mean && ca < C) ? mean[ca] : 0.f;
    const float mu_b = (mean && d < C) ? mean[d] : 0.f;
    const long p_begin = P * blockIdx.z / gridDim.z, p_end = P * (blockIdx.z + 1) / gridDim.z;
    double acc = 0.0;
    for (long p0 = p_begin; p0 < p_end; p0 += 16) {
        const long pp = p0 + threadIdx.y;
        const bool live = pp < p_end;
        sa[threadIdx.y][threadIdx.x] = (live && ca < C) ? f[pp * C + ca] - mu_a : 0.f;  // [p][c]
        sb[threadIdx.y][threadIdx.x] = (live && d < C) ? f[pp * C + d] - mu_b : 0.f;    // [p][d]
        __syncthreads();
        float part = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) part = fmaf(sa[k][threadIdx.y], sb[k][threadIdx.x], part);
        acc += (double)part;
        __syncthreads();
    }
    if (c < C && d < C) dpartial[((size_t)blockIdx.z * C + c) * C + d] = acc;
}
__global__ void gram_exact_sum_kernel(const double* __restrict__ dpartial, int S, long total, float* __restrict__ out) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int k = 0; k < S; ++k) s += dpartial[(size_t)k * total + i];
        out[i] = (float)s;
    }
}

// Sums the split-K partials in a fixed order (deterministic), mirrors the upper-triangular tiles into the lower ones,
// applies the covariance correction and the 1/(C P) normalisation.  Block = (256 / G) consecutive outputs x G interleaved
// groups of partials (G grows with the split count so that short and long reductions both keep the loads coalesced and
// the threads busy).  Optionally fused with the StyleLoss value (loss.py:153-157): diff = gram - target,
// *loss_out = scale * mean(diff^2).
// (min 4 blocks per SM: without the bound ptxas unrolls the partial-sum loop into 128 registers -- 2 blocks per SM, 24 % occupancy --
// and the kernel, which is all load latency, took 26 us for a 512-channel layer)
template <int G>
__global__ void __launch_bounds__(kReduceThreads, 4)
gram_finalize_kernel(const float* __restrict__ partial, int nsplit, int C, int Cn, long P, int full, const float* __restrict__ mean,
                     float* __restrict__ gram, const float* __restrict__ target, float* __restrict__ diff, float scale,
                     float* __restrict__ loss_out, double* red_partials, unsigned int* counter) {
    constexpr int NO = kReduceThreads / G;  // outputs per block round
    __shared__ float sh[G][NO + 1];
    pdl_wait();
    pdl_trigger();
    const int tx = threadIdx.x % NO, ty = threadIdx.x / NO;
    const long total = (long)C * C;
    double acc = 0.0;
    for (long base = (long)blockIdx.x * NO; base < total; base += (long)gridDim.x * NO) {
        const long i = base + tx;
        const int c = (int)(i / C), d = (int)(i % C);
        // only the upper-triangular 128x128 tiles were computed; the source read stays coalesced, the mirror is a write
        const bool live = i < total && (full || ((c >> 7) <= (d >> 7)));
        float s = 0.f;
        if (live)
            for (int k = ty; k < nsplit; k += G) s += partial[(size_t)k * total + i];
        if (G > 1) {
            sh[ty][tx] = s;
            __syncthreads();
        }
        if (ty == 0 && live) {
            float t = s;
            if (G > 1) {
                t = 0.f;
#pragma unroll
                for (int y = 0; y < G; ++y) t += sh[y][tx];
            }
            if (mean) t -= (float)P * mean[c] * mean[d];
            const float g = t / ((float)Cn * (float)P);  // Cn < C: zero-padded channels (pruned VGG-16), normalised as the real count
            const bool mirror = !full && ((c >> 7) != (d >> 7));
            const long im = (long)d * C + c;
            gram[i] = g;
            if (mirror) gram[im] = g;
            if (target) {
                const float dd = g - target[i];
                diff[i] = dd;
                acc += (double)dd * dd;
                if (mirror) {
                    const float dm = g - target[im];
                    diff[im] = dm;
                    acc += (double)dm * dm;
                }
            }
        }
        if (G > 1) __syncthreads();
    }
    if (target) {
        double v[1] = {acc}, tot[1];
        if (grid_sum<1>(v, red_partials, counter, tot)) *loss_out = scale * (float)(tot[0] / ((double)Cn * (double)Cn));
    }
}

__global__ void style_loss_fwd_kernel(const float* __restrict__ gram, const float* __restrict__ target, long n,
                                      float scale, float* __restrict__ loss_out, float* __restrict__ diff,
                                      double* partials, unsigned int* counter) {
    double acc = 0.0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float dd = gram[i] - target[i];
        diff[i] = dd;
        acc += (double)dd * dd;
    }
    double v[1] = {acc}, tot[1];
    if (grid_sum<1>(v, partials, counter, tot)) *loss_out = scale * (float)(tot[0] / (double)n);
}

// aux_d = coef * 4 / (C^3 P) * diff  (TF32-rounded: it is the B operand of the backward MMA)
__global__ void style_bwd_prep_kernel(const float* __restrict__ diff, int C, long P, const float* __restrict__ coef,
                                      float* __restrict__ aux_d) {
    const long total = (long)C * C;
    const float k = *coef * 4.f / ((float)C * (float)C * (float)C * (float)P);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
        aux_d[i] = round_tf32(k * diff[i]);
}
// aux_bias[c] = - sum_d aux_d[c][d] * mean[d]
__global__ void style_bwd_bias_kernel(const float* __restrict__ aux_d, const float* __restrict__ mean, int C,
                                      float* __restrict__ aux_bias) {
    const int c = blockIdx.x;
    float acc = 0.f;
    for (int d = threadIdx.x; d < C; d += blockDim.x) acc += aux_d[(size_t)c * C + d] * mean[d];
    acc = warp_sum(acc);
    __shared__ float sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i];
        aux_bias[c] = -s;
    }
}

__global__ void axpby_kernel(const float* __restrict__ x, float* __restrict__ y, long n, float a, int accumulate) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        y[i] = (accumulate ? y[i] : 0.f) + a * x[i];
}

template <int BN, bool OFFDIAG, int NACC>
int launch_gram_nacc(const CUtensorMap& tm, const GramParams& p, int ntiles, cudaStream_t st) {
    using Cfg = GramCfg<BN, OFFDIAG>;
    static unsigned long long attr_done = 0;
    MAUA_CUDA_CHECK((ensure_dynamic_smem(gram_tc_kernel<BN, OFFDIAG, NACC>, Cfg::SMEM_BYTES, &attr_done)));
    MAUA_CUDA_CHECK(launch_pdl<PDL_GRAM>(gram_tc_kernel<BN, OFFDIAG, NACC>, dim3(ntiles, p.nsplit), dim3(256), Cfg::SMEM_BYTES, st, tm, p));
    return MAUA_OK;
}

template <int BN, bool OFFDIAG>
int launch_gram_tc(const CUtensorMap& tm, const GramParams& p, int ntiles, cudaStream_t st) {
    // four interleaved accumulation chains by default (see gram_tc_kernel; verified on B200 in round 2: covariance error
    // vs fp64 1.0e-3 -> 2.2e-4 at 1024^2, same run time); MAUA_GRAM_NACC=1 selects the single chain
    const char* f = getenv("MAUA_GRAM_NACC");
    if (f && atoi(f) == 1) return launch_gram_nacc<BN, OFFDIAG, 1>(tm, p, ntiles, st);
    return launch_gram_nacc<BN, OFFDIAG, 4>(tm, p, ntiles, st);
}

}  // namespace

// [C][C] float slots of the partial area: the split-K partials of the tensor-core path (kMaxSplit / ntiles of them), and at
// least 3 so that the exact-arithmetic path always has room for one fp64 partial (2 slots) + the fp32 sums (1 slot)
static int gram_slots(int C) {
    const int T = (C + 127) / 128;
    const int ntiles = T * (T + 1) / 2;
    const int s = kMaxSplit / ntiles;
    return s > 3 ? s : 3;
}

size_t gram_workspace_bytes(int C) {
    // [slots][C][C] floats + channel-mean scratch (doubles)
    return (((size_t)gram_slots(C) * C * C * sizeof(float) + 255) & ~size_t(255)) + (size_t)kMaxSplit * 4 * C * sizeof(double) + 1024;
}

int gram_launch(const float* f, long P, int C, int use_cov, float* gram, float* mean_out, void* workspace, int impl,
                cudaStream_t st, const GramLossFuse* fuse, int Cn) {
    if (Cn <= 0) Cn = C;
    // (the [B*C, B*C] dynamic Gram of an img_vid window, loss.py:164-168, is this kernel on the B*C-channel matrix)
    MAUA_REQUIRE(C >= 64 && C % 64 == 0 && C <= 16384,
                 "gram: channel count %d unsupported (need a multiple of 64, <= 16384)", C);
    MAUA_REQUIRE(P >= 1 && P < (1L << 31), "gram: bad pixel count %ld", P);
    const int T = (C + 127) / 128;
    const int ntiles = T * (T + 1) / 2;
    int nsplit = kMaxSplit / ntiles > 0 ? kMaxSplit / ntiles : 1;
    const long total_st = (P + GK - 1) / GK;
    if (nsplit > total_st) nsplit = (int)total_st;
    float* partial = reinterpret_cast<float*>(workspace);
    double* mean_scratch = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) +
                                                     (((size_t)gram_slots(C) * C * C * sizeof(float) + 255) & ~size_t(255)));
    if (use_cov) {
        int rc = channel_mean_launch(f, P, C, mean_out, mean_scratch, kMaxSplit * 4, st);
        if (rc) return rc;
    }
    int full = 0;
    bool centred = false;
    if (impl == 4) {
        // workspace: fp64 partials in the first float slots, the fp32 sums in the last float slot of the partial area
        const int slots = gram_slots(C);
        int S = (slots - 1) / 2;
        if ((long)S > (P + 255) / 256) S = (int)((P + 255) / 256);
        double* dpartial = reinterpret_cast<double*>(workspace);
        const long cc = (long)C * C;
        float* sums = partial + (size_t)(slots - 1) * cc;
        gram_exact_kernel<<<dim3((C + 15) / 16, (C + 15) / 16, S), dim3(16, 16), 0, st>>>(f, P, C, use_cov ? mean_out : nullptr,
                                                                                         dpartial);
        MAUA_CUDA_CHECK(cudaGetLastError());
        gram_exact_sum_kernel<<<(int)((cc + 255) / 256 > 592 ? 592 : (cc + 255) / 256), 256, 0, st>>>(dpartial, S, cc, sums);
        MAUA_CUDA_CHECK(cudaGetLastError());
        partial = sums;
        nsplit = 1;
        full = 1;
        centred = true;
    } else if (impl == 1) {
        nsplit = 1;
        full = 1;
        gram_ref_kernel<<<dim3((C + 15) / 16, (C + 15) / 16), dim3(16, 16), 0, st>>>(f, P, C, partial);
        MAUA_CUDA_CHECK(cudaGetLastError());
    } else {
        CUtensorMap tm;
        int rc = make_tmap_2d(&tm, f, P, C, GK, 1);
        if (rc) return rc;
        GramParams p;
        p.C = C; p.P = P; p.T = T; p.nsplit = nsplit; p.partial = partial;
        if (C == 64) rc = launch_gram_tc<64, false>(tm, p, ntiles, st);
        else if (C == 128) rc = launch_gram_tc<128, false>(tm, p, ntiles, st);
        else rc = launch_gram_tc<128, true>(tm, p, ntiles, st);
        if (rc) return rc;
    }
    const long total = (long)C * C;
    const int G = nsplit >= 64 ? 8 : (nsplit >= 24 ? 4 : (nsplit >= 8 ? 2 : 1));
    const long per_block = kReduceThreads / G;
    const int fgrid = (int)((total + per_block - 1) / per_block > 148 * 8 ? 148 * 8 : (total + per_block - 1) / per_block);
    const float* tgt = nullptr; float* dif = nullptr; float* lout = nullptr; float sc = 0.f;
    double* rp = nullptr; unsigned int* cnt = nullptr;
    if (fuse) {
        MAUA_REQUIRE(fuse->target && fuse->diff && fuse->loss_out && fgrid <= fuse->rs.max_blocks, "gram: bad fused-loss arguments");
        tgt = fuse->target; dif = fuse->diff; lout = fuse->loss_out; sc = fuse->value_scale;
        rp = fuse->rs.partials; cnt = fuse->rs.counter;
    }
    const float* mu = (use_cov && !centred) ? mean_out : nullptr;
#define MAUA_FINALIZE(GG)                                                                                                          \
    MAUA_CUDA_CHECK(launch_pdl<PDL_GRAM>(gram_finalize_kernel<GG>, dim3(fgrid), dim3(kReduceThreads), 0, st, (const float*)partial, nsplit, C, Cn, P, \
                               full, mu, gram, tgt, dif, sc, lout, rp, cnt))
    if (G == 8) MAUA_FINALIZE(8); else if (G == 4) MAUA_FINALIZE(4); else if (G == 2) MAUA_FINALIZE(2); else MAUA_FINALIZE(1);
#undef MAUA_FINALIZE
    return MAUA_OK;
}

int style_loss_fwd_launch(const float* gram, const float* target, int C, float value_scale, float* loss_out,
                          float* diff, ReduceScratch rs, cudaStream_t st) {
    const long n = (long)C * C;
    int grid = (int)((n + 255) / 256);
    if (grid > 148) grid = 148;
    MAUA_REQUIRE(grid <= rs.max_blocks, "reduce scratch too small");
    style_loss_fwd_kernel<<<grid, 256, 0, st>>>(gram, target, n, value_scale, loss_out, diff, rs.partials, rs.counter);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

int style_loss_bwd_prep_launch(const float* diff, const float* mean, int C, long P, const float* coef, float* aux_d,
                               float* aux_bias, cudaStream_t st) {
    const long n = (long)C * C;
    style_bwd_prep_kernel<<<(int)((n + 255) / 256 > 592 ? 592 : (n + 255) / 256), 256, 0, st>>>(diff, C, P, coef, aux_d);
    MAUA_CUDA_CHECK(cudaGetLastError());
    if (mean) {
        style_bwd_bias_kernel<<<C, 128, 0, st>>>(aux_d, mean, C, aux_bias);
        MAUA_CUDA_CHECK(cudaGetLastError());
    }
    return MAUA_OK;
}

int style_loss_bwd_bias_launch(const float* aux_d, const float* mean, int C, float* aux_bias, cudaStream_t st) {
    style_bwd_bias_kernel<<<C, 128, 0, st>>>(aux_d, mean, C, aux_bias);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

int axpby_launch(const float* x, float* y, long n, float a, int accumulate, cudaStream_t st) {
    axpby_kernel<<<(int)((n + 255) / 256 > 592 ? 592 : (n + 255) / 256), 256, 0, st>>>(x, y, n, a, accumulate);
    MAUA_CUDA_CHECK(cudaGetLastError());
    return MAUA_OK;
}

}  // namespace maua
