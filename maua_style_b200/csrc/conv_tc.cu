// conv3x3 (pad 1, stride 1) as a TMA-fed tcgen05 / TMEM implicit GEMM for sm_100a.
//
// Replaces cuDNN fprop / bwd-data behind reference models.py:129 (nn.Conv2d in build_sequential)
// and the autograd dgrad of it (optim.py:213 total_loss.backward()).  Weights are frozen
// (models.py:443-445) so only input-gradients are ever needed: the backward pass is this same
// kernel run on 180-degree-rotated, channel-transposed weights.
//
// GEMM view (NHWC activations):   D[pixel][n] = sum_{tap, c} A_tap[pixel][c] * Wg[n][tap*Cin + c]
//   M = pixels.  One CTA owns MT sub-tiles of 128 pixels (16 wide x 8 high), stacked along h.
//   N = output channels, BN in {32,64,128,256} per CTA.
//   K = 9 taps x Cin, walked in k-steps of 32 channels (= one 128-byte swizzle span of fp32).
// A operand: for every (dx, channel chunk) ONE 4-D TMA box {32 c, 16 w, MT*8 + 2 h, 1 b} whose origin is shifted by
//   (dx, -1); out-of-bounds pixels are zero-filled by TMA, which is exactly the conv zero padding and also handles
//   ragged right/bottom tiles for arbitrary H x W.  The three vertical taps dy = -1, 0, +1 re-use that box: a shift by
//   one image row is 16 pixel-rows x 128 B = 2048 B, a multiple of the 1024-byte swizzle atom, so the UMMA descriptor
//   simply starts 0 / 2048 / 4096 bytes into the box.  This cuts the L2 -> shared-memory traffic of the activations
//   2.7x compared with one box per tap (horizontal shifts would break the swizzle phase, so dx stays a reload).
// B operand: 2-D TMA box {32 k, BN n} of the K-major GEMM weights, one per (tap, chunk), in its own ring.
// Both land in the canonical K-major SWIZZLE_128B layout (8-row x 128-byte atoms, SBO = 1024 B).
// Accumulators: MT x [128 lanes x BN fp32 columns] in TMEM, double-buffered when they fit (<= 256 columns).
// An optional second ("aux") GEMM term sum_{c2} A2[pixel][c2] * W2[n][c2] accumulates into the
// same TMEM tile as extra k-steps: this is the StyleLoss backward 4(G-A)F/(C^3 N) (loss.py:141-157)
// folded into the dgrad of the following layer.
//
// Warp roles (384 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane),
// warp 2 = TMEM allocator, warp 3 = L2 prefetch of the next layer's weights, warps 4-7 and 8-11 = two epilogue groups (TMEM ->
// registers -> fused epilogue -> global) that take alternate 32-column boxes of a tile.
#include "conv_tc.cuh"
#include "pointwise.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace maua {

namespace {

constexpr int TILE_W = 16;
constexpr int TILE_H = 8;   // one 128-pixel sub-tile = 16 x 8
constexpr int KCHUNK = 32;  // fp32 channels per k-step (128 bytes)
constexpr int A_BYTES = 128 * 128;          // one sub-tile: 128 pixel rows x 128 B
constexpr int ROW_BYTES = TILE_W * 128;     // one image row of the tile: 16 pixels x 128 B = 2 swizzle atoms

struct ConvKParams {
    int B, H, W, Cin, Cout, ntaps, K2;
    int tiles_w, tiles_h, n_tiles, total_tiles;
    // Work items: tiles [0, full_tiles) are computed whole; each of the remaining total_tiles - full_tiles tiles (the last,
    // partial wave of the persistent grid) is split along K into `split` parts that run on different CTAs (pairs) at the
    // same time.  Parts 0 .. split-2 dump their raw accumulators into splitk_ws; the last part adds them (fixed order:
    // deterministic) and runs the fused epilogue.
    int full_tiles, split, total_items;
    int tail_halves;  // 1: the tail tiles are computed as two half-N items each (no hand-over) instead of K-split parts
    int split_nowait; // 1 (tail mode 3, only when EVERY tile is split): all parts dump their raw accumulators and return; the
                      // sums and the fused epilogue are done by conv_splitk_reduce_kernel launched right behind (no flags, no wait)
    float* splitk_ws;
    unsigned int* splitk_flags;
    const uint8_t* prefetch;        // next layer's weights: L2 prefetch by the spare warp (ConvArgs::prefetch)
    unsigned long long prefetch_bytes;
    int direct;  // 1: register -> global epilogue (needed for the content / addend / fp32-mask terms), 0: TMA-store epilogue
    ConvEpilogue ep;
};

// Threads: 4 role warps + TWO groups of 4 epilogue warps.  One epilogue warp per scheduler issues ~340 dependent instructions per
// 32-column box at ~6 cycles each (~2200 cycles per box, measured with the stores switched off): for the short-K layers
// (Cin = 64: 6 k-groups per tile) that was longer than the tile's MMAs and sat on the critical path (conv2_1 forward 72.7 us, of
// which 42.7 without the epilogue).  Two warps per scheduler, working on alternate boxes, hide each other's latencies.
constexpr int kConvThreads = 384;
constexpr int STAGE_BOX_BYTES = 128 * 128;  // epilogue staging: one {32 ch, 16 w, 8 h} output box per epilogue group
constexpr int N_STAGE_BOX = 2;

// KS: kernel size of the halo main term -- 3 (VGG, NIN conv3 / conv4) or 5 (NIN conv2, models.py:90: 5x5 / pad 2): the same
// pipeline with a halo of KS / 2 pixels, KS horizontal box loads per channel chunk and KS vertical taps per box.
template <int BN, int MT, int CG, int KS = 3>
struct ConvCfg {
    static constexpr int BNL = BN / CG;                            // weight rows this CTA keeps (half the tile in a pair)
    static constexpr int A_ROWS = MT * TILE_H + KS - 1;            // image rows incl. the vertical halo
    static constexpr int A_STAGE = A_ROWS * ROW_BYTES;             // 36864 (MT=2) / 20480 (MT=1): multiples of 1024
    static constexpr int B_STAGE = BNL * 128;
    static constexpr int BUDGET = 192 * 1024;                      // operand rings; 32 KB more go to the epilogue staging
    // Pipeline shape.  Every barrier hand-over costs the issuing threads a few hundred cycles (try_wait + elect + tcgen05.commit:
    // measured with the loads and the MMAs switched off, tools/exp_conv_chain.py: ~440 cycles per weight step), which a
    // weight step of a narrow tile (4 * MT MMAs of N <= 128: 64..512 tensor cycles) does not cover -- the 64-channel layers
    // ran at <= 50 % tensor-pipe-active and every launch of the small scales was a chain of 3 * 48 such steps.  Narrow shapes
    // therefore use MERGED stages: one activation box + its KS weight tiles behind ONE full / ONE empty barrier (one wait and
    // one commit per k-group instead of KS waits and KS + 1 commits), as many stages as fit (>= 3, <= 8).  Wide shapes, whose
    // weight tiles are too big for three such stages and whose steps are long enough, keep separate activation / weight rings.
    static constexpr int STAGE = A_STAGE + KS * B_STAGE;
    static constexpr int NS_RAW = BUDGET / STAGE;
    static constexpr bool MERGED = NS_RAW >= 3;
    static constexpr int NA = MERGED ? (NS_RAW > 8 ? 8 : NS_RAW) : ((3 * A_STAGE + 4 * B_STAGE <= BUDGET) ? 3 : 2);
    static constexpr int NB_RAW = (BUDGET - NA * A_STAGE) / B_STAGE;
    static constexpr int NB = MERGED ? 0 : (NB_RAW > 8 ? 8 : NB_RAW);
    static constexpr int A_PITCH = MERGED ? STAGE : A_STAGE;      // distance between the activation boxes of two stages
    static constexpr int ACC_COLS = MT * BN;                       // one accumulator set: 32..512 columns
    static constexpr int NACC = ACC_COLS <= 256 ? 2 : 1;           // double-buffered when TMEM has room
    static constexpr int TMEM_COLS_RAW = NACC * ACC_COLS;
    static constexpr int TMEM_COLS = TMEM_COLS_RAW < 32 ? 32 : TMEM_COLS_RAW;  // power of two, 32..512
    static constexpr int RING_BYTES = NA * A_PITCH + NB * B_STAGE;
    static constexpr int SMEM_BYTES = RING_BYTES + 1024 /*align slack*/ + 1024 /*barriers*/ + N_STAGE_BOX * STAGE_BOX_BYTES;
    static_assert(MERGED || NB >= 3, "weight ring too shallow");
    static_assert(A_PITCH % 1024 == 0 && B_STAGE % 1024 == 0, "operand tiles must stay 1024-byte aligned (128B swizzle atoms)");
    static_assert(CG == 1 || CG == 2, "CTA group size");
    static_assert(BNL % 16 == 0 && BNL >= 16, "per-CTA weight rows");
};

// Persistent, warp-specialised kernel: each CTA walks tiles  t = blockIdx.x, blockIdx.x + gridDim.x, ...
// Tile index = pixel_tile * n_tiles + n_tile, so CTAs that run concurrently share the activation tile in L2.
// Pipelines: activation ring (TMA -> MMA), weight ring (TMA -> MMA), TMEM accumulators (MMA -> epilogue,
// double-buffered when they fit) and the tile loop itself, so the epilogue of tile i overlaps the loads and MMAs of
// tile i+1.
//
// K is walked in groups.  3x3 main term: group = (dx, channel chunk), one halo box + 3 weight tiles (dy = -1, 0, 1).
// Pointwise main term (ntaps == 1) and the aux term: group = channel chunk, one plain box + 1 weight tile.
//
// CG = 2 runs the same pipeline on CTA PAIRS (2-CTA clusters, tcgen05 cta_group::2): a pair owns a tile of 2*MT sub-tiles
// (this CTA the ones at rows h0 + rank*MT*8 ...), every MMA is M = 256 across both CTAs, and each CTA loads its own
// activation boxes but only HALF of every weight tile.  All TMA loads complete on the even CTA's barriers (its producer
// arms them for both CTAs' bytes), the even CTA's MMA warp issues for the pair and its commits arrive in both CTAs.
template <int BN, int MT, int CG, bool POOL = false, int KS = 3>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
               const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
               const ConvKParams p) {
    using Cfg = ConvCfg<BN, MT, CG, KS>;
    constexpr int NA = Cfg::NA, NB = Cfg::NB, NACC = Cfg::NACC;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;  // position in the CTA pair; rank 0 leads
    const int cta_tile0 = blockIdx.x / CG, cta_tile_step = gridDim.x / CG;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smemA = smem;
    uint8_t* smemB = smem + NA * Cfg::A_PITCH;  // (separate weight ring; merged stages keep their weight tiles behind the box)
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + Cfg::RING_BYTES);
    uint64_t* a_empty = a_full + NA;
    uint64_t* b_full = a_empty + NA;
    uint64_t* b_empty = b_full + NB;
    uint64_t* tmem_full_bar = b_empty + NB;        // [NACC]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [NACC]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    uint8_t* stage_box = smem + Cfg::RING_BYTES + 1024;  // 1024-byte aligned (ring sizes are multiples of 1024)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int cpt = p.Cin / KCHUNK;                         // channel chunks of the main term
    const bool halo = (p.ntaps == KS * KS);
    const int ng1 = halo ? KS * cpt : (p.ntaps == 1 ? cpt : 0);  // main groups
    const int ng2 = p.K2 / KCHUNK;                          // aux groups
    const int ng = ng1 + ng2;

    struct Item { int tile, g0, g1, part, nparts, slot, nh; };
    auto get_item = [&](int it) {
        Item o;
        o.nh = -1;
        if (it < p.full_tiles) {
            o.tile = it; o.g0 = 0; o.g1 = ng; o.part = 0; o.nparts = 1; o.slot = 0;
        } else if (p.tail_halves) {
            // tail tile r / 2, output-channel half r % 2: the whole K range, half of the tile's channels -- two independent
            // items on two CTA units, nothing to hand over
            const int r = it - p.full_tiles;
            o.tile = p.full_tiles + (r >> 1);
            o.nh = r & 1;
            o.g0 = 0; o.g1 = ng; o.part = 0; o.nparts = 1; o.slot = 0;
        } else {
            const int r = it - p.full_tiles;
            o.slot = r / p.split;
            o.part = r - o.slot * p.split;
            o.nparts = p.split;
            o.tile = p.full_tiles + o.slot;
            o.g0 = ng * o.part / p.split;
            o.g1 = ng * (o.part + 1) / p.split;
        }
        return o;
    };

    // Output channel (relative to the tile's first channel) of accumulator column c.  A half-N item of a CTA pair uses the first
    // or second quarter of EACH CTA's weight rows (the pair's MMA takes N/2 rows from either CTA), i.e. two channel runs.
    auto chan = [&](int c, int nh) -> int {
        if (nh < 0) return c;
        if (CG == 2) return (c < BN / 4 ? c : c - BN / 4 + BN / 2) + nh * (BN / 4);
        return c + nh * (BN / 2);
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (ng2 > 0) {
            tma_prefetch_desc(&tmA2);
            tma_prefetch_desc(&tmB2);
        }
        if (!p.direct) tma_prefetch_desc(&tmOut);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NA; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < NACC; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], 8 * CG);  // one arrival per epilogue warp (of both CTAs of a pair)
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (CG == 2) { tmem_alloc_2sm(tmem_ptr_smem, Cfg::TMEM_COLS); tmem_relinquish_2sm(); }
        else { tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS); tmem_relinquish(); }
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();  // the peer's barriers must be initialised before any TMA / commit targets them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    if (warp == 3 && p.prefetch_bytes) {
        // frozen weights of the next launch -> L2, in 16 KB pieces spread over the grid (independent of the previous kernel: no wait)
        constexpr unsigned long long kPiece = 16384;
        const unsigned long long npieces = (p.prefetch_bytes + kPiece - 1) / kPiece;
        for (unsigned long long i = blockIdx.x * 32ull + lane; i < npieces; i += gridDim.x * 32ull) {
            const unsigned long long off = i * kPiece;
            const unsigned long long left = p.prefetch_bytes - off;
            l2_prefetch_bulk(p.prefetch + off, static_cast<uint32_t>(left < kPiece ? left : kPiece));
        }
    }
    pdl_wait();     // everything above (descriptor prefetch, barrier init, TMEM allocation) overlapped the previous kernel's tail
    pdl_trigger();

    auto decode = [&](int tile, int& b, int& h0, int& w0, int& n0) {
        const int nt = tile % p.n_tiles;
        int pt = tile / p.n_tiles;
        const int tw = pt % p.tiles_w;
        pt /= p.tiles_w;
        const int th = pt % p.tiles_h;
        b = pt / p.tiles_h;
        w0 = tw * TILE_W;
        h0 = th * (TILE_H * MT * CG) + (int)rank * (TILE_H * MT);  // this CTA's rows of the (pair) tile
        n0 = nt * BN;
    };
    const int nb0 = (int)rank * Cfg::BNL;  // this CTA's rows of every weight tile

    // The producer and MMA roles run with the whole warp converged; a single lane chosen by elect.sync issues the TMA /
    // tcgen05 instructions.  (Entering these loops with only lane 0 active makes ptxas wrap every UTCHMMA in an
    // ELECT / BRA.U.ANY loop, which makes the short N = 64 MMAs issue-bound.)
    if (warp == 0) {
        // ===================== TMA producer =====================
        // In a pair both CTAs load (own activations, own half of the weights) but only the even CTA arms the barriers,
        // for the bytes of both.
        auto arm = [&](uint64_t* bar, uint32_t bytes) { if (rank == 0) mbar_arrive_expect_tx(bar, bytes * CG); };
        auto load_a = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
            if (CG == 2) tma_load_4d_2sm(dst, m, bar, c0, c1, c2, c3); else tma_load_4d(dst, m, bar, c0, c1, c2, c3);
        };
        auto load_b = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
            if (CG == 2) tma_load_2d_2sm(dst, m, bar, c0, c1); else tma_load_2d(dst, m, bar, c0, c1);
        };
        uint32_t ia = 0, ib = 0;  // ring counters across tiles
        for (int it = cta_tile0; it < p.total_items; it += cta_tile_step) {
            const Item item = get_item(it);
            int b, h0, w0, n0;
            decode(item.tile, b, h0, w0, n0);
            for (int g = item.g0; g < item.g1; ++g) {
                const int sa = ia % NA;
                mbar_wait(&a_empty[sa], ((ia / NA) & 1) ^ 1);
                ++ia;
                uint8_t* sA = smemA + sa * Cfg::A_PITCH;
                const bool is_halo = g < ng1 && halo;
                const bool aux = g >= ng1;
                const int dxi = is_halo ? g / cpt : 0;        // 0..KS-1  <->  dx = -KS/2 .. +KS/2
                const int c0 = (is_halo ? g - dxi * cpt : (aux ? g - ng1 : g)) * KCHUNK;
                if constexpr (Cfg::MERGED) {
                    // one stage = the activation box + its weight tiles, all against one barrier
                    if (elect_one()) {
                        uint8_t* sB = sA + Cfg::A_STAGE;
                        if (is_halo) {
                            arm(&a_full[sa], Cfg::A_STAGE + KS * Cfg::B_STAGE);
                            load_a(sA, &tmA, &a_full[sa], c0, w0 + dxi - KS / 2, h0 - KS / 2, b);
#pragma unroll
                            for (int dyi = 0; dyi < KS; ++dyi)
                                load_b(sB + dyi * Cfg::B_STAGE, &tmB, &a_full[sa], (dyi * KS + dxi) * p.Cin + c0, n0 + nb0);
                        } else {
                            arm(&a_full[sa], MT * A_BYTES + Cfg::B_STAGE);
                            load_a(sA, aux ? &tmA2 : &tmA, &a_full[sa], c0, w0, h0, b);
                            load_b(sB, aux ? &tmB2 : &tmB, &a_full[sa], c0, n0 + nb0);
                        }
                    }
                    __syncwarp();
                } else if (is_halo) {
                    if (elect_one()) {
                        arm(&a_full[sa], Cfg::A_STAGE);
                        load_a(sA, &tmA, &a_full[sa], c0, w0 + dxi - KS / 2, h0 - KS / 2, b);
                    }
                    __syncwarp();
                    for (int dyi = 0; dyi < KS; ++dyi) {
                        const int sb = ib % NB;
                        mbar_wait(&b_empty[sb], ((ib / NB) & 1) ^ 1);
                        ++ib;
                        if (elect_one()) {
                            arm(&b_full[sb], Cfg::B_STAGE);
                            load_b(smemB + sb * Cfg::B_STAGE, &tmB, &b_full[sb], (dyi * KS + dxi) * p.Cin + c0, n0 + nb0);
                        }
                        __syncwarp();
                    }
                } else {
                    const int sb = ib % NB;
                    mbar_wait(&b_empty[sb], ((ib / NB) & 1) ^ 1);
                    ++ib;
                    if (elect_one()) {
                        arm(&a_full[sa], MT * A_BYTES);
                        load_a(sA, aux ? &tmA2 : &tmA, &a_full[sa], c0, w0, h0, b);
                        arm(&b_full[sb], Cfg::B_STAGE);
                        load_b(smemB + sb * Cfg::B_STAGE, aux ? &tmB2 : &tmB, &b_full[sb], c0, n0 + nb0);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (the even CTA issues for the pair) =====================
        constexpr uint32_t idesc_full = make_idesc_tf32(128 * CG, BN, 0, 0);
        constexpr uint32_t idesc_half = make_idesc_tf32(128 * CG, BN >= 32 ? BN / 2 : BN, 0, 0);
        auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accumulate) {
            if (CG == 2) umma_tf32_2sm(d, ad, bd, idesc, accumulate); else umma_tf32(d, ad, bd, idesc, accumulate);
        };
        auto commit = [&](uint64_t* bar) { if (CG == 2) umma_commit_2sm(bar); else umma_commit(bar); };
        uint32_t ia = 0, ib = 0, lt = 0;
        for (int it = cta_tile0; it < p.total_items; it += cta_tile_step, ++lt) {
            const Item item = get_item(it);
            const int acc = lt % NACC;
            mbar_wait(&tmem_empty_bar[acc], ((lt / NACC) & 1) ^ 1);  // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_base = tmem_base + acc * Cfg::ACC_COLS;
            const uint32_t idesc = item.nh < 0 ? idesc_full : idesc_half;
            // half-N item: this CTA's quarter (pair) / half (single CTA) of the weight rows it holds
            const uint32_t b_off = item.nh < 0 ? 0u : static_cast<uint32_t>(item.nh) * ((CG == 2 ? BN / 4 : BN / 2) * 128);
            uint32_t started = 0;  // 0 until the first MMA of this tile has been issued (per sub-tile: same flag)
            for (int g = item.g0; g < item.g1; ++g) {
                const int sa = ia % NA;
                mbar_wait(&a_full[sa], (ia / NA) & 1);
                ++ia;
                const uint32_t sA = smem_u32(smemA + sa * Cfg::A_PITCH);
                const bool is_halo = (g < ng1) && halo;
                const int nsteps = is_halo ? KS : 1;
                if constexpr (Cfg::MERGED) {
                    tc_fence_after();
                    if (elect_one()) {
                        for (int j = 0; j < nsteps; ++j) {
                            const uint64_t bdesc = make_smem_desc_sw128(sA + Cfg::A_STAGE + j * Cfg::B_STAGE + b_off, 16, 1024);
#pragma unroll
                            for (int m = 0; m < MT; ++m) {
                                const uint32_t a_off = is_halo ? (m * TILE_H + j) * ROW_BYTES : m * A_BYTES;
                                const uint64_t adesc = make_smem_desc_sw128(sA + a_off, 16, 1024);
#pragma unroll
                                for (int kk = 0; kk < KCHUNK / 8; ++kk)
                                    mma(d_base + m * BN, adesc + 2 * kk, bdesc + 2 * kk, idesc, (started | j | kk) ? 1u : 0u);
                            }
                        }
                        commit(&a_empty[sa]);  // frees the stage (in both CTAs) once these MMAs have read it
                        if (g == item.g1 - 1) commit(&tmem_full_bar[acc]);  // accumulator complete
                    }
                    __syncwarp();
                    started = 1;
                } else {
                for (int j = 0; j < nsteps; ++j) {
                    const int sb = ib % NB;
                    mbar_wait(&b_full[sb], (ib / NB) & 1);
                    ++ib;
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smemB + sb * Cfg::B_STAGE) + b_off, 16, 1024);
#pragma unroll
                        for (int m = 0; m < MT; ++m) {
                            // halo box: sub-tile m, vertical tap j starts (m*8 + j) image rows into the box
                            const uint32_t a_off = is_halo ? (m * TILE_H + j) * ROW_BYTES : m * A_BYTES;
                            const uint64_t adesc = make_smem_desc_sw128(sA + a_off, 16, 1024);
#pragma unroll
                            for (int kk = 0; kk < KCHUNK / 8; ++kk) {
                                // +32 bytes along K inside the 128-byte swizzle span = +2 in the (addr >> 4) field
                                mma(d_base + m * BN, adesc + 2 * kk, bdesc + 2 * kk, idesc, (started | kk) ? 1u : 0u);
                            }
                        }
                        commit(&b_empty[sb]);  // frees the weight slot (in both CTAs) once these MMAs have read it
                        if (j == nsteps - 1) commit(&a_empty[sa]);  // ... and the activation box after its last tap
                        if (j == nsteps - 1 && g == item.g1 - 1) commit(&tmem_full_bar[acc]);  // accumulator complete
                    }
                    __syncwarp();
                    started = 1;
                }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;  // TMEM lane quadrant this warp may access
        const int row = q * 32 + lane;
        const int hl = row / TILE_W;
        const int wl = row % TILE_W;
        const ConvEpilogue& ep = p.ep;
        const float ccoef = ep.cont_f ? *ep.cont_coef : 0.f;
        const int words = p.Cout >> 5;  // 32-channel words of the sign bitmaps per pixel
        const int grp = (warp - 4) >> 2;                 // epilogue group 0 / 1
        const int gtid = threadIdx.x - 128 - grp * 128;  // thread of the group
        const bool issuer = (gtid == 0);
        uint8_t* sbox = stage_box + grp * STAGE_BOX_BYTES;  // this group's staging box
        const int bar_a = 1 + 2 * grp, bar_b = 2 + 2 * grp;  // this group's named barriers
        uint32_t lt = 0;
        for (int it = cta_tile0; it < p.total_items; it += cta_tile_step, ++lt) {
            const Item item = get_item(it);
            int b, h0, w0, n0;
            decode(item.tile, b, h0, w0, n0);
            const int acc = lt % NACC;
            mbar_wait(&tmem_full_bar[acc], (lt / NACC) & 1);
            tc_fence_after();
            const uint32_t t_base = tmem_base + acc * Cfg::ACC_COLS + (static_cast<uint32_t>(q * 32) << 16);
            // K-split parts are handled by group 0 alone (the hand-over flags count one warp per lane quadrant)
            const bool solo = item.nparts > 1;
            if (solo && grp == 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (CG == 2) mbar_arrive_leader(&tmem_empty_bar[acc]); else mbar_arrive(&tmem_empty_bar[acc]); }
                continue;
            }
            // split-K: workspace of this tile's partial accumulators, [part][rank][m][row][BN]; flags [slot][rank][quadrant]
            constexpr size_t kPartElems = static_cast<size_t>(MT) * 128 * BN;
            float* ws_tile = p.splitk_ws + static_cast<size_t>(item.slot) * (p.split - (p.split_nowait ? 0 : 1)) * CG * kPartElems;
            unsigned int* flag = p.splitk_flags + (static_cast<size_t>(item.slot) * CG + rank) * 4 + q;
            if (item.part < item.nparts - 1 || (p.split_nowait && item.nparts > 1)) {
                // not the last part of a split tile: hand the raw partial sums to the CTA (pair) that holds the last part
                float* dst = ws_tile + (static_cast<size_t>(item.part) * CG + rank) * kPartElems;
#pragma unroll 1
                for (int m = 0; m < MT; ++m)
#pragma unroll 1
                    for (int c = 0; c < BN; c += 32) {
                        float v[32];
                        tmem_ld_x32(t_base + m * BN + c, v);
                        float4* d4 = reinterpret_cast<float4*>(dst + (static_cast<size_t>(m) * 128 + row) * BN + c);
#pragma unroll
                        for (int k = 0; k < 8; ++k) __stcg(d4 + k, make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
                    }
                __threadfence();
                __syncwarp();
                if (lane == 0 && !p.split_nowait) atomicAdd(flag, 1u);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (CG == 2) mbar_arrive_leader(&tmem_empty_bar[acc]); else mbar_arrive(&tmem_empty_bar[acc]); }
                continue;
            }
            const int nprev = item.nparts - 1;  // partial sums to add before the epilogue (0 for an unsplit tile)
            if (nprev > 0) {
                if (lane == 0) {
                    while (ld_acquire_gpu(flag) < static_cast<unsigned int>(nprev)) __nanosleep(64);
                }
                __syncwarp();
            }
            auto add_partials = [&](float* v, int n, int m, int c) {
                for (int pp = 0; pp < nprev; ++pp) {
                    const float4* s4 = reinterpret_cast<const float4*>(
                        ws_tile + (static_cast<size_t>(pp) * CG + rank) * kPartElems + (static_cast<size_t>(m) * 128 + row) * BN + c);
                    for (int k = 0; k < n / 4; ++k) {
                        const float4 u = __ldcg(s4 + k);
                        v[4 * k] += u.x; v[4 * k + 1] += u.y; v[4 * k + 2] += u.z; v[4 * k + 3] += u.w;
                    }
                }
            };
            const int ncols = item.nh < 0 ? BN : BN / 2;
            if (!p.direct) {
                // TMEM -> registers -> fused epilogue -> swizzled staging box in shared memory -> TMA store.  A warp-wide
                // 16-byte store straight to global would touch 32 different lines (the lanes are 32 different pixels): the
                // LSU retires it at one line per cycle, which is what kept the epilogue of the wide tiles on the critical
                // path.  The TMA engine writes whole lines and clips ragged tiles by itself.
#pragma unroll 1
                for (int m = 0; m < MT; ++m) {
                    const int h = h0 + m * TILE_H + hl;
                    const int w = w0 + wl;
                    const bool valid = (h < p.H) && (w < p.W);
                    const size_t pix = (static_cast<size_t>(b) * p.H + h) * p.W + w;
#pragma unroll 1
                    for (int c = 0; c < ncols; c += 32) {
                        // the two epilogue groups take alternate boxes (the parity flips from tile to tile: single-box tiles)
                        if (!solo && (((m * (ncols >> 5) + (c >> 5) + static_cast<int>(lt)) & 1) != grp)) continue;
                        const int cn = n0 + chan(c, item.nh);  // first output channel of this 32-column chunk
                        float v[32];
                        tmem_ld_x32(t_base + m * BN + c, v);
                        if (nprev > 0) add_partials(v, 32, m, c);
                        if (ep.bias) {
#pragma unroll
                            for (int i = 0; i < 32; i += 4) {
                                const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + cn + i));
                                v[i] += bv.x; v[i + 1] += bv.y; v[i + 2] += bv.z; v[i + 3] += bv.w;
                            }
                        }
                        if (ep.relu) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
                        }
                        if (ep.mask_bits) {
                            const uint32_t mk = valid ? __ldg(ep.mask_bits + pix * words + (cn >> 5)) : 0u;
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = ((mk >> i) & 1u) ? v[i] : 0.f;
                        }
                        if (ep.round) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = round_tf32(v[i]);
                        }
                        if (ep.mask_out && valid) {
                            uint32_t bits = 0;
#pragma unroll
                            for (int i = 0; i < 32; ++i) bits |= (v[i] > 0.f ? 1u : 0u) << i;
                            ep.mask_out[pix * words + (cn >> 5)] = bits;
                        }
                        // the store that last read this group's staging box must have drained it (it was issued a whole box of
                        // arithmetic ago)
                        if (issuer) bulk_wait_group_read<0>();
                        named_bar_sync(bar_a, 128);
                        uint8_t* srow = sbox + row * 128;
#pragma unroll
                        for (int k = 0; k < 8; ++k)  // 128B swizzle: 16-byte chunk k of row r lives at chunk k ^ (r % 8)
                            *reinterpret_cast<float4*>(srow + ((k ^ (row & 7)) << 4)) =
                                make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                        fence_proxy_async_smem();
                        named_bar_sync(bar_b, 128);
                        if (issuer) {
                            tma_store_4d(&tmOut, sbox, cn, w0, h0 + m * TILE_H, b);
                            if (ep.out2) tma_store_4d(&tmOut2, sbox, cn, w0, h0 + m * TILE_H, b);
                            bulk_commit_group();
                        }
                        if (POOL) {
                            // 2x2 / stride-2 pooling of the finished box (models.py:119-122) straight from the staging copy:
                            // the box holds 16 x 8 pixels x 32 channels, i.e. 8 x 4 pooled pixels x eight 16-byte channel chunks
                            // = 256 items, two per epilogue thread.  An item reads its window's four chunks (de-swizzled),
                            // reduces them and writes one 16-byte piece of the pooled NHWC map; eight consecutive threads cover
                            // one pooled pixel's 128-byte line.  Same arithmetic as pool_fwd_kernel, bit for bit.  The box is
                            // not rewritten before every thread has passed the next-but-one box's first barrier.
                            const int PH = p.H >> 1, PW = p.W >> 1;
                            const int tid = gtid;
#pragma unroll
                            for (int it2 = 0; it2 < 2; ++it2) {
                                const int item = tid + 128 * it2;
                                const int k = item & 7;          // 16-byte channel chunk
                                const int pp = item >> 3;        // pooled pixel of the box: 8 wide x 4 high
                                const int pwl = pp & 7, phl = pp >> 3;
                                const int ph = ((h0 + m * TILE_H) >> 1) + phl, pw = (w0 >> 1) + pwl;
                                float4 a[4];
#pragma unroll
                                for (int qd = 0; qd < 4; ++qd) {
                                    const int rr = (2 * phl + (qd >> 1)) * TILE_W + 2 * pwl + (qd & 1);
                                    a[qd] = *reinterpret_cast<const float4*>(sbox + rr * 128 + ((k ^ (rr & 7)) << 4));
                                }
                                float4 o;
                                if (ep.pool_avg) {
                                    o.x = 0.25f * (a[0].x + a[1].x + a[2].x + a[3].x); o.y = 0.25f * (a[0].y + a[1].y + a[2].y + a[3].y);
                                    o.z = 0.25f * (a[0].z + a[1].z + a[2].z + a[3].z); o.w = 0.25f * (a[0].w + a[1].w + a[2].w + a[3].w);
                                    if (ep.round) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
                                } else {
                                    o.x = fmaxf(fmaxf(a[0].x, a[1].x), fmaxf(a[2].x, a[3].x)); o.y = fmaxf(fmaxf(a[0].y, a[1].y), fmaxf(a[2].y, a[3].y));
                                    o.z = fmaxf(fmaxf(a[0].z, a[1].z), fmaxf(a[2].z, a[3].z)); o.w = fmaxf(fmaxf(a[0].w, a[1].w), fmaxf(a[2].w, a[3].w));
                                }
                                if (ph < PH && pw < PW) {
                                    const size_t po = ((static_cast<size_t>(b) * PH + ph) * PW + pw) * p.Cout + cn + 4 * k;
                                    *reinterpret_cast<float4*>(ep.pool_out + po) = o;
                                    if (ep.pool_codes && !ep.pool_avg)  // arg-max codes for the backward pass (ConvEpilogue::pool_codes)
                                        ep.pool_codes[po >> 2] = static_cast<uint8_t>(
                                            first_argmax(a[0].x, a[1].x, a[2].x, a[3].x) | (first_argmax(a[0].y, a[1].y, a[2].y, a[3].y) << 2) |
                                            (first_argmax(a[0].z, a[1].z, a[2].z, a[3].z) << 4) | (first_argmax(a[0].w, a[1].w, a[2].w, a[3].w) << 6));
                                }
                            }
                        }
                    }
                }
            } else {
#pragma unroll 1
            for (int m = 0; m < MT; ++m) {
                const int h = h0 + m * TILE_H + hl;
                const int w = w0 + wl;
                const bool valid = (h < p.H) && (w < p.W);
                const size_t pix = (static_cast<size_t>(b) * p.H + h) * p.W + w;
#pragma unroll 1
                for (int c0 = 0; c0 < ncols; c0 += 16) {
                    if (!solo && (((m * (ncols >> 4) + (c0 >> 4) + static_cast<int>(lt)) & 1) != grp)) continue;
                    const int cn = n0 + chan(c0, item.nh);
                    const size_t off = pix * p.Cout + cn;  // element offset of this 16-channel chunk
                    constexpr int c = 0;                   // (the chunk-relative offsets below were written as off + c + i)
                    float v[16];
                    tmem_ld_x16(t_base + m * BN + c0, v);
                    if (nprev > 0) add_partials(v, 16, m, c0);
                    if (valid) {
                        if (ep.bias) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 bv = *reinterpret_cast<const float4*>(ep.bias + cn + i);
                                v[i] += bv.x; v[i + 1] += bv.y; v[i + 2] += bv.z; v[i + 3] += bv.w;
                            }
                        }
                        if (ep.cont_f) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 f = *reinterpret_cast<const float4*>(ep.cont_f + off + c + i);
                                const float4 tg = *reinterpret_cast<const float4*>(ep.cont_t + off + c + i);
                                v[i] += ccoef * (f.x - tg.x); v[i + 1] += ccoef * (f.y - tg.y);
                                v[i + 2] += ccoef * (f.z - tg.z); v[i + 3] += ccoef * (f.w - tg.w);
                            }
                        }
                        if (ep.addend) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 a = *reinterpret_cast<const float4*>(ep.addend + off + c + i);
                                v[i] += a.x; v[i + 1] += a.y; v[i + 2] += a.z; v[i + 3] += a.w;
                            }
                        }
                        if (ep.relu) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                        }
                        if (ep.mask_src) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4) {
                                const float4 mk = *reinterpret_cast<const float4*>(ep.mask_src + off + c + i);
                                v[i] = mk.x > 0.f ? v[i] : 0.f; v[i + 1] = mk.y > 0.f ? v[i + 1] : 0.f;
                                v[i + 2] = mk.z > 0.f ? v[i + 2] : 0.f; v[i + 3] = mk.w > 0.f ? v[i + 3] : 0.f;
                            }
                        }
                        if (ep.mask_bits) {
                            const uint32_t mk = ep.mask_bits[pix * words + (cn >> 5)] >> (cn & 31);
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = ((mk >> i) & 1u) ? v[i] : 0.f;
                        }
                        if (ep.round) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = round_tf32(v[i]);
                        }
                        if (ep.mask_out) {
                            uint32_t bits = 0;
#pragma unroll
                            for (int i = 0; i < 16; ++i) bits |= (v[i] > 0.f ? 1u : 0u) << i;
                            reinterpret_cast<uint16_t*>(ep.mask_out)[pix * (2 * words) + (cn >> 4)] = (uint16_t)bits;
                        }
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4*>(ep.out + off + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if (ep.out2) {
#pragma unroll
                            for (int i = 0; i < 16; i += 4)
                                *reinterpret_cast<float4*>(ep.out2 + off + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        }
                    }
                }
            }
            }
            // all TMEM reads of this warp are complete (the loads wait): hand the accumulator back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (nprev > 0) *flag = 0u;  // re-arm for the next launch (this warp was the flag's only reader)
                if (CG == 2) mbar_arrive_leader(&tmem_empty_bar[acc]); else mbar_arrive(&tmem_empty_bar[acc]);
            }
        }
        if (issuer && !p.direct) bulk_wait_group<0>();  // outstanding output stores must land before the CTA retires
    }

    tc_fence_before();
    if (CG == 2) {
        cluster_sync_all();  // the peer may still be reading this CTA's shared memory / signalling its barriers
        if (warp == 2) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    } else {
        __syncthreads();
        if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

}  // namespace

// NHWC activation tensor [B][H][W][C] -> 4-D map, box {32, box_w, box_h, 1}, 128B swizzle.
int make_tmap_nhwc(CUtensorMap* m, const float* ptr, int B, int H, int W, int C, int box_w, int box_h) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled entry point not available (driver too old / no GPU)");
        return MAUA_ERR_CUDA;
    }
    MAUA_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "activation pointer must be 16-byte aligned");
    MAUA_REQUIRE(C % 4 == 0, "NHWC channel count %d must be a multiple of 4", C);
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {KCHUNK, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled(NHWC %dx%dx%dx%d) failed: CUresult %d", B, H, W, C, (int)r);
        return MAUA_ERR_CUDA;
    }
    return MAUA_OK;
}

// Row-major matrix [rows][cols] (cols contiguous) -> 2-D map, box {32, box_rows}, 128B swizzle.
int make_tmap_2d(CUtensorMap* m, const float* ptr, long rows, long cols, int box_rows, int atom32) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_last_error("cuTensorMapEncodeTiled entry point not available (driver too old / no GPU)");
        return MAUA_ERR_CUDA;
    }
    MAUA_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "matrix pointer must be 16-byte aligned");
    MAUA_REQUIRE(cols % 4 == 0, "matrix row length %ld must be a multiple of 4", cols);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {KCHUNK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled(2D %ldx%ld) failed: CUresult %d", rows, cols, (int)r);
        return MAUA_ERR_CUDA;
    }
    return MAUA_OK;
}

namespace {

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& n = cached[dev & 63];
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

constexpr int kMaxSplit = 8;
constexpr int kMinGroupsPerPart = 4;

// Split-K plan of a launch: the tiles of the last partial wave (all tiles when there are fewer tiles than CTA units) are
// split into `split` K-ranges each, as long as every range keeps >= kMinGroupsPerPart k-groups.
struct SplitPlan { int full_tiles, split, items, halves; };
// mode 0: every tile whole; 1: K-split tail (needs the workspace); 2: half-N tail -- the R tiles of the last partial wave become
// 2 R independent half-channel items when they fit the CTA units (2 R <= units) and the halves stay MMA / epilogue friendly.
SplitPlan plan_split(long tiles, int units, int ngroups, int mode, int bn = 0, int cg = 1) {
    SplitPlan sp{(int)tiles, 1, (int)tiles, 0};
    const int rem = (int)(tiles % units);
    if (mode == 0 || rem == 0) return sp;
    if (mode == 3 && tiles >= units) mode = 2;  // K-split + reduce kernel only when the whole launch is one partial wave
    if (mode == 2) {
        if (2 * rem > units || bn < (cg == 2 ? 128 : 64)) return sp;
        sp.full_tiles = (int)tiles - rem;
        sp.halves = 1;
        sp.items = sp.full_tiles + 2 * rem;
        return sp;
    }
    int s = units / rem;
    if (s > kMaxSplit) s = kMaxSplit;
    if (s > ngroups / kMinGroupsPerPart) s = ngroups / kMinGroupsPerPart;
    if (s < 2) return sp;
    sp.full_tiles = (int)tiles - rem;
    sp.split = s;
    sp.items = sp.full_tiles + rem * s;
    return sp;
}
int conv_groups(const ConvArgs& a) {
    return (a.ntaps == 9 ? 3 * (a.Cin / KCHUNK) : a.ntaps == 25 ? 5 * (a.Cin / KCHUNK) : (a.ntaps == 1 ? a.Cin / KCHUNK : 0)) +
           a.K2 / KCHUNK;
}
// tail handling requested by the caller: ConvArgs::tail_mode (K-split additionally needs the workspace)
int effective_tail_mode(const ConvArgs& a) {
    if (a.tail_mode == 1) return (a.splitk_ws && a.splitk_flags) ? 1 : 0;
    if (a.tail_mode == 3) return (a.splitk_ws && !a.ep.pool_out) ? 3 : 2;
    return a.tail_mode == 2 ? 2 : 0;
}

// Tile selection: relative MAC rates of the tile shapes measured on B200 (tools/sweep_conv.sh, profiles/ round 1) times
// the occupancy of the waves.  What the sweep shows: with both operands in shared memory an MMA is paced by the
// operand read (4 KB of activations + 32 B per weight row the CTA holds), so wide tiles and CTA pairs (cta_group::2,
// each CTA holds half of the weight rows) win, and a tile needs its accumulator double-buffered in TMEM (MT * BN <= 256)
// to keep the epilogue off the critical path.  With split-K (plan_split) the last partial wave costs 1 / split of a tile
// time instead of a whole one, which is what decides the shape at <= 512^2 where every layer is a partial wave.
// tile shapes the 5x5 kernel is instantiated for (conv_tc_launch)
bool ks5_shape(int bn, int mt, int cg) {
    return (bn == 256 && mt == 1 && cg == 2) || (bn == 128 && mt == 2 && cg == 2) || (bn == 128 && mt == 1 && cg == 1) ||
           (bn == 64 && mt == 2 && cg == 2) || (bn == 64 && mt == 1 && cg == 1);
}

SplitPlan choose_tile(const ConvArgs& a, int sms, int tail_mode, int& bn_out, int& mt_out, int& cg_out) {
    // (re-fitted after the merged pipeline stages: tools/sweep_conv_small.py with SWEEP_BIG=1, full-wave layers, round 2)
    auto shape_rate = [](int bn, int mt, int cg) -> double {
        if (cg == 2) {
            if (bn == 256) return mt == 1 ? 0.96 : 0.90;
            if (bn == 128) return mt == 1 ? 1.00 : 0.98;
            if (bn == 64) return mt == 2 ? 0.83 : 0.78;
            return mt == 2 ? 0.51 : 0.43;
        }
        if (bn == 256) return mt == 1 ? 0.87 : 0.88;
        if (bn == 128) return mt == 2 ? 0.76 : 0.63;
        if (bn == 64) return mt == 2 ? 0.64 : 0.56;
        return mt == 2 ? 0.385 : 0.32;
    };
    int best_bn = 32, best_mt = 1, best_cg = 1;
    double best = -1.0;
    SplitPlan best_sp{0, 1, 0};
    const int ngroups = conv_groups(a);
    for (int bn = 256; bn >= 32; bn >>= 1) {
        if (a.Cout % bn) continue;
        for (int mt = 2; mt >= 1; --mt)
            for (int cg = 2; cg >= 1; --cg) {
                if (a.force_cg && cg != a.force_cg) continue;
                if (a.ntaps == 25 && !ks5_shape(bn, mt, cg)) continue;
                const int units = sms / cg;
                const long tiles = (long)((a.W + TILE_W - 1) / TILE_W) *
                                   ((a.H + cg * mt * TILE_H - 1) / (cg * mt * TILE_H)) * a.B * (a.Cout / bn);
                // time in units of one tile: full waves + the split last wave (1 / split, plus the hand-over of partials)
                const SplitPlan sp = plan_split(tiles, units, ngroups, tail_mode, bn, cg);
                // a half-N item does half the MMA work at the (lower) rate of the half-width shape
                const double waves = sp.halves ? (double)(sp.full_tiles / units) + 0.5 * shape_rate(bn, mt, cg) / shape_rate(bn / 2, mt, cg)
                                     // (tail mode 3 hands the parts over through a second kernel: ~ a third of a short wave, measured)
                                     : sp.split > 1 ? (double)(sp.full_tiles / units) + 1.0 / sp.split + (tail_mode == 3 ? 0.35 : 0.03)
                                                    : (double)((tiles + units - 1) / units);
                const double eff = (double)tiles / (waves * units);
                // rows of the (pair) tile that exist: ragged bottoms waste MMA work
                const double rows = (double)a.H / (double)(((a.H + cg * mt * TILE_H - 1) / (cg * mt * TILE_H)) * cg * mt * TILE_H);
                // fused pooling and short-K layers (Cin = 64: the epilogue weighs more): the 128-wide pair tile does better with two
                // stacked sub-tiles (conv2_2 + pool at 512^2 x 128: 112 vs 121 us; conv2_1 at 512^2: 60.4 vs 64.5 us)
                const double pool_f = ((a.ep.pool_out || (a.ntaps == 9 && a.Cin <= 64)) && bn == 128 && cg == 2 && mt == 1) ? 0.92 : 1.0;
                const double score = eff * rows * shape_rate(bn, mt, cg) * pool_f;
                if (score > best) { best = score; best_bn = bn; best_mt = mt; best_cg = cg; best_sp = sp; }
            }
    }
    int bn = best_bn, mt = best_mt, cg = best_cg;
    if (const char* f = getenv("MAUA_CONV_FORCE")) {  // developer override for tile-shape experiments: "bn,mt,cg"
        int fb = 0, fm = 0, fc = 0;
        if (sscanf(f, "%d,%d,%d", &fb, &fm, &fc) == 3 && a.Cout % fb == 0 && (fm == 1 || fm == 2) && (fc == 1 || fc == 2) &&
            (fb == 32 || fb == 64 || fb == 128 || fb == 256) && (a.ntaps != 25 || ks5_shape(fb, fm, fc))) {
            bn = fb; mt = fm; cg = fc;
            const long tiles = (long)((a.W + TILE_W - 1) / TILE_W) * ((a.H + cg * mt * TILE_H - 1) / (cg * mt * TILE_H)) * a.B * (a.Cout / bn);
            best_sp = plan_split(tiles, sms / cg, ngroups, tail_mode, bn, cg);
        }
    }
    bn_out = bn; mt_out = mt; cg_out = cg;
    return best_sp;
}

// Tail mode 3: sum of the K-split parts of every tile + the fused epilogue of ConvEpilogue (same order of operations as the
// in-kernel epilogue; fixed summation order: deterministic).  Workspace layout: [tile][part][rank][m][row][BN] raw accumulators.
struct SplitReduce {
    const float* ws;
    int split, ntiles, BN, MT, CG;
    int B, H, W, Cout, tiles_w, tiles_h, n_tiles;
    ConvEpilogue ep;
};
__global__ void __launch_bounds__(256)
conv_splitk_reduce_kernel(const SplitReduce r) {
    pdl_wait();
    pdl_trigger();
    const int c4n = r.BN / 4;
    const long total = (long)r.ntiles * r.CG * r.MT * 128 * c4n;
    const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;  // (total is a multiple of 32: whole warps are live or dead together)
    if (idx >= total) return;
    const int c = (int)(idx % c4n) * 4;
    long t = idx / c4n;
    const int row = (int)(t % 128);
    t /= 128;
    const int m = (int)(t % r.MT);
    t /= r.MT;
    const int rank = (int)(t % r.CG);
    const int tile = (int)(t / r.CG);
    const size_t part_elems = (size_t)r.MT * 128 * r.BN;
    const float* src = r.ws + ((size_t)tile * r.split * r.CG + rank) * part_elems + ((size_t)m * 128 + row) * r.BN + c;
    float4 acc = __ldcg(reinterpret_cast<const float4*>(src));
    for (int pp = 1; pp < r.split; ++pp) {
        const float4 u = __ldcg(reinterpret_cast<const float4*>(src + (size_t)pp * r.CG * part_elems));
        acc.x += u.x; acc.y += u.y; acc.z += u.z; acc.w += u.w;
    }
    const int nt = tile % r.n_tiles;
    int pt = tile / r.n_tiles;
    const int tw = pt % r.tiles_w;
    pt /= r.tiles_w;
    const int th = pt % r.tiles_h;
    const int b = pt / r.tiles_h;
    const int h = th * (TILE_H * r.MT * r.CG) + rank * (TILE_H * r.MT) + m * TILE_H + row / TILE_W;
    const int w = tw * TILE_W + row % TILE_W;
    const int n = nt * r.BN + c;
    const bool valid = h < r.H && w < r.W;
    const ConvEpilogue& ep = r.ep;
    float v[4] = {acc.x, acc.y, acc.z, acc.w};
    const size_t pix = ((size_t)b * r.H + h) * r.W + w;
    const size_t off = pix * r.Cout + n;
    if (valid) {
        if (ep.bias) { const float4 bv = __ldg(reinterpret_cast<const float4*>(ep.bias + n)); v[0] += bv.x; v[1] += bv.y; v[2] += bv.z; v[3] += bv.w; }
        if (ep.cont_f) {
            const float cc = *ep.cont_coef;
            const float4 f = *reinterpret_cast<const float4*>(ep.cont_f + off), tg = *reinterpret_cast<const float4*>(ep.cont_t + off);
            v[0] += cc * (f.x - tg.x); v[1] += cc * (f.y - tg.y); v[2] += cc * (f.z - tg.z); v[3] += cc * (f.w - tg.w);
        }
        if (ep.addend) { const float4 a4 = *reinterpret_cast<const float4*>(ep.addend + off); v[0] += a4.x; v[1] += a4.y; v[2] += a4.z; v[3] += a4.w; }
        if (ep.relu) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
        if (ep.mask_src) {
            const float4 mk = *reinterpret_cast<const float4*>(ep.mask_src + off);
            v[0] = mk.x > 0.f ? v[0] : 0.f; v[1] = mk.y > 0.f ? v[1] : 0.f; v[2] = mk.z > 0.f ? v[2] : 0.f; v[3] = mk.w > 0.f ? v[3] : 0.f;
        }
        if (ep.mask_bits) {
            const uint32_t mk = ep.mask_bits[pix * (r.Cout >> 5) + (n >> 5)] >> (n & 31);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = ((mk >> i) & 1u) ? v[i] : 0.f;
        }
        if (ep.round) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = round_tf32(v[i]);
        }
        const float4 o = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(ep.out + off) = o;
        if (ep.out2) *reinterpret_cast<float4*>(ep.out2 + off) = o;
    }
    if (ep.mask_out) {
        // sign bitmap: the 8 threads that hold the 32 channels of one word are 8 consecutive lanes
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) bits |= (v[i] > 0.f ? 1u : 0u) << ((n & 31) + i);
        bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
        bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
        bits |= __shfl_xor_sync(0xffffffffu, bits, 4);
        if (valid && (n & 31) == 0) ep.mask_out[pix * (r.Cout >> 5) + (n >> 5)] = bits;
    }
}

template <int BN, int MT, int CG, bool POOL, int KS = 3>
int launch_cfg(const ConvArgs& a, cudaStream_t st) {
    using Cfg = ConvCfg<BN, MT, CG, KS>;
    static unsigned long long attr_done = 0;  // per (BN, MT, CG, POOL, KS) instantiation; benign race (idempotent call)
    MAUA_CUDA_CHECK((ensure_dynamic_smem(conv_tc_kernel<BN, MT, CG, POOL, KS>, Cfg::SMEM_BYTES, &attr_done)));
    CUtensorMap tmA, tmB, tmA2, tmB2;
    int rc;
    const bool has_main = a.ntaps > 0;
    const bool has_aux = a.K2 > 0;
    if (has_main) {
        // 3x3: halo box (MT*8 + 2 image rows) shared by the three vertical taps; pointwise: plain box
        if ((rc = make_tmap_nhwc(&tmA, a.in, a.B, a.H, a.W, a.Cin, TILE_W, a.ntaps == KS * KS ? Cfg::A_ROWS : MT * TILE_H))) return rc;
        if ((rc = make_tmap_2d(&tmB, a.wg, a.Cout, (long)a.ntaps * a.Cin, Cfg::BNL, 0))) return rc;
    }
    if (has_aux) {
        if ((rc = make_tmap_nhwc(&tmA2, a.in2, a.B, a.H, a.W, a.K2, TILE_W, MT * TILE_H))) return rc;
        if ((rc = make_tmap_2d(&tmB2, a.w2, a.Cout, a.K2, Cfg::BNL, 0))) return rc;
    }
    if (!has_main) { tmA = tmA2; tmB = tmB2; }
    if (!has_aux) { tmA2 = tmA; tmB2 = tmB; }
    // epilogue flavour: the TMA-store path covers bias / ReLU / bitmap mask / rounding; the content, addend and
    // fp32-mask terms need per-element global loads and keep the register -> global path
    const bool direct = a.ep.cont_f || a.ep.addend || a.ep.mask_src;
    MAUA_REQUIRE(!(a.ep.pool_out && direct), "fused pooling needs the TMA-store epilogue (no content / addend / fp32-mask terms)");
    CUtensorMap tmOut, tmOut2;
    if (!direct) {
        if ((rc = make_tmap_nhwc(&tmOut, a.ep.out, a.B, a.H, a.W, a.Cout, TILE_W, TILE_H))) return rc;
        if (a.ep.out2) { if ((rc = make_tmap_nhwc(&tmOut2, a.ep.out2, a.B, a.H, a.W, a.Cout, TILE_W, TILE_H))) return rc; }
        else tmOut2 = tmOut;
    } else {
        tmOut = tmA; tmOut2 = tmA;
    }

    ConvKParams p;
    p.B = a.B; p.H = a.H; p.W = a.W;
    p.Cin = has_main ? a.Cin : KCHUNK;
    p.Cout = a.Cout;
    p.ntaps = has_main ? a.ntaps : 0;
    p.K2 = a.K2;
    p.tiles_w = (a.W + TILE_W - 1) / TILE_W;
    p.tiles_h = (a.H + TILE_H * MT * CG - 1) / (TILE_H * MT * CG);  // (pair) tiles
    p.n_tiles = a.Cout / BN;
    p.total_tiles = p.tiles_w * p.tiles_h * a.B * p.n_tiles;
    p.ep = a.ep;
    p.direct = direct ? 1 : 0;
    p.prefetch = static_cast<const uint8_t*>(a.prefetch);
    p.prefetch_bytes = (a.prefetch && (reinterpret_cast<uintptr_t>(a.prefetch) & 15) == 0) ? (a.prefetch_bytes & ~size_t(15)) : 0;
    const int units = num_sms() / CG;  // persistent: one CTA (pair) per SM (pair)
    const SplitPlan sp = plan_split(p.total_tiles, units, conv_groups(a), effective_tail_mode(a), BN, CG);
    p.full_tiles = sp.full_tiles; p.split = sp.split; p.total_items = sp.items; p.tail_halves = sp.halves;
    p.splitk_ws = a.splitk_ws; p.splitk_flags = a.splitk_flags;
    p.split_nowait = (effective_tail_mode(a) == 3 && sp.split > 1 && !sp.halves && sp.full_tiles == 0) ? 1 : 0;
    const int grid = CG * (p.total_items < units ? p.total_items : units);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kConvThreads);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (CG > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = CG; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled(PDL_CONV)) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    MAUA_CUDA_CHECK((cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, MT, CG, POOL, KS>, tmA, tmB, tmA2, tmB2, tmOut, tmOut2, p)));
    if (p.split_nowait) {
        // sums of the K-split parts + the fused epilogue, one float4 per thread
        SplitReduce r;
        r.ws = p.splitk_ws; r.split = p.split; r.ntiles = p.total_tiles; r.BN = BN; r.MT = MT; r.CG = CG;
        r.B = p.B; r.H = p.H; r.W = p.W; r.Cout = p.Cout; r.tiles_w = p.tiles_w; r.tiles_h = p.tiles_h; r.n_tiles = p.n_tiles;
        r.ep = a.ep;
        const long total = (long)p.total_tiles * CG * MT * 128 * (BN / 4);
        MAUA_CUDA_CHECK(launch_pdl<PDL_CONV>(conv_splitk_reduce_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, r));
    }
    return MAUA_OK;
}

template <int BN, int MT>
int launch_cg(const ConvArgs& a, int cg, cudaStream_t st) {
    if (a.ep.pool_out) return cg == 2 ? launch_cfg<BN, MT, 2, true>(a, st) : launch_cfg<BN, MT, 1, true>(a, st);
    return cg == 2 ? launch_cfg<BN, MT, 2, false>(a, st) : launch_cfg<BN, MT, 1, false>(a, st);
}

}  // namespace

size_t conv_splitk_ws_bytes() { return (size_t)148 * 128 * 512 * sizeof(float); }  // < 148 split parts x (MT * 128 x BN <= 512 cols)
size_t conv_splitk_flag_words() { return 148 * 4 * 2; }

// host-logic view of the launch plan (no GPU needed): tile shape + split-K plan for a layer on a device with `sms` SMs
void conv_tile_plan(const ConvArgs& a, int sms, int tail_mode, int* bn, int* mt, int* cg, int* full_tiles, int* split_tiles, int* split) {
    int b, m, c;
    const SplitPlan sp = choose_tile(a, sms, tail_mode, b, m, c);
    *bn = b; *mt = m; *cg = c;
    *full_tiles = sp.full_tiles;
    *split = sp.halves ? 2 : sp.split;
    *split_tiles = sp.halves ? (sp.items - sp.full_tiles) / 2 : (sp.split > 1 ? (sp.items - sp.full_tiles) / sp.split : 0);
}

int conv_tc_launch(const ConvArgs& a, cudaStream_t st) {
    MAUA_REQUIRE(a.ntaps == 25 || a.ntaps == 9 || a.ntaps == 1 || a.ntaps == 0, "ntaps must be 25, 9, 1 or 0 (got %d)", a.ntaps);
    MAUA_REQUIRE(a.ntaps > 0 || a.K2 > 0, "conv has neither a main nor an aux term");
    MAUA_REQUIRE(a.ntaps == 0 || (a.Cin % KCHUNK == 0 && a.Cin >= KCHUNK),
                 "tcgen05 conv needs Cin %% 32 == 0 (got %d); the 3-channel image layer uses conv_first_*", a.Cin);
    MAUA_REQUIRE(a.K2 % KCHUNK == 0, "aux K2 must be a multiple of 32 (got %d)", a.K2);
    MAUA_REQUIRE(a.Cout % 32 == 0, "Cout must be a multiple of 32 (got %d)", a.Cout);
    MAUA_REQUIRE((reinterpret_cast<uintptr_t>(a.ep.out) & 15) == 0, "output pointer must be 16-byte aligned");
    MAUA_REQUIRE(a.B >= 1 && a.H >= 1 && a.W >= 1, "bad extent B=%d H=%d W=%d", a.B, a.H, a.W);
    MAUA_REQUIRE(a.ep.out != nullptr, "null output pointer");

    int bn, mt, cg;
    const SplitPlan spc = choose_tile(a, num_sms(), effective_tail_mode(a), bn, mt, cg);
    if (getenv("MAUA_CONV_DEBUG")) {  // one line per launch: the tile shape and split-K plan that was chosen
        fprintf(stderr, "conv_tc %dx%d Cin %d Cout %d taps %d K2 %d: BN %d MT %d CG %d, %d whole tiles + %d tail tiles (%s x%d) on %d units\n",
                a.H, a.W, a.Cin, a.Cout, a.ntaps, a.K2, bn, mt, cg, spc.full_tiles,
                spc.halves ? (spc.items - spc.full_tiles) / 2 : (spc.split > 1 ? (spc.items - spc.full_tiles) / spc.split : 0),
                spc.halves ? "half-N" : "K-split", spc.halves ? 2 : spc.split, num_sms() / cg);
    }
    if (a.ntaps == 25) {
        // 5x5: the shapes ks5_shape() lets the chooser pick (each is one more instantiation of the kernel)
        MAUA_REQUIRE(!a.ep.pool_out, "the 5x5 convolution has no fused pooling");
        if (bn == 256 && mt == 1 && cg == 2) return launch_cfg<256, 1, 2, false, 5>(a, st);
        if (bn == 128 && mt == 2 && cg == 2) return launch_cfg<128, 2, 2, false, 5>(a, st);
        if (bn == 128 && mt == 1 && cg == 1) return launch_cfg<128, 1, 1, false, 5>(a, st);
        if (bn == 64 && mt == 2 && cg == 2) return launch_cfg<64, 2, 2, false, 5>(a, st);
        if (bn == 64 && mt == 1 && cg == 1) return launch_cfg<64, 1, 1, false, 5>(a, st);
        MAUA_REQUIRE(false, "5x5 convolution: no kernel for tile shape BN %d MT %d CG %d (Cout %d must be a multiple of 64)", bn, mt, cg, a.Cout);
    }
    if (bn == 256) return mt == 2 ? launch_cg<256, 2>(a, cg, st) : launch_cg<256, 1>(a, cg, st);
    if (bn == 128) return mt == 2 ? launch_cg<128, 2>(a, cg, st) : launch_cg<128, 1>(a, cg, st);
    if (bn == 64) return mt == 2 ? launch_cg<64, 2>(a, cg, st) : launch_cg<64, 1>(a, cg, st);
    return mt == 2 ? launch_cg<32, 2>(a, cg, st) : launch_cg<32, 1>(a, cg, st);
}

// ---------------------------------------------------------------------------------------------
// Naive SIMT cross-check of exactly the same contract (debug / tests only, any channel counts).
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void conv_ref_kernel(ConvArgs a) {
    const long total = (long)a.B * a.H * a.W * a.Cout;
    for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
        const int n = idx % a.Cout;
        long pix = idx / a.Cout;
        const int w = pix % a.W;
        const int h = (pix / a.W) % a.H;
        const int b = pix / ((long)a.W * a.H);
        float acc = 0.f;
        for (int tap = 0; tap < a.ntaps; ++tap) {
            const int ks = a.ntaps == 25 ? 5 : (a.ntaps == 9 ? 3 : 1);
            const int dy = tap / ks - ks / 2, dx = tap % ks - ks / 2;
            const int hh = h + dy, ww = w + dx;
            if (hh < 0 || hh >= a.H || ww < 0 || ww >= a.W) continue;
            const float* ip = a.in + (((long)b * a.H + hh) * a.W + ww) * a.Cin;
            const float* wp = a.wg + (long)n * a.ntaps * a.Cin + (long)tap * a.Cin;
            for (int c = 0; c < a.Cin; ++c) acc = fmaf(ip[c], wp[c], acc);
        }
        for (int c = 0; c < a.K2; ++c) acc = fmaf(a.in2[pix * a.K2 + c], a.w2[(long)n * a.K2 + c], acc);
        const ConvEpilogue& ep = a.ep;
        float v = acc;
        if (ep.bias) v += ep.bias[n];
        if (ep.cont_f) v += *ep.cont_coef * (ep.cont_f[idx] - ep.cont_t[idx]);
        if (ep.addend) v += ep.addend[idx];
        if (ep.relu) v = fmaxf(v, 0.f);
        if (ep.mask_src) v = ep.mask_src[idx] > 0.f ? v : 0.f;
        if (ep.mask_bits) v = ((ep.mask_bits[pix * (a.Cout >> 5) + (n >> 5)] >> (n & 31)) & 1u) ? v : 0.f;
        if (ep.round) v = round_tf32(v);
        ep.out[idx] = v;
        if (ep.out2) ep.out2[idx] = v;
    }
}
}  // namespace

int conv_ref_launch(const ConvArgs& a, cudaStream_t st) {
    MAUA_REQUIRE(a.ep.out != nullptr, "null output pointer");
    const long total = (long)a.B * a.H * a.W * a.Cout;
    const int blocks = (int)((total + 255) / 256 > 148 * 32 ? 148 * 32 : (total + 255) / 256);
    conv_ref_kernel<<<blocks, 256, 0, st>>>(a);
    MAUA_CUDA_CHECK(cudaGetLastError());
    if (a.ep.mask_out) {
        const int rc = relu_mask_bits_launch(a.ep.out, a.ep.mask_out, (long)a.B * a.H * a.W, a.Cout, st);
        if (rc) return rc;
    }
    if (a.ep.pool_out) return pool_fwd_launch(a.ep.out, a.ep.pool_out, a.B, a.H, a.W, a.Cout, a.ep.pool_avg, a.ep.round, st, a.ep.pool_codes);
    return MAUA_OK;
}

}  // namespace maua
