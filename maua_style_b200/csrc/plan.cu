// maua_plan: the whole feval of reference optim.py:201-238 -- net(pastiche) through the truncated VGG stack
// (models.py:351-453), the loss modules spliced into it (loss.py), and the backward pass to the image -- as a
// fixed sequence of kernel launches over library-owned workspaces.
//
// Data layout in HBM: the image is NCHW [1,3,H,W] at the boundary (the reference's pastiche); every feature map
// is NHWC fp32 with values rounded to TF32 in the producing epilogue.  All post-ReLU activations of the current
// forward live in one arena (they double as ReLU masks / max-pool argmax sources in the backward pass, so no
// separate mask or index tensors exist); gradients ping-pong between three scratch buffers.
//
// Backward structure (weights frozen, models.py:443-445 => input gradients only).  Let G_e be the gradient
// w.r.t. the post-ReLU output of conv entry e and Gm_e = G_e * (out_e > 0).  Each launch produces Gm_e directly:
//   consumer is a conv c   : Gm_e = mask( dgrad_c(Gm_c) + style taps at e [aux GEMM k-steps] + content tap [epilogue] )
//   consumer is a pool     : Gp = dgrad_c(Gm_c) ; Gm_e = mask( unpool(Gp) + tap gradients at e )
//   e is the last entry    : Gm_e = mask( tap gradients at e )            (aux-GEMM-only launch)
// and finally pastiche.grad = conv1_1 dgrad(Gm_0) + TV + temporal terms (conv_edge.cu).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "conv_tc.cuh"
#include "gram.cuh"
#include "maua_b200.h"
#include "pointwise.cuh"

namespace maua {
int check_device_arch(int device);
ReduceScratch scratch_from_workspace(void* ws);

namespace {

struct Entry {
    bool pool = false;
    int cin = 0, cout = 0;     // conv only
    int ks = 3, stride = 1, pad = 1;  // conv geometry: 3x3 / 1 / 1 (VGG), 1x1, 5x5 / 1 / 2, 11x11 / 4 / 0 (NIN, models.py:82-111)
    bool pool3 = false;        // pool only: 3x3 / stride 2 / ceil_mode (NIN, models.py:77-80) instead of 2x2 / stride 2
    float *w1g = nullptr, *w1t = nullptr;  // 11x11 / 4 image layer only: GEMM weights [cout][KP] and [KP][cout] of the im2col form
    float* w_flip = nullptr;   // 5x5 only: [cin][cout][5][5] rotated copy = the weights of the input-gradient convolution
    int conv_index = -1;       // = relu index (global, counted from conv1_1)
    bool image_layer = false;  // conv1_1: consumes the NCHW image (conv_edge.cu)
    float* w_raw = nullptr;    // OIHW copy (first conv only needs it, kept for all: 52 MB total)
    float* wg = nullptr;       // forward GEMM weights  [cout][9*cin]
    float* wd = nullptr;       // dgrad GEMM weights    [cin][9*cout]
    float* bias = nullptr;
    float* wt1 = nullptr;      // first conv only: [32][cout] tap-column weights of the dgrad contraction
    float *wg32 = nullptr, *wd32 = nullptr, *wt1_32 = nullptr;  // un-rounded copies, MAUA_IMPL_FP32 only (made on demand)
    // per forward
    int H = 0, W = 0, C = 0;   // output extent
    int Cn = 0;                // conv only: channel count the loss normalisations use (< C when the layer is zero-padded)
    float* out = nullptr;      // arena pointer
    uint32_t* bits = nullptr;  // conv only: sign bitmap of `out` (1 bit / element), the ReLU mask of the backward pass
    uint8_t* codes = nullptr;  // 2x2 pool only: arg-max codes of `out` (1 byte / 4 channels), written by the forward pass
    bool codes_ok = false;     // ... by the last forward pass (max pooling, product path)
};

struct Tap {
    int kind = 0;
    int relu_index = 0;
    int entry = -1;
    int C = 0;
    int Cn = 0;  // normalisation channel count (maua_net_desc::norm_channels)
    float *gram = nullptr, *diff = nullptr, *mean = nullptr, *aux_d = nullptr, *aux_bias = nullptr;
    void* gram_ws = nullptr;
    // state of the last forward
    int mode = 0;
    bool active = false;  // contributes to the backward pass
    int use_cov = 0;
    float* target = nullptr;
    // MAUA_MODE_EXTERNAL (img_vid windows, loss.py:141-181 with B > 1): the caller computes the style loss from the
    // tap features of all frames and hands the backward GEMM term of THIS frame back through maua_plan_set_tap_fold
    const float* ext_in2 = nullptr;   // NHWC [H][W][ext_K2]
    int ext_K2 = 0;
    const float* ext_w2 = nullptr;    // [C][ext_K2]
    const float* ext_bias = nullptr;  // [C] or null
};

// Backward prologue in ONE launch: coef2[i] = coefs[i] * factor[i] for every module slot (block (0,0)), and for every live
// style tap (blockIdx.y) the backward coefficient matrix  aux_d = coef * 4 / (C^3 P) * (G - A), TF32-rounded because it is
// the B operand of the folded StyleLoss-backward MMA (loss.py:141-157).
struct BwdPrep {
    int n_slots, n_style;
    float factors[MAUA_MAX_TAPS + 2];
    const float* diff[MAUA_MAX_TAPS];
    float* aux_d[MAUA_MAX_TAPS];
    int C[MAUA_MAX_TAPS];
    float inv_c3p[MAUA_MAX_TAPS];  // 4 / (C^3 P)
    int slot[MAUA_MAX_TAPS];
    int do_round;
};
__global__ void bwd_prep_kernel(const float* __restrict__ coefs, float* __restrict__ coef2, const BwdPrep bp) {
    pdl_wait();
    pdl_trigger();
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < bp.n_slots)
        coef2[threadIdx.x] = coefs[threadIdx.x] * bp.factors[threadIdx.x];
    if ((int)blockIdx.y >= bp.n_style) return;
    const int j = blockIdx.y;
    const long total = (long)bp.C[j] * bp.C[j];
    const float k = coefs[bp.slot[j]] * bp.factors[bp.slot[j]] * bp.inv_c3p[j];
    const float* __restrict__ diff = bp.diff[j];
    float* __restrict__ aux = bp.aux_d[j];
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
        aux[i] = bp.do_round ? round_tf32(k * diff[i]) : k * diff[i];
}

struct CoefParams {
    int n;
    float strength[MAUA_MAX_TAPS + 2];
    float vsf[MAUA_MAX_TAPS + 2];
    int normalize[MAUA_MAX_TAPS + 2];
    int kind[MAUA_MAX_TAPS + 2];
};
__device__ __forceinline__ float sg(float x) { return x / (fabsf(x) + 1e-8f); }
// ScaleGradients (loss.py:10-20) applied to the scalar loss: grad / (|grad| + 1e-8) * strength^2
__global__ void loss_grad_coefs_kernel(const float* __restrict__ up, float* __restrict__ coefs, CoefParams p) {
    pdl_wait();
    pdl_trigger();
    const int i = threadIdx.x;
    if (i >= p.n) return;
    const float u = up[i], s = p.strength[i], v = p.vsf[i];
    float c;
    if (p.kind[i] == 2) {
        c = u * s;  // TVLoss: never normalised (loss.py:232)
    } else if (p.kind[i] == 1) {
        c = p.normalize[i] ? sg(u * s) * s * s : u * s;
    } else {
        if (p.normalize[i]) c = sg(u * s) * s * s + (v > 0.f ? sg(u * v * s) * s * s : 0.f);
        else c = u * s * (1.f + (v > 0.f ? v : 0.f));
    }
    coefs[i] = c;
}

}  // namespace
}  // namespace maua

using namespace maua;

struct maua_plan {
    int device = 0;
    int impl = MAUA_IMPL_TC;
    // layer-wise split (models.py:503-566): this plan covers global entries [begin, begin + entries.size())
    int begin = 0;
    int in_C = 3;                 // channels of the stage input (3 = the image)
    bool last_stage = true;
    const float* stage_input = nullptr;  // NHWC boundary activation of the last forward (begin > 0)
    std::vector<Entry> entries;
    std::vector<Tap> taps;
    int avg_pool = 0;
    // K-split of the last partial wave of conv tiles (conv_tc.cu).  Measured on B200 in round 2 (profiles/r02_splitk_ab.txt):
    // the hand-over of partial accumulators through L2 costs ~20 us per split wave, more than the idle SMs it recovers at
    // every size from 256^2 to 2048^2, so it is OFF by default (MAUA_SPLITK=1 / maua_plan_set_splitk turn it on).
    bool splitk = false;
    int conv_tail = 2;            // half-N items for the last partial wave (0: whole tiles; K-split: `splitk`)
    bool fuse_pool = true;        // pool inside the producing conv's epilogue (MAUA_FUSE_POOL=0 at plan creation: separate pass)
    // Loss modules on a side stream (forward pass): the Gram SYRK + finalize of a style tap and the MSE of a content tap only
    // read the tap's feature map, so they run on `side` next to the following convolutions (fork: event after the producing
    // conv; join: once, at the end of the forward pass).  -1 = automatic: on while the image has at most kSideAutoPixels
    // pixels -- there the conv launches leave SMs idle and every launch is latency; at 1024^2 the persistent conv CTAs hold
    // every SM and the side work only delays them.  0 / 1 force it (MAUA_SIDE_STREAM, maua_plan_set_side_stream).
    bool pool_codes = true;        // un-pool from arg-max codes + sign bitmaps (pool_bwd_codes_kernel; MAUA_POOL_CODES=0: re-read the activation)
    bool prefetch_weights = true;  // L2 prefetch of the next conv's weights by the running conv kernel (MAUA_PREFETCH_W=0: off)
    int side_mode = -1;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    size_t weight_bytes = 0;
    // workspaces
    float* arena = nullptr;
    size_t arena_elems = 0;
    int ws_generation = 0;        // bumped whenever a workspace is (re)allocated: a CUDA graph captured before is stale
    uint32_t* bits_arena = nullptr;
    size_t bits_words = 0;
    float* gbuf[3] = {nullptr, nullptr, nullptr};
    size_t gbuf_elems = 0;
    void* reduce_ws = nullptr;
    float* coef2 = nullptr;       // [n_taps + 2] scaled coefficients
    float* im2col_ws = nullptr;   // NIN image layer (11x11 / 4): im2col matrix / its gradient [OH*OW][kIm2colK]
    size_t im2col_elems = 0;
    float* splitk_ws = nullptr;   // partial accumulators of K-split conv tiles (conv_tc.cu)
    unsigned int* splitk_flags = nullptr;
    size_t tap_bytes = 0;
    // last forward
    int H = 0, W = 0;
    int last_entry = -1;          // last executed entry
    bool can_backward = false;
    const float* image = nullptr;
    maua_image_io img_io;
    float factors[MAUA_MAX_TAPS + 2];
    int launches_fwd = 0, launches_bwd = 0;
    // optional per-launch timing (bench.py roofline): one event after every launch group on the caller's stream
    struct Prof { const char* name; int layer; double flops, bytes; cudaEvent_t ev; };
    bool profile = false;
    std::vector<Prof> prof;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
};

static void prof_mark(maua_plan* p, cudaStream_t st, const char* name, int layer, double flops, double bytes) {
    if (!p->profile) return;
    if (p->ev_used == p->ev_pool.size()) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        p->ev_pool.push_back(ev);
    }
    cudaEvent_t ev = p->ev_pool[p->ev_used++];
    cudaEventRecord(ev, st);
    p->prof.push_back({name, layer, flops, bytes, ev});
}

namespace {

constexpr int kIm2colK = 384;  // 3 * 11 * 11 = 363 columns of the NIN image layer's im2col matrix, padded to a multiple of 128

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// output extent of one entry given its input extent
void entry_extent(const Entry& e, int& h, int& w) {
    if (e.pool) {
        if (e.pool3) pool3_out_extent(h, w, &h, &w);
        else { h /= 2; w /= 2; }
    } else {
        h = h + 2 * e.pad >= e.ks ? (h + 2 * e.pad - e.ks) / e.stride + 1 : 0;
        w = w + 2 * e.pad >= e.ks ? (w + 2 * e.pad - e.ks) / e.stride + 1 : 0;
    }
}

int ensure_workspaces(maua_plan* p, int H, int W) {
    // extents per entry
    size_t need = 0, max_act = 0, need_bits = 0;
    int h = H, w = W;
    for (auto& e : p->entries) {
        entry_extent(e, h, w);
        MAUA_REQUIRE(h >= 1 && w >= 1, "image %dx%d is too small for this network (an activation would be empty)", H, W);
        e.H = h; e.W = w;
        const size_t n = (size_t)h * w * e.C;
        need += (n + 63) & ~size_t(63);
        if (!e.pool) need_bits += (n / 32 + 63) & ~size_t(63);
        else if (!e.pool3) need_bits += (n / 16 + 63) & ~size_t(63);  // arg-max codes: one byte per 4 channels
        if (n > max_act) max_act = n;
    }
    if (need_bits > p->bits_words) {
        p->ws_generation++;
        if (p->bits_arena) cudaFree(p->bits_arena);
        p->bits_arena = nullptr;
        p->bits_words = 0;
        if (cudaMalloc(&p->bits_arena, need_bits * sizeof(uint32_t)) != cudaSuccess) {
            cudaGetLastError();
            set_last_error("out of device memory allocating the ReLU sign bitmaps for %dx%d", H, W);
            return MAUA_ERR_OOM;
        }
        p->bits_words = need_bits;
    }
    if (need > p->arena_elems) {
        p->ws_generation++;
        if (p->arena) cudaFree(p->arena);
        p->arena = nullptr;
        p->arena_elems = 0;
        if (cudaMalloc(&p->arena, need * sizeof(float)) != cudaSuccess) {
            cudaGetLastError();
            set_last_error("out of device memory allocating %.2f GB of activations for %dx%d", need * 4e-9, H, W);
            return MAUA_ERR_OOM;
        }
        p->arena_elems = need;
    }
    if (max_act > p->gbuf_elems) {
        p->ws_generation++;
        for (int i = 0; i < 3; ++i) {
            if (p->gbuf[i]) cudaFree(p->gbuf[i]);
            p->gbuf[i] = nullptr;
        }
        p->gbuf_elems = 0;
        for (int i = 0; i < 3; ++i) {
            if (cudaMalloc(&p->gbuf[i], max_act * sizeof(float)) != cudaSuccess) {
                cudaGetLastError();
                set_last_error("out of device memory allocating gradient buffers for %dx%d", H, W);
                return MAUA_ERR_OOM;
            }
        }
        p->gbuf_elems = max_act;
    }
    if (!p->entries.empty() && p->entries[0].image_layer && p->entries[0].ks == 11) {
        const size_t need_i = (size_t)p->entries[0].H * p->entries[0].W * kIm2colK;
        if (need_i > p->im2col_elems) {
            p->ws_generation++;
            if (p->im2col_ws) cudaFree(p->im2col_ws);
            p->im2col_ws = nullptr;
            p->im2col_elems = 0;
            if (cudaMalloc(&p->im2col_ws, need_i * sizeof(float)) != cudaSuccess) {
                cudaGetLastError();
                set_last_error("out of device memory allocating the im2col workspace (%.2f GB) for %dx%d", need_i * 4e-9, H, W);
                return MAUA_ERR_OOM;
            }
            p->im2col_elems = need_i;
        }
    }
    size_t off = 0, boff = 0;
    for (auto& e : p->entries) {
        e.out = p->arena + off;
        off += ((size_t)e.H * e.W * e.C + 63) & ~size_t(63);
        e.bits = nullptr;
        e.codes = nullptr;
        if (!e.pool) {
            e.bits = p->bits_arena + boff;
            boff += ((size_t)e.H * e.W * e.C / 32 + 63) & ~size_t(63);
        } else if (!e.pool3) {
            e.codes = reinterpret_cast<uint8_t*>(p->bits_arena + boff);
            boff += ((size_t)e.H * e.W * e.C / 16 + 63) & ~size_t(63);
        }
    }
    return MAUA_OK;
}

}  // namespace

static int plan_create_impl(int device, const maua_net_desc* d, int begin, int end, maua_plan_t** out) {
    MAUA_REQUIRE(d && out, "maua_plan_create: null argument");
    MAUA_REQUIRE(d->n_entries >= 1 && d->n_entries <= MAUA_MAX_LAYERS, "maua_plan_create: bad n_entries %d", d->n_entries);
    MAUA_REQUIRE(d->n_taps >= 0 && d->n_taps <= MAUA_MAX_TAPS, "maua_plan_create: bad n_taps %d", d->n_taps);
    MAUA_REQUIRE(begin >= 0 && begin < end && end <= d->n_entries, "maua_plan_create_stage: bad entry range [%d, %d) of %d",
                 begin, end, d->n_entries);
    int rc = check_device_arch(device);
    if (rc) return rc;
    DeviceGuard guard(device);
    MAUA_REQUIRE(d->channels[0] > 0, "maua_plan_create: the network must start with a conv layer");
    MAUA_REQUIRE(d->channels[begin] > 0, "maua_plan_create_stage: a stage must begin with a conv layer (entry %d is a pool)", begin);

    maua_plan* p = new maua_plan();
    p->device = device;
    p->avg_pool = d->avg_pool;
    if (const char* f = getenv("MAUA_FUSE_POOL")) p->fuse_pool = atoi(f) != 0;
    if (const char* f = getenv("MAUA_SIDE_STREAM")) p->side_mode = atoi(f) < 0 ? -1 : (atoi(f) != 0);
    if (const char* f = getenv("MAUA_PREFETCH_W")) p->prefetch_weights = atoi(f) != 0;
    if (const char* f = getenv("MAUA_POOL_CODES")) p->pool_codes = atoi(f) != 0;
    // (created here, not at first use: a first forward pass may already run inside a stream capture)
    if (cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        set_last_error("maua_plan_create: cannot create the side stream / its events: %s", cudaGetErrorString(cudaGetLastError()));
        maua_plan_destroy(p);
        return MAUA_ERR_CUDA;
    }
    if (const char* f = getenv("MAUA_SPLITK")) p->splitk = atoi(f) != 0;
    if (const char* f = getenv("MAUA_CONV_TAIL")) p->conv_tail = atoi(f);
    p->begin = begin;
    p->last_stage = (end == d->n_entries);
    memset(&p->img_io, 0, sizeof(p->img_io));
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** ptr, size_t bytes) {
        if (e == cudaSuccess) {
            e = cudaMalloc(ptr, bytes);
            if (e == cudaSuccess) p->weight_bytes += bytes;
        }
    };
    int cin = 3, convs = 0;
    bool ok = true;
    for (int i = 0; i < end && ok; ++i) {
        if (i == begin) p->in_C = cin;
        if (i < begin) {  // entries of earlier stages: only track channel / conv counters
            if (d->channels[i] > 0) { cin = d->channels[i]; convs++; }
            continue;
        }
        Entry en;
        if (d->channels[i] == 0) {
            en.pool = true;
            en.pool3 = d->pool_kind == 1;
            en.C = cin;
            MAUA_REQUIRE(cin % 4 == 0, "pool over %d channels unsupported", cin);
        } else {
            en.cin = cin;
            en.cout = d->channels[i];
            en.C = en.cout;
            en.Cn = d->norm_channels[i] > 0 ? d->norm_channels[i] : en.cout;
            MAUA_REQUIRE(en.Cn <= en.cout, "conv entry %d: norm_channels %d exceeds channels %d", i, en.Cn, en.cout);
            en.conv_index = convs++;
            en.image_layer = (i == 0);
            {
                const int kind = d->conv_kind[i];
                if (kind == 1) { en.ks = 1; en.pad = 0; }
                else if (kind == 5) { en.ks = 5; en.pad = 2; }
                else if (kind == 11) { en.ks = 11; en.stride = 4; en.pad = 0; }
                else if (kind != 0) {
                    set_last_error("conv entry %d: unknown conv_kind %d (0: 3x3, 1: 1x1, 5: 5x5 pad 2, 11: 11x11 stride 4)", i, kind);
                    ok = false;
                    break;
                }
                if ((kind == 11) != (i == 0 && kind != 0)) {
                    set_last_error("conv entry %d: the 11x11 / 4 layer is the image layer and only that", i);
                    ok = false;
                    break;
                }
                if (i == 0 && kind != 0 && kind != 11) {
                    set_last_error("the image layer must be a 3x3 or the 11x11 / 4 convolution");
                    ok = false;
                    break;
                }
            }
            if (i > 0 && !(en.cin % 32 == 0 && en.cout % 64 == 0)) {
                set_last_error("conv %d: %d -> %d channels unsupported by the tcgen05 path (need Cin %% 32 == 0, Cout %% 64 == 0); "
                               "only VGG-16/19-shaped stacks are supported", en.conv_index, en.cin, en.cout);
                ok = false;
                break;
            }
            if (i == 0 && en.ks == 3 && !(en.cout % 16 == 0 && en.cout <= 128 && 256 % (en.cout / 16) == 0 && en.cout % 64 == 0)) {
                set_last_error("first conv: 3 -> %d channels unsupported", en.cout);
                ok = false;
                break;
            }
            if (!d->weights[i] || !d->biases[i]) {
                set_last_error("conv %d: null weight / bias pointer", en.conv_index);
                ok = false;
                break;
            }
            const size_t wn = (size_t)en.cout * en.cin * en.ks * en.ks;
            const bool gemm_layer = i > 0 && en.ks <= 5;  // 5x5, 3x3, 1x1: tcgen05 implicit GEMM on rounded GEMM-layout copies
            alloc((void**)&en.w_raw, wn * sizeof(float));
            alloc((void**)&en.bias, (size_t)en.cout * sizeof(float));
            if (gemm_layer) {
                alloc((void**)&en.wg, wn * sizeof(float));
                alloc((void**)&en.wd, wn * sizeof(float));
            }
            if (en.ks == 5) alloc((void**)&en.w_flip, wn * sizeof(float));
            // the checkpoint tensors may live on another device (the first stage's): peer-capable default copy
            if (e == cudaSuccess) e = cudaMemcpy(en.w_raw, d->weights[i], wn * sizeof(float), cudaMemcpyDefault);
            if (e == cudaSuccess) e = cudaMemcpy(en.bias, d->biases[i], (size_t)en.cout * sizeof(float), cudaMemcpyDefault);
            if (i == 0 && en.ks == 3) {
                alloc((void**)&en.wt1, (size_t)32 * en.cout * sizeof(float));
                if (e == cudaSuccess && conv_first_dgrad_prep_weights(en.w_raw, en.wt1, en.cout, 1, 0)) ok = false;
            }
            if (e == cudaSuccess && gemm_layer) {
                if (prep_weights_launch(en.w_raw, en.wg, en.cout, en.cin, 0, 1, 0, en.ks * en.ks) ||
                    prep_weights_launch(en.w_raw, en.wd, en.cout, en.cin, 1, 1, 0, en.ks * en.ks))
                    ok = false;
            }
            if (e == cudaSuccess && en.ks == 5 && conv_gen_flip_weights_launch(en.w_raw, en.w_flip, en.cout, en.cin, 5, 0)) ok = false;
            if (i == 0 && en.ks == 11) {
                alloc((void**)&en.w1g, (size_t)en.cout * kIm2colK * sizeof(float));
                alloc((void**)&en.w1t, (size_t)en.cout * kIm2colK * sizeof(float));
                if (e == cudaSuccess && im2col_weights_launch(en.w_raw, en.w1g, en.w1t, en.cout, 3 * 11 * 11, kIm2colK, 0)) ok = false;
            }
            cin = en.cout;
        }
        p->entries.push_back(en);
    }
    // taps (global indexing; taps that belong to another stage keep entry = -1 and are ignored by this plan)
    for (int t = 0; t < d->n_taps && ok && e == cudaSuccess; ++t) {
        Tap tp;
        tp.kind = d->tap_kind[t];
        tp.relu_index = d->tap_relu_index[t];
        bool in_net = false;
        {
            int convs_seen = 0;
            for (int i = 0; i < d->n_entries; ++i)
                if (d->channels[i] > 0 && convs_seen++ == tp.relu_index) in_net = true;
        }
        for (size_t i = 0; i < p->entries.size(); ++i)
            if (!p->entries[i].pool && p->entries[i].conv_index == tp.relu_index) tp.entry = (int)i;
        if (!in_net) {
            set_last_error("tap %d refers to relu index %d which is not in the network", t, tp.relu_index);
            ok = false;
            break;
        }
        if (t > 0 && (d->tap_relu_index[t] < d->tap_relu_index[t - 1])) {
            set_last_error("taps must be ordered by relu index");
            ok = false;
            break;
        }
        for (int u = 0; u < t; ++u)
            if (d->tap_relu_index[u] == tp.relu_index && d->tap_kind[u] == tp.kind) {
                set_last_error("two taps of the same kind on relu index %d", tp.relu_index);
                ok = false;
            }
        if (tp.entry >= 0) {
            tp.C = p->entries[tp.entry].cout;
            tp.Cn = p->entries[tp.entry].Cn;
            if (tp.kind == MAUA_TAP_STYLE) {
                if (tp.C % 64 != 0) {
                    set_last_error("style tap on %d channels unsupported", tp.C);
                    ok = false;
                    break;
                }
                const size_t cc = (size_t)tp.C * tp.C * sizeof(float);
                alloc((void**)&tp.gram, cc);
                alloc((void**)&tp.diff, cc);
                alloc((void**)&tp.aux_d, cc);
                alloc((void**)&tp.mean, tp.C * sizeof(float));
                alloc((void**)&tp.aux_bias, tp.C * sizeof(float));
                alloc(&tp.gram_ws, gram_workspace_bytes(tp.C));
            }
        }
        p->taps.push_back(tp);
    }
    alloc(&p->reduce_ws, maua_reduce_workspace_bytes());
    alloc((void**)&p->coef2, (MAUA_MAX_TAPS + 2) * sizeof(float));
    alloc((void**)&p->splitk_ws, conv_splitk_ws_bytes());
    alloc((void**)&p->splitk_flags, conv_splitk_flag_words() * sizeof(unsigned int));
    if (e == cudaSuccess && ok) e = cudaMemset(p->splitk_flags, 0, conv_splitk_flag_words() * sizeof(unsigned int));
    if (e == cudaSuccess && ok) e = cudaMemset(p->reduce_ws, 0, maua_reduce_workspace_bytes());
    if (e == cudaSuccess && ok) e = cudaDeviceSynchronize();
    if (e != cudaSuccess || !ok) {
        if (e != cudaSuccess) {
            set_last_error("maua_plan_create: CUDA failure: %s", cudaGetErrorString(e));
            cudaGetLastError();
        }
        maua_plan_destroy(p);
        return e == cudaErrorMemoryAllocation ? MAUA_ERR_OOM : (e != cudaSuccess ? MAUA_ERR_CUDA : MAUA_ERR_ARG);
    }
    *out = p;
    return MAUA_OK;
}

extern "C" {

MAUA_API int maua_plan_create(int device, const maua_net_desc* d, maua_plan_t** out) {
    MAUA_REQUIRE(d && out, "maua_plan_create: null argument");
    return plan_create_impl(device, d, 0, d->n_entries, out);
}

MAUA_API int maua_plan_create_stage(int device, const maua_net_desc* d, int entry_begin, int entry_end, maua_plan_t** out) {
    MAUA_REQUIRE(d && out, "maua_plan_create_stage: null argument");
    return plan_create_impl(device, d, entry_begin, entry_end, out);
}

MAUA_API void maua_plan_destroy(maua_plan_t* p) {
    if (!p) return;
    DeviceGuard guard(p->device);
    for (auto& e : p->entries) {
        cudaFree(e.w_raw); cudaFree(e.wg); cudaFree(e.wd); cudaFree(e.bias); cudaFree(e.wt1);
        cudaFree(e.wg32); cudaFree(e.wd32); cudaFree(e.wt1_32); cudaFree(e.w_flip); cudaFree(e.w1g); cudaFree(e.w1t);
    }
    for (auto& t : p->taps) {
        cudaFree(t.gram); cudaFree(t.diff); cudaFree(t.aux_d); cudaFree(t.mean); cudaFree(t.aux_bias); cudaFree(t.gram_ws);
    }
    cudaFree(p->arena);
    cudaFree(p->im2col_ws);
    cudaFree(p->bits_arena);
    for (int i = 0; i < 3; ++i) cudaFree(p->gbuf[i]);
    cudaFree(p->reduce_ws); cudaFree(p->coef2);
    cudaFree(p->splitk_ws); cudaFree(p->splitk_flags);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    if (p->side) cudaStreamDestroy(p->side);
    for (cudaEvent_t ev : p->ev_pool) cudaEventDestroy(ev);
    delete p;
}

MAUA_API size_t maua_plan_device_bytes(const maua_plan_t* p) {
    if (!p) return 0;
    return p->weight_bytes + (p->arena_elems + p->im2col_elems) * sizeof(float) + p->bits_words * sizeof(uint32_t) +
           3 * p->gbuf_elems * sizeof(float);
}

MAUA_API int maua_plan_workspace_generation(const maua_plan_t* p) { return p ? p->ws_generation : -1; }

MAUA_API int maua_plan_set_impl(maua_plan_t* p, int impl) {
    MAUA_REQUIRE(p && impl >= MAUA_IMPL_TC && impl <= MAUA_IMPL_FP32, "maua_plan_set_impl: bad arguments");
    if (impl == MAUA_IMPL_FP32) {
        // un-rounded GEMM-layout weights for the exact-arithmetic kernels, made once
        DeviceGuard guard(p->device);
        for (auto& e : p->entries) {
            if (e.pool) continue;
            const size_t wn = (size_t)e.cout * e.cin * e.ks * e.ks;
            if (e.ks > 3) continue;  // direct fp32 convolutions (conv_gen.cu) read the raw weights in every mode
            if (e.image_layer) {
                if (e.wt1_32) continue;
                MAUA_CUDA_CHECK(cudaMalloc(&e.wt1_32, (size_t)32 * e.cout * sizeof(float)));
                int rc = conv_first_dgrad_prep_weights(e.w_raw, e.wt1_32, e.cout, 0, 0);
                if (rc) return rc;
            } else {
                if (e.wg32) continue;
                MAUA_CUDA_CHECK(cudaMalloc(&e.wg32, wn * sizeof(float)));
                MAUA_CUDA_CHECK(cudaMalloc(&e.wd32, wn * sizeof(float)));
                p->weight_bytes += 2 * wn * sizeof(float);
                int rc = prep_weights_launch(e.w_raw, e.wg32, e.cout, e.cin, 0, 0, 0, e.ks * e.ks);
                if (!rc) rc = prep_weights_launch(e.w_raw, e.wd32, e.cout, e.cin, 1, 0, 0, e.ks * e.ks);
                if (rc) return rc;
            }
        }
        MAUA_CUDA_CHECK(cudaDeviceSynchronize());
    }
    p->impl = impl;
    return MAUA_OK;
}

MAUA_API int maua_plan_set_fuse_pool(maua_plan_t* p, int enable) {
    MAUA_REQUIRE(p, "maua_plan_set_fuse_pool: null plan");
    p->fuse_pool = enable != 0;
    return MAUA_OK;
}

MAUA_API int maua_plan_set_side_stream(maua_plan_t* p, int mode) {
    MAUA_REQUIRE(p && mode >= -1 && mode <= 1, "maua_plan_set_side_stream: mode must be -1 (automatic), 0 or 1");
    p->side_mode = mode;
    return MAUA_OK;
}

MAUA_API int maua_plan_set_splitk(maua_plan_t* p, int enable) {
    MAUA_REQUIRE(p, "maua_plan_set_splitk: null plan");
    p->splitk = enable != 0;
    return MAUA_OK;
}

MAUA_API int maua_plan_set_conv_tail(maua_plan_t* p, int mode) {
    MAUA_REQUIRE(p && mode >= 0 && mode <= 3, "maua_plan_set_conv_tail: bad arguments");
    p->splitk = mode == 1;
    p->conv_tail = mode == 1 ? 0 : mode;
    return MAUA_OK;
}

MAUA_API int maua_plan_set_profile(maua_plan_t* p, int enable) {
    MAUA_REQUIRE(p, "maua_plan_set_profile: null plan");
    p->profile = enable != 0;
    p->prof.clear();
    p->ev_used = 0;
    return MAUA_OK;
}

MAUA_API long maua_plan_profile_json(maua_plan_t* p, char* buf, long cap, maua_stream_t stream) {
    if (!p || !buf || cap < 3) return -1;
    cudaStreamSynchronize((cudaStream_t)stream);
    std::string out = "[";
    for (size_t i = 1; i < p->prof.size(); ++i) {
        const auto& r = p->prof[i];
        if (!strncmp(r.name, "begin", 5)) continue;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p->prof[i - 1].ev, r.ev) != cudaSuccess) { cudaGetLastError(); continue; }
        char line[256];
        snprintf(line, sizeof(line), "%s{\"name\":\"%s\",\"layer\":%d,\"ms\":%.6f,\"flops\":%.6e,\"bytes\":%.6e}",
                 out.size() > 1 ? "," : "", r.name, r.layer, ms, r.flops, r.bytes);
        out += line;
    }
    out += "]";
    if ((long)out.size() + 1 > cap) return (long)out.size() + 1;
    memcpy(buf, out.c_str(), out.size() + 1);
    return (long)out.size() + 1;
}

MAUA_API int maua_plan_last_launches(const maua_plan_t* p, int* fwd, int* bwd) {
    MAUA_REQUIRE(p, "maua_plan_last_launches: null plan");
    if (fwd) *fwd = p->launches_fwd;
    if (bwd) *bwd = p->launches_bwd;
    return MAUA_OK;
}

}  // extern "C"

// `image` is the NCHW image for a plan that starts at conv1_1, else the NHWC [H][W][in_C] boundary activation produced
// by the previous stage.  boundary_out (optional, may be peer memory) additionally receives the last entry's output.
static int plan_forward_impl(maua_plan_t* p, const float* image, int H, int W, const maua_tap_io* tio,
                             const maua_image_io* iio, float* losses_out, int keep_for_backward, float* boundary_out,
                             maua_stream_t stream) {
    MAUA_REQUIRE(p && image && H >= 1 && W >= 1, "maua_plan_forward: bad arguments");
    MAUA_REQUIRE(p->taps.empty() || tio, "maua_plan_forward: tap io missing");
    MAUA_REQUIRE(p->begin == 0 || (reinterpret_cast<uintptr_t>(image) & 15) == 0, "stage input must be 16-byte aligned");
    DeviceGuard guard(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_workspaces(p, H, W);
    if (rc) return rc;
    p->can_backward = false;
    p->H = H; p->W = W; p->image = image;
    p->stage_input = p->begin > 0 ? image : nullptr;
    if (p->begin > 0) iio = nullptr;  // TVLoss / temporal ContentLoss sit on the image: first stage only
    p->launches_fwd = 0;
    const int nt = (int)p->taps.size();
    ReduceScratch rs = scratch_from_workspace(p->reduce_ws);
    if (losses_out) {
        MAUA_CUDA_CHECK(cudaMemsetAsync(losses_out, 0, (nt + 2) * sizeof(float), st));
    }
    memset(p->factors, 0, sizeof(p->factors));
    if (iio) p->img_io = *iio; else memset(&p->img_io, 0, sizeof(p->img_io));
    const long img_elems = 3L * H * W;
    const bool exact = p->impl == MAUA_IMPL_FP32;  // nothing is rounded to TF32 in the exact-arithmetic mode
    const int rnd = exact ? 0 : 1;
    p->prof.clear();
    p->ev_used = 0;
    prof_mark(p, st, "begin_fwd", -1, 0, 0);

    // ---- image-side modules: TVLoss, temporal ContentLoss ----
    bool tv_pending = false;  // TVLoss value: folded into conv1_1's forward (which reads the image anyway) when that runs
    if (p->img_io.tv_mode == MAUA_MODE_LOSS) {
        MAUA_REQUIRE(losses_out, "losses_out is required in loss mode");
        tv_pending = true;
        p->factors[nt] = 1.f;
    }
    bool temporal_active = false;
    if (p->img_io.temporal_mode == MAUA_MODE_CAPTURE) {
        MAUA_REQUIRE(p->img_io.temporal_target && p->img_io.temporal_target_elems == img_elems,
                     "temporal capture needs a target buffer of %ld elements", img_elems);
        MAUA_CUDA_CHECK(cudaMemcpyAsync(p->img_io.temporal_target, image, img_elems * sizeof(float),
                                        cudaMemcpyDeviceToDevice, st));
    } else if (p->img_io.temporal_mode == MAUA_MODE_LOSS && p->img_io.temporal_target &&
               p->img_io.temporal_target_elems == img_elems) {
        MAUA_REQUIRE(losses_out, "losses_out is required in loss mode");
        // loss.py:53-54: MSE(x * weights, target), weights [1,1,H,W] broadcast over channels
        if ((rc = wmse_value_launch(image, p->img_io.temporal_weights, p->img_io.temporal_target, img_elems, (long)H * W,
                                    p->img_io.temporal_strength / (float)img_elems, losses_out + nt + 1, rs, st)))
            return rc;
        p->launches_fwd++;
        prof_mark(p, st, "temporal_value", -1, 0, 8.0 * img_elems);
        temporal_active = true;
        p->factors[nt + 1] = 2.f / (float)img_elems;
    }
    if (!temporal_active) p->img_io.temporal_mode = MAUA_MODE_NONE;

    // ---- which taps are live, and where the forward may stop (models.py:382 truncation, per call) ----
    int last_needed = -1;
    for (int t = 0; t < nt; ++t) {
        Tap& tp = p->taps[t];
        tp.mode = tp.entry >= 0 ? tio[t].mode : MAUA_MODE_NONE;
        tp.active = false;
        tp.use_cov = tio[t].use_covariance;
        tp.target = tio[t].target;
        if (tp.mode != MAUA_MODE_NONE) last_needed = tp.entry > last_needed ? tp.entry : last_needed;
    }
    const int n_ent = (int)p->entries.size();
    if (boundary_out) last_needed = n_ent - 1;  // the next stage consumes this stage's last activation
    p->last_entry = last_needed;
    if (tv_pending && !(last_needed >= 0 && p->entries[0].image_layer && p->entries[0].ks == 3)) {
        if ((rc = tv_value_launch(image, 3, H, W, p->img_io.tv_strength, losses_out + nt, rs, st))) return rc;
        p->launches_fwd++;
        prof_mark(p, st, "tv_value", -1, 0, 4.0 * img_elems);
        tv_pending = false;
    }

    // ---- side stream for the loss modules (see maua_plan::side_mode) ----
    constexpr long kSideAutoPixels = 640L * 640L;
    const bool use_side = !p->profile && (p->side_mode == 1 || (p->side_mode < 0 && (long)H * W <= kSideAutoPixels));
    bool forked = false;
    // stream the loss modules of the tap that has just been produced on `st` are launched on
    cudaError_t fork_err = cudaSuccess;
    auto loss_stream = [&]() -> cudaStream_t {
        if (!use_side) return st;
        cudaError_t e1 = cudaEventRecord(p->ev_fork, st);
        if (e1 == cudaSuccess) e1 = cudaStreamWaitEvent(p->side, p->ev_fork, 0);
        if (e1 != cudaSuccess) { fork_err = e1; return st; }  // (reported below; the work then simply stays on the caller's stream)
        forked = true;
        return p->side;
    };

    // ---- feature stack ----
    const float* cur = p->stage_input;
    int curH = H, curW = W;
    bool pool_done = false;
    for (int i = 0; i <= last_needed; ++i) {
        Entry& e = p->entries[i];
        float* hand_off = (boundary_out && i == n_ent - 1) ? boundary_out : nullptr;
        if (e.image_layer && e.ks != 3 && !exact) {
            // NIN conv1 (models.py:83: 11x11 / stride 4 on the NCHW image) on the product path: im2col (363 -> 384 columns, TF32-
            // rounded) + the pointwise tcgen05 GEMM with the fused bias / ReLU / sign-bitmap epilogue
            if ((rc = im2col_img_launch(image, p->im2col_ws, H, W, e.ks, e.stride, kIm2colK, 1, st))) return rc;
            p->launches_fwd++;
            ConvArgs a;
            a.B = 1; a.H = e.H; a.W = e.W; a.Cin = kIm2colK; a.Cout = e.cout; a.ntaps = 1;
            a.in = p->im2col_ws; a.wg = e.w1g;
            a.ep.out = e.out; a.ep.bias = e.bias; a.ep.relu = 1; a.ep.round = rnd; a.ep.mask_out = e.bits; a.ep.out2 = hand_off;
            a.tail_mode = p->conv_tail;
            if ((rc = conv_dispatch(a, p->impl, st))) return rc;
        } else if (e.image_layer && e.ks != 3) {
            // ... and in the exact-arithmetic mode: direct fp32 convolution
            if ((rc = conv_gen_fwd_launch(image, 1, e.w_raw, e.bias, e.out, e.bits, 1, H, W, 3, e.cout, e.ks, e.stride, e.pad, 1, rnd, st)))
                return rc;
            if (hand_off)
                MAUA_CUDA_CHECK(cudaMemcpyAsync(hand_off, e.out, (size_t)e.H * e.W * e.C * sizeof(float), cudaMemcpyDefault, st));
        } else if (e.image_layer) {
            ConvFirstTV tv;
            if (tv_pending) {
                tv.strength = p->img_io.tv_strength; tv.out = losses_out + nt;
                tv.partials = rs.partials; tv.counter = rs.counter; tv.max_blocks = rs.max_blocks;
            }
            // tcgen05 kernel (3xTF32) on the product path; the FFMA kernel in exact mode and for the SIMT cross-check plan
            if ((rc = conv_first_fwd_launch(image, e.w_raw, e.bias, e.out, e.bits, 1, H, W, e.cout, rnd, st, tv_pending ? &tv : nullptr,
                                            exact || p->impl == MAUA_IMPL_REF)))
                return rc;
            if (hand_off)
                MAUA_CUDA_CHECK(cudaMemcpyAsync(hand_off, e.out, (size_t)e.H * e.W * e.C * sizeof(float), cudaMemcpyDefault, st));
        } else if (e.pool) {
            if (pool_done) {  // produced by the epilogue of the conv below (fused pooling)
                pool_done = false;
                cur = e.out; curH = e.H; curW = e.W;
                continue;
            }
            // a pooled map is not needed again by this stage (the backward pass reads the pre-pool activation), so at a
            // stage boundary it is written straight into the next stage's memory
            if (e.pool3) {
                if ((rc = pool3_fwd_launch(cur, hand_off ? hand_off : e.out, 1, curH, curW, e.C, p->avg_pool, rnd, st))) return rc;
            } else {
                e.codes_ok = p->pool_codes && !p->avg_pool && !exact && p->impl != MAUA_IMPL_REF && e.codes && e.C % 32 == 0;
                if ((rc = pool_fwd_launch(cur, hand_off ? hand_off : e.out, 1, curH, curW, e.C, p->avg_pool, rnd, st,
                                          e.codes_ok ? e.codes : nullptr)))
                    return rc;
            }
        } else if (e.ks == 5 && exact) {
            // NIN conv2 (models.py:90) in the exact-arithmetic mode: 5x5 / pad 2 as a direct fp32 convolution (conv_gen.cu); the
            // product path runs it through conv_tc_kernel<.., KS = 5> below
            if ((rc = conv_gen_fwd_launch(cur, 0, e.w_raw, e.bias, e.out, e.bits, 1, curH, curW, e.cin, e.cout, 5, 1, 2, 1, rnd, st)))
                return rc;
            if (hand_off)
                MAUA_CUDA_CHECK(cudaMemcpyAsync(hand_off, e.out, (size_t)e.H * e.W * e.C * sizeof(float), cudaMemcpyDefault, st));
        } else {
            ConvArgs a;
            a.B = 1; a.H = e.H; a.W = e.W; a.Cin = e.cin; a.Cout = e.cout; a.ntaps = e.ks * e.ks;
            a.in = cur; a.wg = exact ? e.wg32 : e.wg;
            a.ep.out = e.out; a.ep.bias = e.bias; a.ep.relu = 1; a.ep.round = rnd;
            a.ep.mask_out = e.bits;
            if (p->splitk || p->conv_tail == 3) { a.splitk_ws = p->splitk_ws; a.splitk_flags = p->splitk_flags; }
            a.tail_mode = p->splitk ? 1 : p->conv_tail;
            a.ep.out2 = hand_off;  // dual store: tile by tile into the peer's memory while the GEMM runs
            if (p->prefetch_weights && !exact) {
                // weights of the next launch of the iteration: the next conv of the stack, or -- from the last layer -- its own
                // rotated copy, which the backward pass starts with
                int j = i + 1;
                while (j <= last_needed && p->entries[j].pool) ++j;
                const Entry* nx = j <= last_needed ? &p->entries[j] : (keep_for_backward ? &e : nullptr);
                if (nx && !nx->image_layer && nx->ks == 3) {
                    a.prefetch = j <= last_needed ? nx->wg : nx->wd;
                    a.prefetch_bytes = (size_t)nx->cout * nx->cin * 9 * sizeof(float);
                }
            }
            bool split_layer = false;  // tail mode 3: a launch of fewer tiles than SMs is K-split + reduced; its pool stays separate
            if (a.tail_mode == 3 && p->impl != MAUA_IMPL_REF && !exact) {
                int bn, mt, cg, full, st_, sp_;
                conv_tile_plan(a, 148, 3, &bn, &mt, &cg, &full, &st_, &sp_);
                split_layer = full == 0 && sp_ > 1 && st_ > 0;
                // a layer that feeds a pool keeps the pooling fused unless the split is deep enough to pay for a separate pool pass
                if (split_layer && sp_ < 8 && i + 1 <= last_needed && p->entries[i + 1].pool) { split_layer = false; a.tail_mode = 2; }
            }
            if (p->fuse_pool && !split_layer && !boundary_out && i + 1 <= last_needed && p->entries[i + 1].pool && !p->entries[i + 1].pool3 &&
                e.ks == 3 && e.H >= 2 && e.W >= 2) {
                a.ep.pool_out = p->entries[i + 1].out;  // models.py:119-122 pooled from the accumulator registers
                a.ep.pool_avg = p->avg_pool;
                Entry& pe = p->entries[i + 1];
                pe.codes_ok = p->pool_codes && !p->avg_pool && !exact && p->impl != MAUA_IMPL_REF && pe.codes && e.C % 32 == 0;
                a.ep.pool_codes = pe.codes_ok ? pe.codes : nullptr;
                pool_done = true;
            }
            if ((rc = conv_dispatch(a, p->impl, st))) return rc;
        }
        p->launches_fwd++;
        {
            const double px = (double)e.H * e.W;
            if (e.image_layer) prof_mark(p, st, "conv_first_fwd", i, 2.0 * 3 * e.ks * e.ks * e.cout * px, 4.0 * (3.0 * H * W + e.cout * px));
            else if (e.pool) prof_mark(p, st, "pool_fwd", i, 0, 4.0 * 5 * e.C * px);
            else prof_mark(p, st, "conv_fwd", i, 2.0 * e.ks * e.ks * e.cin * e.cout * px, 4.0 * (e.cin + e.cout) * px);
        }
        cur = e.out; curH = e.H; curW = e.W;
        if (e.pool) continue;
        // loss modules spliced after this ReLU (models.py:403-431)
        for (int t = 0; t < nt; ++t) {
            Tap& tp = p->taps[t];
            if (tp.entry != i || tp.mode == MAUA_MODE_NONE) continue;
            const long P = (long)e.H * e.W;
            const long numel = P * e.C;
            const float numel_n = (float)(P * e.Cn);  // nn.MSELoss mean over the REAL elements (padded channels are zero)
            const bool has_work = tp.kind == MAUA_TAP_STYLE ? tp.mode != MAUA_MODE_EXTERNAL
                                                            : (tp.mode != MAUA_MODE_CAPTURE && tp.target && tio[t].target_elems == numel);
            cudaStream_t ls = has_work ? loss_stream() : st;
            if (tp.kind == MAUA_TAP_STYLE && tp.mode == MAUA_MODE_EXTERNAL) {
                // the feature map stays in the arena; loss value and backward term come from the caller
                tp.active = true;
                tp.ext_in2 = nullptr; tp.ext_K2 = 0; tp.ext_w2 = nullptr; tp.ext_bias = nullptr;
                p->factors[t] = 1.f;
            } else if (tp.kind == MAUA_TAP_STYLE) {
                MAUA_REQUIRE(tp.target && tio[t].target_elems == (long)tp.C * tp.C,
                             "style tap %d: target must be a [%d,%d] device tensor", t, tp.C, tp.C);
                GramLossFuse fuse;
                const bool loss_mode = tp.mode != MAUA_MODE_CAPTURE;
                if (loss_mode) {
                    MAUA_REQUIRE(losses_out, "losses_out is required in loss mode");
                    fuse.target = tp.target; fuse.diff = tp.diff; fuse.loss_out = losses_out + t;
                    fuse.value_scale = tio[t].value_scale; fuse.rs = rs;
                }
                // one SYRK + one finalize kernel (which also produces G - A and the loss value in loss mode)
                if ((rc = gram_launch(e.out, P, tp.C, tp.use_cov, tp.gram, tp.mean, tp.gram_ws, p->impl, ls,
                                      loss_mode ? &fuse : nullptr, tp.Cn)))
                    return rc;
                p->launches_fwd += 2 + (tp.use_cov ? 2 : 0);
                prof_mark(p, st, "gram_syrk", i, (double)tp.C * (tp.C + 1) * P, 4.0 * P * tp.C);
                if (!loss_mode) {
                    // loss.py:146-151: target (+)= blend_weight * gram   (B = 1)
                    if ((rc = axpby_launch(tp.gram, tp.target, (long)tp.C * tp.C, tio[t].capture_weight,
                                           tio[t].capture_accumulate, ls)))
                        return rc;
                    p->launches_fwd++;
                } else {
                    tp.active = true;
                    p->factors[t] = 1.f;
                }
            } else {
                if (tp.mode == MAUA_MODE_CAPTURE) {
                    // loss.py:61-62: target = input.detach()
                    MAUA_REQUIRE(tp.target && tio[t].target_elems == numel,
                                 "content tap %d: capture needs a target buffer of %ld elements", t, numel);
                    MAUA_CUDA_CHECK(cudaMemcpyAsync(tp.target, e.out, numel * sizeof(float), cudaMemcpyDeviceToDevice, st));
                } else if (tp.target && tio[t].target_elems == numel) {  // loss.py:44: silently skipped on shape mismatch
                    MAUA_REQUIRE(losses_out, "losses_out is required in loss mode");
                    if ((rc = mse_value_launch(e.out, tp.target, numel, tio[t].value_scale / numel_n, losses_out + t,
                                               rs, ls)))
                        return rc;
                    p->launches_fwd++;
                    prof_mark(p, st, "content_loss", i, 0, 8.0 * numel);
                    tp.active = true;
                    p->factors[t] = 2.f / numel_n;
                }
            }
        }
    }
    if (forked) {  // everything behind the forward pass on `st` (loss read-back, backward pass) sees the loss modules' results
        MAUA_CUDA_CHECK(cudaEventRecord(p->ev_join, p->side));
        MAUA_CUDA_CHECK(cudaStreamWaitEvent(st, p->ev_join, 0));
    }
    MAUA_CUDA_CHECK(fork_err);
    p->can_backward = keep_for_backward != 0;
    return MAUA_OK;
}

extern "C" {

MAUA_API int maua_plan_forward(maua_plan_t* p, const float* image, int H, int W, const maua_tap_io* tio,
                               const maua_image_io* iio, float* losses_out, int keep_for_backward,
                               maua_stream_t stream) {
    MAUA_REQUIRE(p && p->begin == 0 && p->last_stage,
                 "maua_plan_forward: this plan is one stage of a layer-wise split, use maua_plan_forward_stage");
    return plan_forward_impl(p, image, H, W, tio, iio, losses_out, keep_for_backward, nullptr, stream);
}

MAUA_API int maua_plan_forward_stage(maua_plan_t* p, const float* input, int H, int W, const maua_tap_io* tio,
                                     const maua_image_io* iio, float* losses_out, int keep_for_backward,
                                     float* boundary_out, maua_stream_t stream) {
    MAUA_REQUIRE(p, "maua_plan_forward_stage: null plan");
    return plan_forward_impl(p, input, H, W, tio, iio, losses_out, keep_for_backward, boundary_out, stream);
}

MAUA_API int maua_loss_grad_coefs(const float* upstream, float* coefs, int n, const float* strength, const float* vsf,
                                  const int* normalize, const int* kind, maua_stream_t stream) {
    MAUA_REQUIRE(upstream && coefs && strength && vsf && normalize && kind && n >= 1 && n <= MAUA_MAX_TAPS + 2,
                 "maua_loss_grad_coefs: bad arguments");
    CoefParams cp;
    cp.n = n;
    for (int i = 0; i < n; ++i) {
        cp.strength[i] = strength[i]; cp.vsf[i] = vsf[i]; cp.normalize[i] = normalize[i]; cp.kind[i] = kind[i];
    }
    MAUA_CUDA_CHECK(launch_pdl(loss_grad_coefs_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, upstream, coefs, cp));
    return MAUA_OK;
}

}  // extern "C"

// grad_top (optional): d(sum of the later stages' losses) / d(this stage's last activation), produced by the next stage
// of a layer-wise split.  grad_image: NCHW image gradient for a plan that starts at conv1_1, else the NHWC gradient
// w.r.t. the stage input (unmasked, unrounded: the previous stage owns the activation and applies the ReLU mask).
static int plan_backward_impl(maua_plan_t* p, const float* grad_coefs, const float* grad_top, float* grad_image,
                              maua_stream_t stream) {
    MAUA_REQUIRE(p && grad_coefs && grad_image, "maua_plan_backward: null argument");
    if (!p->can_backward) {
        set_last_error("maua_plan_backward: no forward pass with keep_for_backward is pending");
        return MAUA_ERR_STATE;
    }
    DeviceGuard guard(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int nt = (int)p->taps.size();
    const int H = p->H, W = p->W;
    p->launches_bwd = 0;
    int rc;
    const bool exact = p->impl == MAUA_IMPL_FP32;
    const int rnd = exact ? 0 : 1;
    prof_mark(p, st, "begin_bwd", -1, 0, 0);

    // scaled coefficients (content / temporal: 2/numel) and the style taps' scaled (G - A) matrices, one launch
    {
        BwdPrep bp;
        memset(&bp, 0, sizeof(bp));
        bp.n_slots = nt + 2;
        bp.do_round = rnd;
        for (int i = 0; i < nt + 2; ++i) bp.factors[i] = p->factors[i];
        for (int t = 0; t < nt; ++t) {
            Tap& tp = p->taps[t];
            if (!tp.active || tp.kind != MAUA_TAP_STYLE) continue;
            if (tp.mode == MAUA_MODE_EXTERNAL) {
                MAUA_REQUIRE(tp.ext_in2 && tp.ext_w2 && tp.ext_K2 > 0,
                             "style tap %d is in external mode but maua_plan_set_tap_fold was not called after the forward pass", t);
                continue;
            }
            const Entry& e = p->entries[tp.entry];
            const int j = bp.n_style++;
            bp.diff[j] = tp.diff; bp.aux_d[j] = tp.aux_d; bp.C[j] = tp.C; bp.slot[j] = t;
            bp.inv_c3p[j] = 4.f / ((float)tp.Cn * (float)tp.Cn * (float)tp.Cn * (float)((long)e.H * e.W));
        }
        MAUA_CUDA_CHECK(launch_pdl(bwd_prep_kernel, dim3(128, bp.n_style > 0 ? bp.n_style : 1), dim3(256), 0, st, grad_coefs, p->coef2, bp));
        p->launches_bwd++;
        for (int t = 0; t < nt; ++t) {  // covariance: aux_bias = -aux_d @ mean (loss.py:87-89)
            Tap& tp = p->taps[t];
            if (!tp.active || tp.kind != MAUA_TAP_STYLE || !tp.use_cov || tp.mode == MAUA_MODE_EXTERNAL) continue;
            if ((rc = style_loss_bwd_bias_launch(tp.aux_d, tp.mean, tp.C, tp.aux_bias, st))) return rc;
            p->launches_bwd++;
        }
    }
    prof_mark(p, st, "bwd_prep", -1, 0, 0);

    auto taps_at = [&](int entry, Tap*& style, Tap*& content, int& content_idx) {
        style = content = nullptr;
        content_idx = -1;
        for (int t = 0; t < nt; ++t) {
            Tap& tp = p->taps[t];
            if (tp.entry != entry || !tp.active) continue;
            if (tp.kind == MAUA_TAP_STYLE) style = &tp;
            else { content = &tp; content_idx = t; }
        }
    };
    auto add_taps = [&](ConvArgs& a, const Entry& e, Tap* style, Tap* content, int content_idx) {
        if (style && style->mode == MAUA_MODE_EXTERNAL) {
            a.K2 = style->ext_K2; a.in2 = style->ext_in2; a.w2 = style->ext_w2;
            a.ep.bias = style->ext_bias;
        } else if (style) {
            a.K2 = e.C; a.in2 = e.out; a.w2 = style->aux_d;
            a.ep.bias = style->use_cov ? style->aux_bias : nullptr;
        }
        if (content) {
            a.ep.cont_f = e.out; a.ep.cont_t = content->target; a.ep.cont_coef = p->coef2 + content_idx;
        }
    };
    auto run_conv = [&](ConvArgs& a) -> int {
        p->launches_bwd++;
        if (p->splitk || p->conv_tail == 3) { a.splitk_ws = p->splitk_ws; a.splitk_flags = p->splitk_flags; }
        a.tail_mode = p->splitk ? 1 : p->conv_tail;
        const int r = conv_dispatch(a, p->impl, st);
        const double px = (double)a.H * a.W;
        // algorithmic work: dgrad GEMM + StyleLoss backward GEMM; bytes: gradient in + out, mask / feature read
        prof_mark(p, st, a.ntaps ? "conv_dgrad" : "tap_grad", a.Cout, 2.0 * (a.ntaps * (double)a.Cin + a.K2) * a.Cout * px,
                  4.0 * ((a.ntaps ? a.Cin : 0) + 2.0 * a.Cout) * px);
        return r;
    };

    // un-pool + ReLU mask (+ tap gradients of the pre-pool layer) through the pool entry that follows conv entry `prod`
    auto unpool = [&](const Entry& ep_, const Entry& epool, const float* gpool, const float* addend, float* outb) -> int {
        p->launches_bwd++;
        // 2x2 pools of the product path un-pool from the forward pass's arg-max codes (max) and the sign bitmap of the pre-pool
        // activation instead of the activation itself: same decisions, bit-identical gradient, 42 % less traffic
        const bool compact = !epool.pool3 && p->pool_codes && !exact && p->impl != MAUA_IMPL_REF && ep_.bits && ep_.C % 32 == 0 &&
                             (p->avg_pool || epool.codes_ok);
        const int r = epool.pool3 ? pool3_bwd_launch(ep_.out, gpool, addend, outb, 1, ep_.H, ep_.W, ep_.C, p->avg_pool, rnd, st)
                      : compact   ? pool_bwd_codes_launch(ep_.bits, epool.codes, gpool, addend, outb, 1, ep_.H, ep_.W, ep_.C, p->avg_pool, rnd, st)
                                  : pool_bwd_launch(ep_.out, gpool, addend, outb, 1, ep_.H, ep_.W, ep_.C, p->avg_pool, rnd, st);
        prof_mark(p, st, "pool_bwd", ep_.C, 0, 4.0 * 2.25 * ep_.C * ep_.H * ep_.W);
        return r;
    };

    // Walk down from the last executed conv.  `gm` = Gm of the conv entry above the current position.
    float* gm = nullptr;       // masked gradient w.r.t. the output of entry `gm_entry`
    int gm_entry = -1;
    int next_buf = 0;
    auto take_buf = [&]() { float* b = p->gbuf[next_buf]; next_buf = (next_buf + 1) % 3; return b; };

    int e_idx = -1;  // highest entry with a live loss module
    for (int t = 0; t < nt; ++t)
        if (p->taps[t].active && p->taps[t].entry > e_idx) e_idx = p->taps[t].entry;
    if (grad_top) {
        MAUA_REQUIRE(p->last_entry == (int)p->entries.size() - 1, "maua_plan_backward_stage: grad_top given but the last "
                     "forward did not run to the end of the stage");
        e_idx = p->last_entry;
    }
    if (e_idx >= 0 && p->entries[e_idx].pool) {
        // (stage boundary after a pool) grad_top is the gradient w.r.t. the pooled map: un-pool + mask (+ tap gradients)
        const int prod = e_idx - 1;
        MAUA_REQUIRE(prod >= 0 && !p->entries[prod].pool, "unsupported network: two pools in a row");
        Entry& ep_ = p->entries[prod];
        Tap *style, *content; int ci;
        taps_at(prod, style, content, ci);
        float* addend = nullptr;
        if (style || content) {
            ConvArgs t;
            t.B = 1; t.H = ep_.H; t.W = ep_.W; t.Cin = 32; t.Cout = ep_.C; t.ntaps = 0;
            add_taps(t, ep_, style, content, ci);
            t.ep.out = take_buf(); t.ep.round = 0;
            if (style) { if ((rc = run_conv(t))) return rc; }
            else { if ((rc = conv_ref_launch(t, st))) return rc; p->launches_bwd++; }
            addend = t.ep.out;
        }
        float* outb = take_buf();
        if ((rc = unpool(ep_, p->entries[e_idx], grad_top, addend, outb))) return rc;
        gm = outb;
        gm_entry = prod;
    } else if (e_idx >= 0) {
        // top of the stack: only tap gradients (+ the gradient handed down by the next stage)
        Entry& e = p->entries[e_idx];
        Tap *style, *content; int ci;
        taps_at(e_idx, style, content, ci);
        if (style || content || grad_top) {
            ConvArgs a;
            a.B = 1; a.H = e.H; a.W = e.W; a.Cin = 32; a.Cout = e.C; a.ntaps = 0; a.K2 = 0;
            add_taps(a, e, style, content, ci);
            a.ep.out = take_buf(); a.ep.mask_bits = e.bits; a.ep.round = rnd; a.ep.addend = grad_top;
            if (style) {
                if ((rc = run_conv(a))) return rc;
            } else {
                // no GEMM term (content-only tap and / or a handed-down gradient): element-wise SIMT epilogue
                if ((rc = conv_ref_launch(a, st))) return rc;
                p->launches_bwd++;
                prof_mark(p, st, "tap_grad", e.C, 0, 16.0 * e.C * e.H * e.W);
            }
            gm = a.ep.out;
            gm_entry = e_idx;
        }
    }
    // descend
    for (int c = gm_entry; c > 0 && gm;) {
        // c is a conv entry with masked gradient gm; produce Gm of the conv entry below it
        Entry& ec = p->entries[c];
        int below = c - 1;
        const bool through_pool = p->entries[below].pool;
        const int prod = through_pool ? below - 1 : below;  // conv entry that produced conv c's input (maybe pooled)
        MAUA_REQUIRE(prod >= 0 && !p->entries[prod].pool, "unsupported network: two pools in a row");
        Entry& ep_ = p->entries[prod];
        Tap *style, *content; int ci;
        taps_at(prod, style, content, ci);
        ConvArgs a;
        a.B = 1; a.H = ec.H; a.W = ec.W;
        a.Cin = ec.cout; a.Cout = ec.cin; a.ntaps = ec.ks * ec.ks;
        a.in = gm; a.wg = exact ? ec.wd32 : ec.wd;
        if (p->prefetch_weights && !exact && prod > 0 && !ep_.image_layer && ep_.ks == 3) {
            a.prefetch = ep_.wd;  // the next dgrad's weights (see ConvArgs::prefetch)
            a.prefetch_bytes = (size_t)ep_.cout * ep_.cin * 9 * sizeof(float);
        }
        if (ec.ks == 5 && exact) {
            // NIN conv2 (5x5 / pad 2), exact mode: the input gradient is the direct convolution of Gm with the rotated, transposed weights
            // (conv_gen.cu, fp32); the ReLU mask, tap gradients and rounding follow in the un-pool kernel or an epilogue-only launch
            float* raw = take_buf();
            if ((rc = conv_gen_fwd_launch(gm, 0, ec.w_flip, nullptr, raw, nullptr, 1, ec.H, ec.W, ec.cout, ec.cin, 5, 1, 2, 0, 0, st)))
                return rc;
            p->launches_bwd++;
            prof_mark(p, st, "conv_dgrad", ec.cin, 2.0 * 25 * ec.cin * ec.cout * ec.H * ec.W, 4.0 * (ec.cin + ec.cout) * ec.H * ec.W);
            if (!through_pool) {
                ConvArgs t;
                t.B = 1; t.H = ep_.H; t.W = ep_.W; t.Cin = 32; t.Cout = ep_.C; t.ntaps = 0;
                add_taps(t, ep_, style, content, ci);
                t.ep.out = take_buf(); t.ep.mask_bits = ep_.bits; t.ep.round = rnd; t.ep.addend = raw;
                if (style) { if ((rc = run_conv(t))) return rc; }
                else { if ((rc = conv_ref_launch(t, st))) return rc; p->launches_bwd++; }
                gm = t.ep.out;
            } else {
                float* addend = nullptr;
                if (style || content) {
                    ConvArgs t;
                    t.B = 1; t.H = ep_.H; t.W = ep_.W; t.Cin = 32; t.Cout = ep_.C; t.ntaps = 0;
                    add_taps(t, ep_, style, content, ci);
                    t.ep.out = take_buf(); t.ep.round = 0;
                    if (style) { if ((rc = run_conv(t))) return rc; }
                    else { if ((rc = conv_ref_launch(t, st))) return rc; p->launches_bwd++; }
                    addend = t.ep.out;
                }
                float* outb = take_buf();
                if (outb == raw || outb == addend) outb = take_buf();
                if ((rc = unpool(ep_, p->entries[below], raw, addend, outb))) return rc;
                gm = outb;
            }
        } else if (!through_pool) {
            add_taps(a, ep_, style, content, ci);
            a.ep.out = take_buf(); a.ep.mask_bits = ep_.bits; a.ep.round = rnd;
            if ((rc = run_conv(a))) return rc;
            gm = a.ep.out;
        } else {
            // gradient w.r.t. the pooled map, unrounded; then un-pool + mask (+ tap gradients of the pre-pool layer)
            a.ep.out = take_buf(); a.ep.round = 0;
            if ((rc = run_conv(a))) return rc;
            float* gpool = a.ep.out;
            float* addend = nullptr;
            if (style || content) {
                ConvArgs t;
                t.B = 1; t.H = ep_.H; t.W = ep_.W; t.Cin = 32; t.Cout = ep_.C; t.ntaps = 0;
                add_taps(t, ep_, style, content, ci);
                t.ep.out = take_buf(); t.ep.round = 0;
                if (style) { if ((rc = run_conv(t))) return rc; }
                else { if ((rc = conv_ref_launch(t, st))) return rc; p->launches_bwd++; }
                addend = t.ep.out;
            }
            float* outb = take_buf();
            if (outb == gpool || outb == addend) outb = take_buf();
            if ((rc = unpool(ep_, p->entries[below], gpool, addend, outb))) return rc;
            gm = outb;
        }
        c = prod;
        gm_entry = prod;
    }

    if (p->begin > 0) {
        // stage of a layer-wise split: hand d/d(stage input) to the previous stage (may be a peer-memory store)
        Entry& e0 = p->entries[0];
        const size_t in_elems = (size_t)e0.H * e0.W * e0.cin;
        if (gm && gm_entry == 0) {
            ConvArgs a;
            a.B = 1; a.H = e0.H; a.W = e0.W; a.Cin = e0.cout; a.Cout = e0.cin; a.ntaps = e0.ks * e0.ks;
            a.in = gm; a.wg = exact ? e0.wd32 : e0.wd;
            a.ep.out = grad_image; a.ep.round = 0;
            if (e0.ks == 5 && exact) {
                if ((rc = conv_gen_fwd_launch(gm, 0, e0.w_flip, nullptr, grad_image, nullptr, 1, e0.H, e0.W, e0.cout, e0.cin, 5, 1, 2, 0, 0, st)))
                    return rc;
                p->launches_bwd++;
            } else if ((rc = run_conv(a))) return rc;
        } else {
            MAUA_CUDA_CHECK(cudaMemsetAsync(grad_image, 0, in_elems * sizeof(float), st));
        }
        return MAUA_OK;
    }

    // image-side tail
    ImageTail tail;
    const bool tv = p->img_io.tv_mode == MAUA_MODE_LOSS;
    const bool temporal = p->img_io.temporal_mode == MAUA_MODE_LOSS;
    if (tv || temporal) tail.img = p->image;
    if (tv) tail.tv_coef = p->coef2 + nt;
    if (temporal) {
        tail.temp_target = p->img_io.temporal_target;
        tail.temp_weights = p->img_io.temporal_weights;
        tail.temp_coef = p->coef2 + nt + 1;
    }
    Entry& e0 = p->entries[0];
    if (e0.ks != 3 && !exact) {
        // NIN conv1 (11x11 / 4) on the product path: T = Gm . W^T as a pointwise tcgen05 GEMM (128 -> 384 tap columns), then col2im
        // (<= 3 x 3 positions per pixel) fused with the image-side terms -- the scheme of conv1_1's dgrad (conv_edge.cu)
        const bool have = gm && gm_entry == 0;
        if (have) {
            ConvArgs a;
            a.B = 1; a.H = e0.H; a.W = e0.W; a.Cin = e0.cout; a.Cout = kIm2colK; a.ntaps = 1;
            a.in = gm; a.wg = e0.w1t;
            a.ep.out = p->im2col_ws; a.ep.round = 0;
            if ((rc = run_conv(a))) return rc;
        }
        if ((rc = col2im_img_launch(have ? p->im2col_ws : nullptr, grad_image, H, W, e0.ks, e0.stride, kIm2colK, tail, st))) return rc;
        p->launches_bwd++;
        prof_mark(p, st, "conv_first_dgrad", 0, 2.0 * 3 * e0.ks * e0.ks * e0.cout * e0.H * e0.W, 4.0 * ((e0.cout + 2.0 * kIm2colK) * e0.H * e0.W + 6.0 * H * W));
    } else if (e0.ks != 3) {
        // ... exact mode: gather-form input gradient + the image-side terms in one kernel (conv_gen.cu)
        if ((rc = conv_gen_dgrad_img_launch((gm && gm_entry == 0) ? gm : nullptr, e0.w_raw, grad_image, 1, H, W, e0.cout, e0.ks, e0.stride,
                                            tail, st)))
            return rc;
        p->launches_bwd++;
        prof_mark(p, st, "conv_first_dgrad", 0, 2.0 * 3 * e0.ks * e0.ks * e0.cout * e0.H * e0.W, 4.0 * (e0.cout * (double)e0.H * e0.W + 6.0 * H * W));
    } else if (gm && gm_entry == 0) {
        float* T = take_buf();
        if (T == gm) T = take_buf();
        if ((rc = conv_first_dgrad_launch(gm, exact ? e0.wt1_32 : e0.wt1, grad_image, 1, H, W, e0.cout, tail, T, p->impl, st)))
            return rc;
        p->launches_bwd += 2;
        prof_mark(p, st, "conv_first_dgrad", 0, 2.0 * 27 * e0.cout * H * W, 4.0 * (e0.cout + 6) * H * W);
    } else {
        // no feature-space loss is active: only TV / temporal terms (or nothing at all)
        MAUA_CUDA_CHECK(cudaMemsetAsync(p->gbuf[0], 0, (size_t)H * W * e0.cout * sizeof(float), st));
        if ((rc = conv_first_dgrad_launch(p->gbuf[0], exact ? e0.wt1_32 : e0.wt1, grad_image, 1, H, W, e0.cout, tail, p->gbuf[1],
                                          p->impl, st)))
            return rc;
        p->launches_bwd += 2;
    }
    return MAUA_OK;
}

extern "C" {

MAUA_API int maua_plan_backward(maua_plan_t* p, const float* grad_coefs, float* grad_image, maua_stream_t stream) {
    MAUA_REQUIRE(p && p->begin == 0 && p->last_stage,
                 "maua_plan_backward: this plan is one stage of a layer-wise split, use maua_plan_backward_stage");
    return plan_backward_impl(p, grad_coefs, nullptr, grad_image, stream);
}

MAUA_API int maua_plan_backward_stage(maua_plan_t* p, const float* grad_coefs, const float* grad_top, float* grad_out,
                                      maua_stream_t stream) {
    MAUA_REQUIRE(p, "maua_plan_backward_stage: null plan");
    return plan_backward_impl(p, grad_coefs, grad_top, grad_out, stream);
}

MAUA_API int maua_enable_peer_access(int a, int b) {
    if (a == b) return MAUA_OK;
    int prev = 0;
    MAUA_CUDA_CHECK(cudaGetDevice(&prev));
    for (int k = 0; k < 2; ++k) {
        const int from = k ? b : a, to = k ? a : b;
        int can = 0;
        MAUA_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, from, to));
        if (!can) {
            cudaSetDevice(prev);
            set_last_error("devices %d and %d are not peers (no NVLink / P2P path)", from, to);
            return MAUA_ERR_CUDA;
        }
        MAUA_CUDA_CHECK(cudaSetDevice(from));
        cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) {
            cudaSetDevice(prev);
            set_last_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", from, to, cudaGetErrorString(e));
            return MAUA_ERR_CUDA;
        }
    }
    MAUA_CUDA_CHECK(cudaSetDevice(prev));
    return MAUA_OK;
}

MAUA_API int maua_plan_stage_output_shape(const maua_plan_t* p, int h, int w, int* oh, int* ow, int* oc) {
    MAUA_REQUIRE(p && h >= 1 && w >= 1, "maua_plan_stage_output_shape: bad arguments");
    int c = p->in_C;
    for (const auto& e : p->entries) {
        entry_extent(e, h, w);
        c = e.C;
    }
    MAUA_REQUIRE(h >= 1 && w >= 1, "image too small for this stage");
    if (oh) *oh = h;
    if (ow) *ow = w;
    if (oc) *oc = c;
    return MAUA_OK;
}

MAUA_API int maua_plan_tap_gram(maua_plan_t* p, int tap, float* dst, int* c, maua_stream_t stream) {
    MAUA_REQUIRE(p && tap >= 0 && tap < (int)p->taps.size() && p->taps[tap].kind == MAUA_TAP_STYLE,
                 "maua_plan_tap_gram: bad tap index");
    const Tap& tp = p->taps[tap];
    if (c) *c = tp.C;
    if (dst) {
        MAUA_REQUIRE(tp.entry <= p->last_entry, "maua_plan_tap_gram: tap was not reached by the last forward");
        MAUA_CUDA_CHECK(cudaMemcpyAsync(dst, tp.gram, (size_t)tp.C * tp.C * sizeof(float), cudaMemcpyDeviceToDevice,
                                        (cudaStream_t)stream));
    }
    return MAUA_OK;
}

MAUA_API int maua_plan_entry_output(maua_plan_t* p, int entry, float* dst, int* h, int* w, int* c, int* is_pool,
                                    maua_stream_t stream) {
    MAUA_REQUIRE(p && entry >= 0 && entry < (int)p->entries.size(), "maua_plan_entry_output: bad entry index");
    const Entry& e = p->entries[entry];
    MAUA_REQUIRE(entry <= p->last_entry, "maua_plan_entry_output: entry was not reached by the last forward");
    if (h) *h = e.H;
    if (w) *w = e.W;
    if (c) *c = e.C;
    if (is_pool) *is_pool = e.pool ? 1 : 0;
    if (dst)
        MAUA_CUDA_CHECK(cudaMemcpyAsync(dst, e.out, (size_t)e.H * e.W * e.C * sizeof(float), cudaMemcpyDeviceToDevice,
                                        (cudaStream_t)stream));
    return MAUA_OK;
}

MAUA_API int maua_plan_set_tap_fold(maua_plan_t* p, int tap, const float* in2, int k2, const float* w2, const float* bias) {
    MAUA_REQUIRE(p && tap >= 0 && tap < (int)p->taps.size() && p->taps[tap].kind == MAUA_TAP_STYLE && p->taps[tap].entry >= 0,
                 "maua_plan_set_tap_fold: bad tap index");
    MAUA_REQUIRE(in2 && w2 && k2 > 0 && k2 % 32 == 0, "maua_plan_set_tap_fold: need in2, w2 and k2 %% 32 == 0 (got %d)", k2);
    Tap& tp = p->taps[tap];
    MAUA_REQUIRE(tp.mode == MAUA_MODE_EXTERNAL && tp.active, "maua_plan_set_tap_fold: tap %d was not in external mode in the last forward", tap);
    tp.ext_in2 = in2; tp.ext_K2 = k2; tp.ext_w2 = w2; tp.ext_bias = bias;
    return MAUA_OK;
}

MAUA_API int maua_plan_tap_feature_strided(maua_plan_t* p, int tap, float* dst, long dst_pixel_stride, maua_stream_t stream) {
    MAUA_REQUIRE(p && dst && tap >= 0 && tap < (int)p->taps.size() && p->taps[tap].entry >= 0,
                 "maua_plan_tap_feature_strided: bad arguments");
    const Entry& e = p->entries[p->taps[tap].entry];
    MAUA_REQUIRE(p->taps[tap].entry <= p->last_entry, "maua_plan_tap_feature_strided: tap was not reached by the last forward");
    MAUA_REQUIRE(dst_pixel_stride >= e.C, "maua_plan_tap_feature_strided: stride %ld < %d channels", dst_pixel_stride, e.C);
    DeviceGuard guard(p->device);
    MAUA_CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)dst_pixel_stride * sizeof(float), e.out, (size_t)e.C * sizeof(float),
                                      (size_t)e.C * sizeof(float), (size_t)e.H * e.W, cudaMemcpyDeviceToDevice,
                                      (cudaStream_t)stream));
    return MAUA_OK;
}

MAUA_API int maua_plan_tap_feature(maua_plan_t* p, int tap, float* dst, int* h, int* w, int* c, maua_stream_t stream) {
    MAUA_REQUIRE(p && tap >= 0 && tap < (int)p->taps.size(), "maua_plan_tap_feature: bad tap index");
    const Entry& e = p->entries[p->taps[tap].entry];
    MAUA_REQUIRE(p->taps[tap].entry <= p->last_entry, "maua_plan_tap_feature: tap was not reached by the last forward");
    if (h) *h = e.H;
    if (w) *w = e.W;
    if (c) *c = e.C;
    if (dst)
        MAUA_CUDA_CHECK(cudaMemcpyAsync(dst, e.out, (size_t)e.H * e.W * e.C * sizeof(float), cudaMemcpyDeviceToDevice,
                                        (cudaStream_t)stream));
    return MAUA_OK;
}

}  // extern "C"
