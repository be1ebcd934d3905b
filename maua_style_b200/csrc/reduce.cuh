// Deterministic grid-wide sums without float atomics.
//
// Every block reduces its values with warp shuffles, writes one partial per value, and takes a ticket; the block
// that draws the last ticket re-reads all partials (coalesced, all threads) and reduces them in a fixed order, so the
// result does not depend on block scheduling.  The counter re-arms itself for the next launch on the stream.
#pragma once
#include "common.cuh"

namespace maua {

constexpr int kReduceThreads = 256;

// Call from all threads of every block (blockDim.x == kReduceThreads).  Returns true in thread 0 of the last block
// only, with tot[] holding the grid totals.
template <int NV, int NT = kReduceThreads>
__device__ __forceinline__ bool grid_sum(const double (&v)[NV], double* __restrict__ partials, unsigned int* counter,
                                         double (&tot)[NV]) {
    constexpr int kReduceThreads = NT;  // blockDim.x of the calling kernel
    __shared__ double sh[NV][kReduceThreads / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const double s = warp_sum(v[k]);
        if (lane == 0) sh[k][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = 0;
#pragma unroll
            for (int i = 0; i < kReduceThreads / 32; ++i) s += sh[k][i];
            partials[(size_t)blockIdx.x * NV + k] = s;
        }
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += kReduceThreads)
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[k] += __ldcg(partials + (size_t)b * NV + k);
    __syncthreads();  // sh[] reuse
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const double s = warp_sum(acc[k]);
        if (lane == 0) sh[k][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x != 0) return false;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = 0;
#pragma unroll
        for (int i = 0; i < kReduceThreads / 32; ++i) s += sh[k][i];
        tot[k] = s;
    }
    *counter = 0;
    return true;
}

}  // namespace maua
