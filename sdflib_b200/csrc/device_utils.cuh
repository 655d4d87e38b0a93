// Small device-side utilities shared by the builders: exclusive scan, launch helpers, timers.
// (Included by several translation units; every kernel here is `static` to its TU.)
#pragma once
#include <chrono>
#include <cstdint>

#include "sdf_internal.h"

namespace sdfb200 {

inline uint32_t divUp(uint64_t a, uint64_t b) { return uint32_t((a + b - 1) / b); }

inline double msSince(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// ---- exclusive scan of uint32 values (three small kernels; sums must fit in 32 bits) -----------------
constexpr int kScanBlock = 1024;

static __global__ void scanBlockSums(const uint32_t* in, uint32_t* blockSums, uint32_t n) {
    __shared__ uint32_t warpSums[32];
    const uint32_t i = blockIdx.x * kScanBlock + threadIdx.x;
    uint32_t v = i < n ? in[i] : 0u;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t s = warpSums[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) blockSums[blockIdx.x] = s;
    }
}

// in-block inclusive scan of one value per thread (warp shuffles + one smem pass)
static __device__ __forceinline__ uint32_t blockInclusiveScan(uint32_t v, uint32_t* warpTotals /*[32] shared*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) warpTotals[warp] = v;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warpTotals[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warpTotals[lane] = w;
    }
    __syncthreads();
    if (warp > 0) v += warpTotals[warp - 1];
    __syncthreads();
    return v;
}

static __global__ void scanOfBlockSums(uint32_t* blockSums, uint32_t nBlocks, uint32_t* total) {
    __shared__ uint32_t warpTotals[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t start = 0; start < nBlocks; start += kScanBlock) {   // single CTA, sequential over chunks
        const uint32_t i = start + threadIdx.x;
        const uint32_t v = i < nBlocks ? blockSums[i] : 0u;
        const uint32_t inc = blockInclusiveScan(v, warpTotals);
        const uint32_t base = carry;
        if (i < nBlocks) blockSums[i] = base + inc - v;   // exclusive
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) carry = base + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

static __global__ void scanFinalize(const uint32_t* in, const uint32_t* blockSums, uint32_t* out, uint32_t n) {
    __shared__ uint32_t warpTotals[32];
    const uint32_t i = blockIdx.x * kScanBlock + threadIdx.x;
    const uint32_t v = i < n ? in[i] : 0u;
    const uint32_t inc = blockInclusiveScan(v, warpTotals);
    if (i < n) out[i] = blockSums[blockIdx.x] + inc - v;
}

// Exclusive scan helper: out[i] = sum_{j<i} in[j]; returns the total (one 4-byte D2H copy = one sync).
// `out` may alias `in`.
struct Scanner {
    DevBuf<uint32_t> blockSums, total;
    uint64_t launches = 0;
    uint32_t run(const uint32_t* in, uint32_t* out, uint32_t n, cudaStream_t st = 0) {
        if (n == 0) return 0;
        const uint32_t nBlocks = divUp(n, kScanBlock);
        if (blockSums.n < nBlocks) blockSums.alloc(nBlocks + 64);
        if (!total.p) total.alloc(1);
        scanBlockSums<<<nBlocks, kScanBlock, 0, st>>>(in, blockSums.p, n);
        scanOfBlockSums<<<1, kScanBlock, 0, st>>>(blockSums.p, nBlocks, total.p);
        scanFinalize<<<nBlocks, kScanBlock, 0, st>>>(in, blockSums.p, out, n);
        launches += 3;
        uint32_t t = 0;
        SDFB_CUDA(cudaMemcpyAsync(&t, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SDFB_CUDA(cudaStreamSynchronize(st));
        return t;
    }
};

}  // namespace sdfb200
