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

// ---- exclusive scan (three small kernels) ---------------------------------------------------------------
// In = uint8_t / uint32_t values, Out = uint32_t / uint64_t running sums (choose Out wide enough for the total).
constexpr int kScanBlock = 1024;

template <class Out> static __device__ __forceinline__ Out shflUp(Out v, int o) { return __shfl_up_sync(0xffffffffu, v, o); }
template <class Out> static __device__ __forceinline__ Out shflDown(Out v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }

// in-block inclusive scan of one value per thread (warp shuffles + one smem pass)
template <class Out>
static __device__ __forceinline__ Out blockInclusiveScan(Out v, Out* warpTotals /*[32] shared*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const Out t = shflUp<Out>(v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) warpTotals[warp] = v;
    __syncthreads();
    if (warp == 0) {
        Out w = warpTotals[lane];
        for (int o = 1; o < 32; o <<= 1) {
            const Out t = shflUp<Out>(w, o);
            if (lane >= o) w += t;
        }
        warpTotals[lane] = w;
    }
    __syncthreads();
    if (warp > 0) v += warpTotals[warp - 1];
    __syncthreads();
    return v;
}

template <class In, class Out>
static __global__ void scanBlockSums(const In* in, Out* blockSums, uint64_t n) {
    __shared__ Out warpSums[32];
    const uint64_t i = uint64_t(blockIdx.x) * kScanBlock + threadIdx.x;
    Out v = i < n ? Out(in[i]) : Out(0);
    for (int o = 16; o > 0; o >>= 1) v += shflDown<Out>(v, o);
    if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        Out s = warpSums[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) s += shflDown<Out>(s, o);
        if (threadIdx.x == 0) blockSums[blockIdx.x] = s;
    }
}

template <class Out>
static __global__ void scanOfBlockSums(Out* blockSums, uint32_t nBlocks, Out* total) {
    __shared__ Out warpTotals[32];
    __shared__ Out carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t start = 0; start < nBlocks; start += kScanBlock) {   // single CTA, sequential over chunks
        const uint32_t i = start + threadIdx.x;
        const Out v = i < nBlocks ? blockSums[i] : Out(0);
        const Out inc = blockInclusiveScan<Out>(v, warpTotals);
        const Out base = carry;
        if (i < nBlocks) blockSums[i] = base + inc - v;   // exclusive
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) carry = base + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// out[i] = sum_{j<i} in[j] for i < n, and out[n] = total when writeTotal (out then has n + 1 entries)
template <class In, class Out>
static __global__ void scanFinalize(const In* in, const Out* blockSums, Out* out, uint64_t n, const Out* total, bool writeTotal) {
    __shared__ Out warpTotals[32];
    const uint64_t i = uint64_t(blockIdx.x) * kScanBlock + threadIdx.x;
    const Out v = i < n ? Out(in[i]) : Out(0);
    const Out inc = blockInclusiveScan<Out>(v, warpTotals);
    if (i < n) out[i] = blockSums[blockIdx.x] + inc - v;
    if (writeTotal && i == n) out[n] = *total;
}

// Exclusive scan helper; returns the total (one small D2H copy = one sync). `out` may alias `in` when the
// element types match.
template <class In, class Out> struct ScannerT {
    DevBuf<Out> blockSums, total;
    uint64_t launches = 0;
    Out run(const In* in, Out* out, uint64_t n, bool writeTotal = false, cudaStream_t st = 0) {
        if (!total.p) total.alloc(1);
        if (n == 0) {
            if (writeTotal) SDFB_CUDA(cudaMemsetAsync(out, 0, sizeof(Out), st));
            return 0;
        }
        const uint32_t nBlocks = divUp(n + (writeTotal ? 1 : 0), kScanBlock);
        if (blockSums.n < nBlocks) blockSums.alloc(nBlocks + 64);
        scanBlockSums<In, Out><<<nBlocks, kScanBlock, 0, st>>>(in, blockSums.p, n);
        scanOfBlockSums<Out><<<1, kScanBlock, 0, st>>>(blockSums.p, nBlocks, total.p);
        scanFinalize<In, Out><<<nBlocks, kScanBlock, 0, st>>>(in, blockSums.p, out, n, total.p, writeTotal);
        launches += 3;
        Out t = 0;
        SDFB_CUDA(cudaMemcpyAsync(&t, total.p, sizeof(Out), cudaMemcpyDeviceToHost, st));
        SDFB_CUDA(cudaStreamSynchronize(st));
        return t;
    }
};
using Scanner = ScannerT<uint32_t, uint32_t>;

// 64-bit count of set flags: guards the 32-bit flag scans (their totals would wrap silently) once a pass has
// more than 2^32 pairs.
static __global__ void countFlagsKernel(const uint8_t* flags, uint64_t n, unsigned long long* total) {
    unsigned long long c = 0;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) c += flags[i];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, c);
}
inline uint64_t countFlags64(const uint8_t* flags, uint64_t n) {
    DevBuf<unsigned long long> total(1);
    SDFB_CUDA(cudaMemsetAsync(total.p, 0, 8));
    countFlagsKernel<<<148 * 8, 256>>>(flags, n, total.p);
    unsigned long long t = 0;
    SDFB_CUDA(cudaMemcpy(&t, total.p, 8, cudaMemcpyDeviceToHost));
    return t;
}

}  // namespace sdfb200
