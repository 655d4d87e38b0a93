// Small device-side utilities shared by the builders: exclusive scan, launch helpers, timers.
// (Included by several translation units; every kernel here is `static` to its TU.)
#pragma once
#include <chrono>
#include <cstdint>
#include <cstring>

#include "sdf_internal.h"

namespace sdfb200 {

inline uint32_t divUp(uint64_t a, uint64_t b) { return uint32_t((a + b - 1) / b); }

inline double msSince(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

// ---- small device -> host readbacks ----------------------------------------------------------------------------
// Scan totals and counters come back through a PINNED slot owned by the calling thread: a copy into pageable memory
// (a stack variable) is staged by the driver under a process-wide lock, which serialises the host threads of a
// single-process multi-device build against each other (measured on 8 x B200: profiles/r2_summary.md).
inline void* pinnedScratch() {   // 256 bytes per host thread, allocated on first use, freed at process exit by the driver
    static thread_local void* slot = nullptr;
    if (!slot && cudaHostAlloc(&slot, 256, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); slot = nullptr; }
    return slot;
}
template <class T> inline T readScalar(const T* dSrc, cudaStream_t st = 0) {
    static_assert(sizeof(T) <= 256, "readScalar payload");
    T out;
    if (void* pin = pinnedScratch()) {
        SDFB_CUDA(cudaMemcpyAsync(pin, dSrc, sizeof(T), cudaMemcpyDeviceToHost, st));
        SDFB_CUDA(cudaStreamSynchronize(st));
        std::memcpy(&out, pin, sizeof(T));
    } else {
        SDFB_CUDA(cudaMemcpyAsync(&out, dSrc, sizeof(T), cudaMemcpyDeviceToHost, st));
        SDFB_CUDA(cudaStreamSynchronize(st));
    }
    return out;
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
        return readScalar<Out>(total.p, st);
    }
};
using Scanner = ScannerT<uint32_t, uint32_t>;

// ---- exclusive scan of 0/1 byte flags into 32-bit positions, 16 flags per thread ---------------------------
// The generic scan above moves one element per thread; the flag passes of the ExactOctreeSdf builder run over up to
// billions of (node, triangle) pairs, so this variant reads the flags as 128-bit words (popcount per word) and writes
// the positions as four 128-bit stores per thread.
constexpr int kFlagThreads = 256, kFlagsPerThread = 16, kFlagsPerBlock = kFlagThreads * kFlagsPerThread;

static __device__ __forceinline__ uint4 loadFlags16(const uint8_t* in, uint64_t i, uint64_t n) {
    if (i + 16 <= n) return *reinterpret_cast<const uint4*>(in + i);
    uint32_t w[4] = {0, 0, 0, 0};
    for (int k = 0; k < 16; k++)
        if (i + k < n) w[k >> 2] |= uint32_t(in[i + k]) << (8 * (k & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

static __global__ void __launch_bounds__(kFlagThreads) flagBlockSums(const uint8_t* in, uint32_t* blockSums, uint64_t n) {
    __shared__ uint32_t warpSums[32];
    const uint64_t i = (uint64_t(blockIdx.x) * kFlagThreads + threadIdx.x) * kFlagsPerThread;
    uint32_t v = 0;
    if (i < n) { const uint4 f = loadFlags16(in, i, n); v = __popc(f.x) + __popc(f.y) + __popc(f.z) + __popc(f.w); }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t s = threadIdx.x < kFlagThreads / 32 ? warpSums[threadIdx.x] : 0u;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) blockSums[blockIdx.x] = s;
    }
}

// kGroups: out[i / 16] = position of the first flag of every group of 16 only (flagPositionAt recovers any position from it and
// the flags: the pair passes of the ExactOctreeSdf builder run over billions of pairs, and 4 bytes per pair were half of its memory)
template <bool kGroups>
static __global__ void __launch_bounds__(kFlagThreads) flagFinalize(const uint8_t* in, const uint32_t* blockSums, uint32_t* out, uint64_t n) {
    __shared__ uint32_t warpTotals[32];
    const uint64_t i = (uint64_t(blockIdx.x) * kFlagThreads + threadIdx.x) * kFlagsPerThread;
    uint4 f = make_uint4(0, 0, 0, 0);
    if (i < n) f = loadFlags16(in, i, n);
    const uint32_t mine = __popc(f.x) + __popc(f.y) + __popc(f.z) + __popc(f.w);
    // in-block inclusive scan of the per-thread sums (8 warps)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warpTotals[warp] = inc;
    __syncthreads();
    uint32_t base = blockSums[blockIdx.x] + inc - mine;
    for (int w = 0; w < warp; w++) base += warpTotals[w];
    if (i >= n) return;
    if (kGroups) { out[i >> 4] = base; return; }
    const uint32_t words[4] = {f.x, f.y, f.z, f.w};
    uint32_t pos[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        pos[k] = base;
        base += (words[k >> 2] >> (8 * (k & 3))) & 1u;
    }
    if (i + 16 <= n) {
#pragma unroll
        for (int q = 0; q < 4; q++) reinterpret_cast<uint4*>(out + i)[q] = make_uint4(pos[4 * q], pos[4 * q + 1], pos[4 * q + 2], pos[4 * q + 3]);
    } else {
        for (int k = 0; k < 16; k++) if (i + k < n) out[i + k] = pos[k];
    }
}

// Flags must be exactly 0 or 1. Returns the number of set flags (32-bit: callers with n >= 2^32 check countFlags64 first).
struct FlagScanner {
    DevBuf<uint32_t> blockSums, total;
    uint32_t run(const uint8_t* in, uint32_t* out, uint64_t n, cudaStream_t st = 0) {
        if (!total.p) total.alloc(1);
        if (n == 0) return 0;
        const uint32_t nBlocks = divUp(n, kFlagsPerBlock);
        if (blockSums.n < nBlocks) blockSums.alloc(size_t(nBlocks) + 64);
        flagBlockSums<<<nBlocks, kFlagThreads, 0, st>>>(in, blockSums.p, n);
        scanOfBlockSums<uint32_t><<<1, kScanBlock, 0, st>>>(blockSums.p, nBlocks, total.p);
        flagFinalize<false><<<nBlocks, kFlagThreads, 0, st>>>(in, blockSums.p, out, n);
        return readScalar<uint32_t>(total.p, st);
    }
    // out: (n + 15) / 16 entries, the position of the first flag of each group of 16
    uint32_t runGroups(const uint8_t* in, uint32_t* outGroups, uint64_t n, cudaStream_t st = 0) {
        if (!total.p) total.alloc(1);
        if (n == 0) return 0;
        const uint32_t nBlocks = divUp(n, kFlagsPerBlock);
        if (blockSums.n < nBlocks) blockSums.alloc(size_t(nBlocks) + 64);
        flagBlockSums<<<nBlocks, kFlagThreads, 0, st>>>(in, blockSums.p, n);
        scanOfBlockSums<uint32_t><<<1, kScanBlock, 0, st>>>(blockSums.p, nBlocks, total.p);
        flagFinalize<true><<<nBlocks, kFlagThreads, 0, st>>>(in, blockSums.p, outGroups, n);
        return readScalar<uint32_t>(total.p, st);
    }
};

// position (exclusive count of set flags before i) from the group positions of FlagScanner::runGroups; i < n
static __device__ __forceinline__ uint32_t flagPositionAt(const uint8_t* flags, const uint32_t* groupPos, uint64_t i, uint64_t n) {
    const uint64_t g = i & ~uint64_t(15);
    const uint32_t k = uint32_t(i - g);
    uint32_t at = groupPos[g >> 4];
    if (k == 0) return at;
    const uint4 f = loadFlags16(flags, g, n);
    const uint32_t w[4] = {f.x, f.y, f.z, f.w};
    const uint32_t whole = k >> 2, rest = k & 3u;                          // whole words, then the low bytes of the next one
    for (uint32_t q = 0; q < whole; q++) at += uint32_t(__popc(w[q]));
    if (rest) at += uint32_t(__popc(w[whole] & ((1u << (8 * rest)) - 1u)));
    return at;
}

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
    return readScalar<unsigned long long>(total.p);
}

}  // namespace sdfb200
