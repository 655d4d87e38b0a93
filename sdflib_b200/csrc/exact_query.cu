// Bulk ExactOctreeSdf::getDistance / getDistance(p, grad) (hot path 2, exact variant).
//
// Reference: src/sdf/ExactOctreeSdf.cpp:38-178 and :180-320 — descend to the node at
// bitEncodingStartDepth, unpack its triangle set, filter it through the chain of per-child bit masks
// down to the leaf, brute-force the nearest triangle (first strict minimum, ascending order), then
// TriangleUtils::getSignedDistPointAndTriangle (include/SdfLib/utils/TriangleUtils.h:137-196, :292-376).
//
// B200 design: the per-query mask-chain decode of the reference is the same work for every query that
// lands in a leaf, so it is done ONCE per leaf: prepareExactQuery() decodes the public arrays
// (mOctreeData / mTrianglesSets / mTrianglesMasks — freshly built or loaded from a .bin) level by level
// into a private pool of explicit per-leaf triangle lists with flat (node, entry) pair passes and scans.
// The public arrays stay bit-identical to the reference; the query kernel walks the node array per lane, then
// the warp streams each query's leaf list cooperatively (frames fetched as 5 x 128-bit read-only loads).
// Compiled with -fmad=false: distances are bit-identical to the CPU reference.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "device_utils.cuh"
#include "sdf_internal.h"

namespace sdfb200 {

namespace {

constexpr uint32_t kNone = 0xFFFFFFFFu;
enum : uint32_t { kErrSetRange = 1, kErrMaskRange = 2, kErrTriangle = 4 };

__device__ __forceinline__ uint32_t unpackIndex(const uint32_t* words, uint64_t bitIdx, uint32_t bits) {   // ExactOctreeSdf.cpp:73-77
    const uint64_t w = bitIdx >> 5;
    const uint32_t bit = uint32_t(bitIdx & 31u);
    return ((words[w] << bit) >> (32 - bits)) | uint32_t(uint64_t(words[w + 1]) >> (64 - (bit + bits)));
}

template <class T> __device__ __forceinline__ uint32_t lastLessEqual(const T* a, uint32_t n, T key) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] <= key) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void framesFromTriData(const TriData* tris, float4* frames, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 5u) return;
    const uint32_t t = i / 5u, q = i % 5u;
    const float* src = reinterpret_cast<const float*>(tris + t) + 4 * q;
    frames[i] = make_float4(src[0], src[1], src[2], q == 4 ? 0.0f : src[3]);
}

struct PublicView {
    const uint32_t* nodes;   // (childrenIndex | leaf, trianglesArrayIndex) pairs
    const uint32_t* sets;
    const uint8_t* masks;
    uint64_t numNodes, numSets, numMasks;
    uint32_t numTriangles, bits, bitEnc;
};

// decoded-list length of every frontier node at depth <= bitEnc, and whether it has children
__global__ void frontierCountsKernel(PublicView pv, const uint32_t* nodeIdx, uint32_t n, uint32_t depth, uint32_t* cnt,
                                     uint32_t* inner, uint32_t* leafCnt, uint32_t* err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t w0 = pv.nodes[2 * size_t(nodeIdx[i])], w1 = pv.nodes[2 * size_t(nodeIdx[i]) + 1];
    const bool leaf = (w0 & kLeafBit) != 0;
    uint32_t c = 0;
    if (leaf || depth == pv.bitEnc) {
        if (uint64_t(w1) + 2 > pv.numSets) atomicOr(err, kErrSetRange);
        else {
            c = pv.sets[w1];
            if (uint64_t(w1) + 2 + (uint64_t(c) * pv.bits + 31) / 32 > pv.numSets) { atomicOr(err, kErrSetRange); c = 0; }
        }
    }
    cnt[i] = c;
    inner[i] = leaf ? 0u : 1u;
    leafCnt[i] = leaf ? c : 0u;
}

// unpack the packed set of every frontier node that owns one (CTA per node)
__global__ void __launch_bounds__(128)
unpackSetsKernel(PublicView pv, const uint32_t* nodeIdx, const uint64_t* lo, const uint32_t* cnt, uint32_t* dec, uint32_t* err) {
    const uint32_t i = blockIdx.x;
    const uint32_t c = cnt[i];
    if (c == 0) return;
    const uint32_t* words = pv.sets + pv.nodes[2 * size_t(nodeIdx[i]) + 1] + 1;
    uint32_t* dst = dec + lo[i];
    for (uint32_t t = threadIdx.x; t < c; t += blockDim.x) {
        uint32_t tri = unpackIndex(words, uint64_t(t) * pv.bits, pv.bits);
        if (tri >= pv.numTriangles) { atomicOr(err, kErrTriangle); tri = 0; }
        dst[t] = tri;
    }
}

// frontier below bitEnc: a node's list = its parent's decoded list filtered by the node's bit mask (MSB first,
// ExactOctreeSdf.cpp:105-131). One warp per node, twice: count the set bits, then (after the scan of the counts)
// copy the selected parent entries in order. No per-pair arrays: on the Dragon-class structure a level has more than
// 2^32 (node, parent entry) pairs.
__global__ void __launch_bounds__(256)
maskCountKernel(PublicView pv, const uint32_t* __restrict__ nodeIdx, const uint32_t* __restrict__ parentCnt, uint32_t n, uint32_t* cnt,
                uint32_t* inner, uint32_t* leafCnt, uint32_t* err) {
    const uint32_t i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint32_t w0 = pv.nodes[2 * size_t(nodeIdx[i])], w1 = pv.nodes[2 * size_t(nodeIdx[i]) + 1];
    const uint32_t pc = parentCnt[i], bytes = (pc + 7u) / 8u;
    uint32_t c = 0;
    if (uint64_t(w1) + bytes > pv.numMasks) { if (lane == 0) atomicOr(err, kErrMaskRange); }
    else
        for (uint32_t b = lane; b < bytes; b += 32) {
            uint32_t byte = pv.masks[uint64_t(w1) + b];
            if (b == bytes - 1 && (pc & 7u)) byte &= 0xFF00u >> (pc & 7u);   // bits beyond the parent's count are padding
            c += __popc(byte);
        }
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) {
        const bool leaf = (w0 & kLeafBit) != 0;
        cnt[i] = c;
        inner[i] = leaf ? 0u : 1u;
        leafCnt[i] = leaf ? c : 0u;
    }
}

__global__ void __launch_bounds__(256)
maskSelectKernel(PublicView pv, const uint32_t* __restrict__ nodeIdx, const uint64_t* __restrict__ parentLo, const uint32_t* __restrict__ parentCnt,
                 const uint64_t* __restrict__ lo, uint32_t n, const uint32_t* __restrict__ parentDec, uint32_t* __restrict__ dec) {
    const uint32_t i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint32_t w1 = pv.nodes[2 * size_t(nodeIdx[i]) + 1];
    const uint32_t pc = parentCnt[i];
    if (uint64_t(w1) + (pc + 7u) / 8u > pv.numMasks) return;   // reported by maskCountKernel
    const uint32_t* src = parentDec + parentLo[i];
    uint32_t* dst = dec + lo[i];
    uint32_t written = 0;
    for (uint32_t j0 = 0; j0 < pc; j0 += 32) {
        const uint32_t j = j0 + lane;
        const bool keep = j < pc && (pv.masks[uint64_t(w1) + (j >> 3)] & (0x80u >> (j & 7u)));
        const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
        if (keep) dst[written + __popc(ballot & ((1u << lane) - 1u))] = src[j];
        written += __popc(ballot);
    }
}

__global__ void nextFrontierKernel(PublicView pv, const uint32_t* nodeIdx, const uint32_t* inner, const uint32_t* childSlot,
                                   const uint64_t* lo, const uint32_t* cnt, uint32_t n, uint32_t* nNodeIdx, uint64_t* nParentLo,
                                   uint32_t* nParentCnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !inner[i]) return;
    const uint32_t c0 = pv.nodes[2 * size_t(nodeIdx[i])] & kExactIndexMask;
    for (uint32_t c = 0; c < 8; c++) {
        const uint32_t k = childSlot[i] * 8u + c;
        nNodeIdx[k] = c0 + c;
        nParentLo[k] = lo[i];
        nParentCnt[k] = cnt[i];
    }
}

// leaves of a frontier -> private pool (CTA per frontier node)
__global__ void __launch_bounds__(128)
leafPoolKernel(const uint32_t* nodeIdx, const uint32_t* inner, const uint64_t* lo, const uint32_t* cnt, const uint64_t* poolOff,
               const uint32_t* dec, uint32_t* pool, uint64_t* leafLo, uint32_t* leafCnt) {
    const uint32_t i = blockIdx.x;
    if (inner[i]) return;
    const uint32_t c = cnt[i];
    const uint64_t dst = poolOff[i];
    if (threadIdx.x == 0) { leafLo[nodeIdx[i]] = dst; leafCnt[nodeIdx[i]] = c; }
    for (uint32_t t = threadIdx.x; t < c; t += blockDim.x) pool[dst + t] = dec[lo[i] + t];
}

__global__ void rebaseLeavesKernel(const uint32_t* nodeIdx, const uint32_t* inner, uint32_t n, uint64_t base, uint64_t* leafLo) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !inner[i]) leafLo[nodeIdx[i]] += base;
}

// ---- the query kernel -----------------------------------------------------------------------------------------
struct ExactQueryParams {
    float minx, miny, minz, maxx, maxy, maxz;
    float cell;
    int grid;
    float outside;   // sqrt(3) * boxSize.x
};

__device__ __forceinline__ TriFrame loadFrame(const float4* __restrict__ frames, uint32_t t) {
    const float4 a = __ldg(frames + 5 * size_t(t)), b = __ldg(frames + 5 * size_t(t) + 1), c = __ldg(frames + 5 * size_t(t) + 2),
                 d = __ldg(frames + 5 * size_t(t) + 3), e = __ldg(frames + 5 * size_t(t) + 4);
    TriFrame f;
    f.ox = a.x; f.oy = a.y; f.oz = a.z; f.t00 = a.w;
    f.t01 = b.x; f.t02 = b.y; f.t10 = b.z; f.t11 = b.w;
    f.t12 = c.x; f.t20 = c.y; f.t21 = c.z; f.t22 = c.w;
    f.bx = d.x; f.by = d.y; f.cx = d.z; f.cy = d.w;
    f.v2 = e.x; f.v3x = e.y; f.v3y = e.z;
    return f;
}

__device__ __forceinline__ float boxDistance(const ExactQueryParams& q, f3 p) {   // Mesh.h:42-46
    const f3 size = mk3(q.maxx - q.minx, q.maxy - q.miny, q.maxz - q.minz);
    const f3 center = mk3(q.minx, q.miny, q.minz) + 0.5f * size;
    const f3 d = p - center;
    const f3 h = 0.5f * size;
    const f3 a = mk3(gabs(d.x) - h.x, gabs(d.y) - h.y, gabs(d.z) - h.z);
    const f3 ap = mk3(gmax(a.x, 0.0f), gmax(a.y, 0.0f), gmax(a.z, 0.0f));
    return sqrtf(dot3(ap, ap)) + gmin(gmax(a.x, gmax(a.y, a.z)), 0.0f);
}

// Warp-cooperative: each lane walks the node array for its own query, then the warp works through its 32
// queries one leaf at a time with the LANES STRIDING OVER THE TRIANGLE LIST (two queries per pass when the next
// one landed in the same leaf, so a frame is fetched once for both). Every lane runs the same trip count whatever
// the list lengths of the other queries are. The first version of this kernel looped over the list per thread and
// ran with 10 of 32 lanes active on the 256^3 grid (the 32 x-consecutive queries of a warp fall into 16 different
// leaves): C3, 16.7 M queries, 9.97 ms grid / 19.5 ms random points against 8.6 / 14.0 ms here. The serial rule
// "first strict minimum over ascending list positions" is kept exactly: per-lane strict '<' over ascending
// positions, then redux.min over the distance bits (non-negative floats order like their bit patterns) and
// redux.min over the positions of the lanes that hold that minimum.
template <bool kGrad>
__global__ void __launch_bounds__(256)
exactQueryWarpKernel(const uint32_t* __restrict__ nodes, const uint64_t* __restrict__ leafLo, const uint32_t* __restrict__ leafCnt,
                     const uint32_t* __restrict__ pool, const float4* __restrict__ frames, const TriData* __restrict__ tris,
                     const ExactQueryParams q, const float* __restrict__ xyz, uint64_t n, float* __restrict__ dist,
                     float* __restrict__ grad, const uint32_t* __restrict__ runIfZero) {
    constexpr unsigned kFull = 0xffffffffu;
    if (runIfZero && *runIfZero != 0u) return;   // large batches: the binned kernel was chosen on the device
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = i < n;
    f3 p = mk3(0.0f, 0.0f, 0.0f);
    if (valid) p = mk3(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2));
    float fx = (p.x - q.minx) / q.cell, fy = (p.y - q.miny) / q.cell, fz = (p.z - q.minz) / q.cell;
    const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    const int ix = int(flx), iy = int(fly), iz = int(flz);
    fx -= flx; fy -= fly; fz -= flz;
    const bool inside = valid && !(ix < 0 || ix >= q.grid || iy < 0 || iy >= q.grid || iz < 0 || iz >= q.grid);
    unsigned long long lo = 0;
    uint32_t cnt = 0;
    if (inside) {
        uint32_t idx = uint32_t((iz * q.grid + iy) * q.grid + ix);
        uint32_t w0 = __ldg(nodes + 2 * size_t(idx));
        while (!(w0 & kLeafBit)) {   // roundFloat of ExactOctreeSdf.cpp:33-36 is a strict '>'
            const uint32_t child = ((fz > 0.5f) ? 4u : 0u) + ((fy > 0.5f) ? 2u : 0u) + ((fx > 0.5f) ? 1u : 0u);
            idx = (w0 & kExactIndexMask) + child;
            w0 = __ldg(nodes + 2 * size_t(idx));
            fx = 2.0f * fx; fy = 2.0f * fy; fz = 2.0f * fz;
            fx -= floorf(fx); fy -= floorf(fy); fz -= floorf(fz);
        }
        lo = leafLo[idx];
        cnt = leafCnt[idx];
    }
    uint32_t myTri = 0;   // the serial loop's initial bestTri
    unsigned todo = __ballot_sync(kFull, inside && cnt > 0);
    while (todo) {
        const int j = __ffs(todo) - 1;
        todo &= todo - 1;
        const unsigned long long loJ = __shfl_sync(kFull, lo, j);
        const uint32_t cntJ = __shfl_sync(kFull, cnt, j);
        const f3 p1 = mk3(__shfl_sync(kFull, p.x, j), __shfl_sync(kFull, p.y, j), __shfl_sync(kFull, p.z, j));
        int j2 = todo ? __ffs(todo) - 1 : 0;
        const bool pair = todo && __shfl_sync(kFull, lo, j2) == loJ;   // lists of non-empty leaves start at distinct offsets
        const f3 p2 = mk3(__shfl_sync(kFull, p.x, j2), __shfl_sync(kFull, p.y, j2), __shfl_sync(kFull, p.z, j2));
        if (pair) todo &= todo - 1;
        const uint32_t* lst = pool + loJ;
        float best1 = INFINITY, best2 = INFINITY;
        uint32_t k1 = 0xFFFFFFFFu, k2 = 0xFFFFFFFFu;
        for (uint32_t k = lane; k < cntJ; k += 32) {
            const TriFrame f = loadFrame(frames, __ldg(lst + k));
            const float sq1 = sqDistPointTriangle(p1, f);
            if (sq1 < best1) { best1 = sq1; k1 = k; }
            if (pair) {
                const float sq2 = sqDistPointTriangle(p2, f);
                if (sq2 < best2) { best2 = sq2; k2 = k; }
            }
        }
        const uint32_t m1 = __reduce_min_sync(kFull, __float_as_uint(best1));
        const uint32_t w1 = __reduce_min_sync(kFull, (__float_as_uint(best1) == m1) ? k1 : 0xFFFFFFFFu);
        const uint32_t t1 = (w1 != 0xFFFFFFFFu) ? __ldg(lst + w1) : 0u;
        if (lane == j) myTri = t1;
        if (pair) {
            const uint32_t m2 = __reduce_min_sync(kFull, __float_as_uint(best2));
            const uint32_t w2 = __reduce_min_sync(kFull, (__float_as_uint(best2) == m2) ? k2 : 0xFFFFFFFFu);
            const uint32_t t2 = (w2 != 0xFFFFFFFFu) ? __ldg(lst + w2) : 0u;
            if (lane == j2) myTri = t2;
        }
    }
    if (!valid) return;
    f3 g = mk3(0.0f, 0.0f, 0.0f);
    float d;
    if (!inside) d = boxDistance(q, p) + q.outside;   // the reference leaves the caller's gradient untouched here (zero-initialised by us)
    else d = kGrad ? signedDistGradSelf(p, tris[myTri], g) : signedDistPointTriangle(p, tris[myTri]);
    dist[i] = d;
    if (kGrad) { grad[3 * i] = g.x; grad[3 * i + 1] = g.y; grad[3 * i + 2] = g.z; }
}

// ---- large batches: queries binned by leaf ----------------------------------------------------------------------
// The kernel above fetches an 80-byte frame per (query, triangle) pair and is bound by the L1 data pipe (65 % busy,
// profiles/r1_summary.md). Queries that land in the same leaf scan the same list, but on a regular grid the 8 queries
// of a finest leaf sit in 4 different warps. For large batches the queries are therefore counting-sorted by leaf
// first (walk + histogram, exclusive scan, scatter — no host synchronisation), then every warp takes 64 consecutive
// sorted queries and, per run of up to 8 queries of one leaf, strides the leaf's list ONCE: each lane loads its frame
// into registers and evaluates it against all queries of the run. Per query the arithmetic and the tie rule (first
// strict minimum over ascending list positions) are unchanged, so distances are bit-identical to the other kernel.
// Measured on C3, 16.7 M queries (value / with gradient identical): 256^3 grid 8.60 -> 8.00 ms, uniform random points
// (5 % outside the box) 10.9 -> 6.6 ms. The grid gains little: once the frame traffic is shared the kernel is bound
// by the ~100 non-contracted float instructions of each point-triangle evaluation (-fmad=false for bit parity).
constexpr uint32_t kBinSegment = 64;   // sorted queries per warp
constexpr uint32_t kBinRun = 8;        // queries of one leaf evaluated per pass over its list

template <bool kGrad>
__global__ void __launch_bounds__(256)
exactWalkKernel(const uint32_t* __restrict__ nodes, const ExactQueryParams q, const float* __restrict__ xyz, uint64_t n,
                uint32_t* __restrict__ leafOf, uint32_t* __restrict__ leafQueries, float* __restrict__ dist, float* __restrict__ grad) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f3 p = mk3(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2));
    float fx = (p.x - q.minx) / q.cell, fy = (p.y - q.miny) / q.cell, fz = (p.z - q.minz) / q.cell;
    const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    const int ix = int(flx), iy = int(fly), iz = int(flz);
    fx -= flx; fy -= fly; fz -= flz;
    if (ix < 0 || ix >= q.grid || iy < 0 || iy >= q.grid || iz < 0 || iz >= q.grid) {
        leafOf[i] = kNone;
        dist[i] = boxDistance(q, p) + q.outside;
        if (kGrad) { grad[3 * i] = 0.0f; grad[3 * i + 1] = 0.0f; grad[3 * i + 2] = 0.0f; }
        return;
    }
    uint32_t idx = uint32_t((iz * q.grid + iy) * q.grid + ix);
    uint32_t w0 = __ldg(nodes + 2 * size_t(idx));
    while (!(w0 & kLeafBit)) {
        const uint32_t child = ((fz > 0.5f) ? 4u : 0u) + ((fy > 0.5f) ? 2u : 0u) + ((fx > 0.5f) ? 1u : 0u);
        idx = (w0 & kExactIndexMask) + child;
        w0 = __ldg(nodes + 2 * size_t(idx));
        fx = 2.0f * fx; fy = 2.0f * fy; fz = 2.0f * fz;
        fx -= floorf(fx); fy -= floorf(fy); fz -= floorf(fz);
    }
    leafOf[i] = idx;
    atomicAdd(&leafQueries[idx], 1u);
}

// exclusive scan of the per-node query counts, no host round trip (three launches on the caller's stream)
__global__ void binBlockSums(const uint32_t* in, uint32_t* blockSums, uint64_t n) {
    __shared__ uint32_t warpSums[32];
    const uint64_t i = uint64_t(blockIdx.x) * kScanBlock + threadIdx.x;
    uint32_t v = i < n ? in[i] : 0u;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) warpSums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t t = warpSums[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) blockSums[blockIdx.x] = t;
    }
}

// Binning pays when leaves hold several queries each (a frame is then loaded once per run of up to 8 queries). With
// one query per leaf — e.g. a 256^3 grid over a depth-8 structure — it is pure overhead (C4: 30.7 ms against 26.9 ms),
// so the choice is made on the device from the histogram: bins iff queries >= 2 x non-empty leaves.
__global__ void countNonEmptyKernel(const uint32_t* __restrict__ counts, uint64_t n, uint32_t* nonEmpty) {
    uint32_t c = 0;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += uint64_t(gridDim.x) * blockDim.x) c += counts[i] ? 1u : 0u;
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(nonEmpty, c);
}
__global__ void chooseBinsKernel(const uint32_t* nonEmpty, const uint32_t* numSorted, uint32_t* useBins) {
    *useBins = (*numSorted >= 2u * *nonEmpty && *nonEmpty > 0u) ? 1u : 0u;
}

__global__ void exactScatterKernel(const uint32_t* __restrict__ leafOf, uint64_t n, const uint32_t* __restrict__ start,
                                   uint32_t* __restrict__ remaining, uint32_t* __restrict__ order) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t leaf = leafOf[i];
    if (leaf == kNone) return;
    order[start[leaf] + atomicSub(&remaining[leaf], 1u) - 1u] = uint32_t(i);
}

template <bool kGrad>
__global__ void __launch_bounds__(256)
exactBinnedKernel(const uint64_t* __restrict__ leafLo, const uint32_t* __restrict__ leafCnt, const uint32_t* __restrict__ pool,
                  const float4* __restrict__ frames, const TriData* __restrict__ tris, const float* __restrict__ xyz,
                  const uint32_t* __restrict__ leafOf, const uint32_t* __restrict__ order, const uint32_t* __restrict__ numSorted,
                  const uint32_t* __restrict__ useBins, float* __restrict__ dist, float* __restrict__ grad) {
    constexpr unsigned kFull = 0xffffffffu;
    __shared__ float sP[8][kBinSegment][3];
    __shared__ uint32_t sLeaf[8][kBinSegment], sTri[8][kBinSegment];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (*useBins == 0u) return;   // too few queries per leaf to share anything: the warp kernel runs instead
    const uint32_t total = *numSorted;
    const uint64_t seg0 = (uint64_t(blockIdx.x) * 8 + warp) * kBinSegment;
    if (seg0 >= total) return;
    const uint32_t m = uint32_t(total - seg0 < kBinSegment ? total - seg0 : kBinSegment);
    uint32_t qi[2] = {0, 0};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t k = lane + 32 * h;
        if (k < m) {
            qi[h] = order[seg0 + k];
            sP[warp][k][0] = __ldg(xyz + 3 * size_t(qi[h]));
            sP[warp][k][1] = __ldg(xyz + 3 * size_t(qi[h]) + 1);
            sP[warp][k][2] = __ldg(xyz + 3 * size_t(qi[h]) + 2);
            sLeaf[warp][k] = leafOf[qi[h]];
            sTri[warp][k] = 0u;   // the serial loop's initial bestTri (empty lists keep it)
        }
    }
    __syncwarp();
    uint32_t r = 0;
    while (r < m) {
        const uint32_t leaf = sLeaf[warp][r];
        uint32_t e = r + 1;
        while (e < m && e < r + kBinRun && sLeaf[warp][e] == leaf) e++;
        const uint32_t len = e - r;
        const uint32_t cnt = leafCnt[leaf];
        const uint32_t* lst = pool + leafLo[leaf];
        float best[kBinRun];
        uint32_t bestK[kBinRun];
#pragma unroll
        for (uint32_t c = 0; c < kBinRun; c++) { best[c] = INFINITY; bestK[c] = 0xFFFFFFFFu; }
        for (uint32_t k = lane; k < cnt; k += 32) {
            const TriFrame f = loadFrame(frames, __ldg(lst + k));
#pragma unroll
            for (uint32_t c = 0; c < kBinRun; c++)
                if (c < len) {
                    const float sq = sqDistPointTriangle(mk3(sP[warp][r + c][0], sP[warp][r + c][1], sP[warp][r + c][2]), f);
                    if (sq < best[c]) { best[c] = sq; bestK[c] = k; }
                }
        }
#pragma unroll
        for (uint32_t c = 0; c < kBinRun; c++)
            if (c < len) {
                const uint32_t mn = __reduce_min_sync(kFull, __float_as_uint(best[c]));
                const uint32_t w = __reduce_min_sync(kFull, (__float_as_uint(best[c]) == mn) ? bestK[c] : 0xFFFFFFFFu);
                if (lane == 0 && w != 0xFFFFFFFFu) sTri[warp][r + c] = __ldg(lst + w);
            }
        r = e;
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint32_t k = lane + 32 * h;
        if (k < m) {
            const f3 p = mk3(sP[warp][k][0], sP[warp][k][1], sP[warp][k][2]);
            f3 g = mk3(0.0f, 0.0f, 0.0f);
            const float d = kGrad ? signedDistGradSelf(p, tris[sTri[warp][k]], g) : signedDistPointTriangle(p, tris[sTri[warp][k]]);
            dist[qi[h]] = d;
            if (kGrad) { grad[3 * size_t(qi[h])] = g.x; grad[3 * size_t(qi[h]) + 1] = g.y; grad[3 * size_t(qi[h]) + 2] = g.z; }
        }
    }
}

}  // namespace

// Decode the public arrays into the private per-leaf pool (see the header comment). Throws ERR_IO when the
// arrays are inconsistent (untrusted .bin input).
void prepareExactQuery(sdfb200_sdf& s) {
    const uint32_t nT = s.numTris;
    const uint64_t numNodes = s.nOctree / 2;
    if (!s.qFrames) {   // loaded from a .bin: frames from the file's TriangleData (built structures share their mesh's)
        s.dFrames.alloc(size_t(nT) * 5);
        if (nT) framesFromTriData<<<divUp(uint64_t(nT) * 5, 256), 256>>>(s.dTris.p, s.dFrames.p, nT);
        s.qFrames = s.dFrames.p;
        s.qTris = s.dTris.p;
    }
    s.dLeafLo.alloc(numNodes);
    s.dLeafCnt.alloc(numNodes);
    SDFB_CUDA(cudaMemsetAsync(s.dLeafLo.p, 0, numNodes * 8));
    SDFB_CUDA(cudaMemsetAsync(s.dLeafCnt.p, 0, numNodes * 4));
    PublicView pv{s.dOctree.p, s.dSets.p, s.dMasks.p, numNodes, s.nSets, s.nMasks, nT, s.bitsPerIndex, s.bitEncodingStartDepth};
    DevBuf<uint32_t> err(1);
    SDFB_CUDA(cudaMemsetAsync(err.p, 0, 4));

    struct Frontier {
        uint32_t n = 0;
        DevBuf<uint32_t> nodeIdx, cnt, inner, leafCnt, childSlot, parentCnt, dec, pos;
        DevBuf<uint64_t> lo, parentLo, pairOff, poolOff;
        DevBuf<uint8_t> flags;
        uint64_t poolTotal = 0;
        DevBuf<uint32_t> leafPool;   // this level's leaves, concatenated
    };
    ScannerT<uint32_t, uint32_t> scan32;
    ScannerT<uint32_t, uint64_t> scan64;
    FlagScanner scanFlags;
    const uint64_t G3 = uint64_t(s.startGridSize) * s.startGridSize * s.startGridSize;
    std::vector<std::unique_ptr<Frontier>> fr;
    fr.emplace_back(new Frontier());
    {
        Frontier& F = *fr[0];
        F.n = uint32_t(G3);
        std::vector<uint32_t> ids(G3);
        for (uint32_t i = 0; i < G3; i++) ids[i] = i;
        F.nodeIdx.alloc(G3);
        F.nodeIdx.upload(ids.data(), G3);
        SDFB_CUDA(cudaDeviceSynchronize());
    }
    uint64_t poolTotal = 0;
    for (uint32_t depth = s.startDepth; !fr.empty() && fr.back()->n > 0; depth++) {
        if (depth > 32) throw Error(SDFB200_ERR_IO, "exact octree deeper than 32 levels");
        Frontier& F = *fr.back();
        const uint32_t n = F.n;
        F.cnt.alloc(n); F.inner.alloc(n); F.leafCnt.alloc(n); F.childSlot.alloc(n); F.lo.alloc(size_t(n) + 1); F.poolOff.alloc(size_t(n) + 1);
        if (depth <= s.bitEncodingStartDepth) {
            frontierCountsKernel<<<divUp(n, 256), 256>>>(pv, F.nodeIdx.p, n, depth, F.cnt.p, F.inner.p, F.leafCnt.p, err.p);
            const uint64_t total = scan64.run(F.cnt.p, F.lo.p, n, true);
            if (total >= (uint64_t(1) << 32)) throw Error(SDFB200_ERR_IO, "decoded triangle sets of one level exceed 2^32 entries");
            F.dec.alloc(total + 1);
            unpackSetsKernel<<<n, 128>>>(pv, F.nodeIdx.p, F.lo.p, F.cnt.p, F.dec.p, err.p);
        } else {
            Frontier& P = *fr[fr.size() - 2];
            maskCountKernel<<<divUp(n, 8), 256>>>(pv, F.nodeIdx.p, F.parentCnt.p, n, F.cnt.p, F.inner.p, F.leafCnt.p, err.p);
            const uint64_t total = scan64.run(F.cnt.p, F.lo.p, n, true);
            if (total >= (uint64_t(1) << 32)) throw Error(SDFB200_ERR_IO, "decoded triangle lists of one level exceed 2^32 entries");
            F.dec.alloc(total + 1);
            maskSelectKernel<<<divUp(n, 8), 256>>>(pv, F.nodeIdx.p, F.parentLo.p, F.parentCnt.p, F.lo.p, n, P.dec.p, F.dec.p);
        }
        SDFB_CUDA(cudaGetLastError());
        // leaves -> this level's pool segment
        F.poolTotal = scan64.run(F.leafCnt.p, F.poolOff.p, n, true);
        F.leafPool.alloc(F.poolTotal + 1);
        leafPoolKernel<<<n, 128>>>(F.nodeIdx.p, F.inner.p, F.lo.p, F.cnt.p, F.poolOff.p, F.dec.p, F.leafPool.p, s.dLeafLo.p, s.dLeafCnt.p);
        // next frontier
        const uint32_t nInner = scan32.run(F.inner.p, F.childSlot.p, n);
        if (uint64_t(nInner) * 8 > numNodes) throw Error(SDFB200_ERR_IO, "exact octree node array contains a cycle");
        std::unique_ptr<Frontier> N(new Frontier());
        N->n = nInner * 8;
        if (N->n) {
            N->nodeIdx.alloc(N->n); N->parentLo.alloc(N->n); N->parentCnt.alloc(N->n);
            nextFrontierKernel<<<divUp(n, 256), 256>>>(pv, F.nodeIdx.p, F.inner.p, F.childSlot.p, F.lo.p, F.cnt.p, n, N->nodeIdx.p,
                                                       N->parentLo.p, N->parentCnt.p);
        }
        SDFB_CUDA(cudaGetLastError());
        poolTotal += F.poolTotal;
        if (fr.size() >= 2) fr[fr.size() - 2]->dec.release();   // the parent's decoded lists have been consumed
        fr.emplace_back(std::move(N));
    }
    uint32_t hErr = 0;
    SDFB_CUDA(cudaMemcpy(&hErr, err.p, 4, cudaMemcpyDeviceToHost));
    if (hErr) throw Error(SDFB200_ERR_IO, std::string("inconsistent ExactOctreeSdf arrays:") + ((hErr & kErrSetRange) ? " set offset out of range" : "") +
                                              ((hErr & kErrMaskRange) ? " mask offset out of range" : "") +
                                              ((hErr & kErrTriangle) ? " triangle index out of range" : ""));
    // concatenate the per-level segments; leafLo of a level's leaves is rebased by the segment start
    s.dLeafPool.alloc(poolTotal + 1);
    uint64_t base = 0;
    for (auto& fp : fr) {
        Frontier& F = *fp;
        if (!F.n || !F.poolTotal) continue;
        SDFB_CUDA(cudaMemcpyAsync(s.dLeafPool.p + base, F.leafPool.p, F.poolTotal * 4, cudaMemcpyDeviceToDevice));
        if (base) rebaseLeavesKernel<<<divUp(F.n, 256), 256>>>(F.nodeIdx.p, F.inner.p, F.n, base, s.dLeafLo.p);
        base += F.poolTotal;
    }
    SDFB_CUDA(cudaDeviceSynchronize());
}

void launchExactQuery(const sdfb200_sdf& s, const float* dXyz, uint64_t n, float* dDist, float* dGrad, cudaStream_t st) {
    if (n == 0) return;
    if (!s.dLeafLo.p) throw Error(SDFB200_ERR_CUDA, "ExactOctreeSdf is not resident on a CUDA device");
    ExactQueryParams q;
    q.minx = s.boxMin[0]; q.miny = s.boxMin[1]; q.minz = s.boxMin[2];
    q.maxx = s.boxMax[0]; q.maxy = s.boxMax[1]; q.maxz = s.boxMax[2];
    q.cell = s.cellSize;
    q.grid = s.startGridSize;
    q.outside = sqrtf(3.0f) * (s.boxMax[0] - s.boxMin[0]);   // ExactOctreeSdf.cpp:48
    const uint32_t grid = uint32_t((n + 255) / 256);
    const uint64_t numNodes = s.dLeafCnt.n;
    if (n >= (uint64_t(1) << 15) && n < (uint64_t(1) << 32)) {   // binned path: two uint32 per query + three per node of scratch, stream-ordered
        uint32_t *leafOf = nullptr, *order = nullptr, *counts = nullptr, *start = nullptr, *blockSums = nullptr, *total = nullptr;
        const uint32_t nBlocks = divUp(numNodes + 1, kScanBlock);
        SDFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&leafOf), n * 4, st));
        SDFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&order), n * 4, st));
        SDFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&counts), (numNodes + 1) * 4, st));
        SDFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&start), (numNodes + 1) * 4, st));
        SDFB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&blockSums), (size_t(nBlocks) + 3) * 4, st));
        total = blockSums + nBlocks;
        uint32_t *nonEmpty = total + 1, *useBins = total + 2;
        SDFB_CUDA(cudaMemsetAsync(counts, 0, (numNodes + 1) * 4, st));
        SDFB_CUDA(cudaMemsetAsync(nonEmpty, 0, 4, st));
        if (dGrad) exactWalkKernel<true><<<grid, 256, 0, st>>>(s.dOctree.p, q, dXyz, n, leafOf, counts, dDist, dGrad);
        else exactWalkKernel<false><<<grid, 256, 0, st>>>(s.dOctree.p, q, dXyz, n, leafOf, counts, dDist, nullptr);
        countNonEmptyKernel<<<148 * 4, 256, 0, st>>>(counts, numNodes, nonEmpty);
        binBlockSums<<<nBlocks, kScanBlock, 0, st>>>(counts, blockSums, numNodes + 1);
        scanOfBlockSums<uint32_t><<<1, kScanBlock, 0, st>>>(blockSums, nBlocks, total);
        scanFinalize<uint32_t, uint32_t><<<nBlocks, kScanBlock, 0, st>>>(counts, blockSums, start, numNodes + 1, total, false);
        chooseBinsKernel<<<1, 1, 0, st>>>(nonEmpty, total, useBins);
        exactScatterKernel<<<grid, 256, 0, st>>>(leafOf, n, start, counts, order);
        const uint32_t segGrid = divUp(n, 8 * kBinSegment);
        if (dGrad) {
            exactBinnedKernel<true><<<segGrid, 256, 0, st>>>(s.dLeafLo.p, s.dLeafCnt.p, s.dLeafPool.p, s.qFrames, s.qTris, dXyz, leafOf, order, total, useBins, dDist, dGrad);
            exactQueryWarpKernel<true><<<grid, 256, 0, st>>>(s.dOctree.p, s.dLeafLo.p, s.dLeafCnt.p, s.dLeafPool.p, s.qFrames, s.qTris, q, dXyz, n, dDist, dGrad, useBins);
        } else {
            exactBinnedKernel<false><<<segGrid, 256, 0, st>>>(s.dLeafLo.p, s.dLeafCnt.p, s.dLeafPool.p, s.qFrames, s.qTris, dXyz, leafOf, order, total, useBins, dDist, nullptr);
            exactQueryWarpKernel<false><<<grid, 256, 0, st>>>(s.dOctree.p, s.dLeafLo.p, s.dLeafCnt.p, s.dLeafPool.p, s.qFrames, s.qTris, q, dXyz, n, dDist, nullptr, useBins);
        }
        SDFB_CUDA(cudaGetLastError());
        cudaFreeAsync(leafOf, st); cudaFreeAsync(order, st); cudaFreeAsync(counts, st); cudaFreeAsync(start, st); cudaFreeAsync(blockSums, st);
        return;
    }
    if (dGrad)
        exactQueryWarpKernel<true><<<grid, 256, 0, st>>>(s.dOctree.p, s.dLeafLo.p, s.dLeafCnt.p, s.dLeafPool.p, s.qFrames, s.qTris, q, dXyz, n, dDist, dGrad, nullptr);
    else
        exactQueryWarpKernel<false><<<grid, 256, 0, st>>>(s.dOctree.p, s.dLeafLo.p, s.dLeafCnt.p, s.dLeafPool.p, s.qFrames, s.qTris, q, dXyz, n, dDist, nullptr, nullptr);
    SDFB_CUDA(cudaGetLastError());
}

}  // namespace sdfb200
