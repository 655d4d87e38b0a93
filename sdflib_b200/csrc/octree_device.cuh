// Device code shared by the OctreeSdf builders (octree_build.cu: NO_CONTINUITY, octree_cont.cu: CONTINUITY):
// float64 sphere-BVH nearest-triangle descent, point sampling, the 64x64 Hermite map and the exact-order
// polynomial evaluation. Everything has internal linkage: each translation unit gets its own copy of the
// __constant__ tables (no relocatable device code).
#pragma once
#include <cfloat>
#include <cmath>
#include <algorithm>
#include <cstdlib>

#include "device_utils.cuh"
#include "sdf_internal.h"

namespace sdfb200 {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr uint32_t kNoChild = 0xFFFFFFFFu;

#include "bvh_sampler.cuh"   // cSampleLattice, BVH traversal, samplePoint, latticeSamplePosition, sampleOwners*Kernel


// Non-zero entries of the 64x64 Hermite map W = H(x)H(x)H, row-major, columns ascending.
struct alignas(16) HermiteTable {   // sizeof is a multiple of 16, so the word-wise copy below is exact
    uint16_t rowStart[65];
    uint8_t col[1000];
    int8_t weight[1000];
    uint8_t order[64];   // rows sorted by descending number of terms (load balance across lanes)
};
__constant__ HermiteTable cHermite;

HermiteTable makeHermiteTable() {
    static const int H[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {-3, -2, 3, -1}, {2, 1, -2, 1}};
    static const int slot[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
    HermiteTable t;
    int n = 0;
    for (int row = 0; row < 64; row++) {
        t.rowStart[row] = uint16_t(n);
        const int i = row & 3, j = (row >> 2) & 3, k = row >> 4;
        for (int c = 0; c < 8; c++)
            for (int s = 0; s < 8; s++) {
                const int w = H[i][2 * (c & 1) + slot[s][0]] * H[j][2 * ((c >> 1) & 1) + slot[s][1]] *
                              H[k][2 * ((c >> 2) & 1) + slot[s][2]];
                if (w != 0) { t.col[n] = uint8_t(c * 8 + s); t.weight[n] = int8_t(w); n++; }
            }
    }
    t.rowStart[64] = uint16_t(n);
    int o = 0;
    for (int want : {64, 16, 4, 1})
        for (int row = 0; row < 64; row++)
            if (t.rowStart[row + 1] - t.rowStart[row] == want) t.order[o++] = uint8_t(row);
    return t;
}

// The 19 mid-point samples of every node of a level, one thread per (node, sample): all 32 lanes of a warp
// traverse the BVH (a warp-per-node layout leaves 13 of 32 lanes idle during the dominant phase).
// out[(node * 19 + s) * stride] = (d, gx, gy, gz); with stride 2 the mixed-derivative half is zeroed.
__global__ void __launch_bounds__(kBvhThreads) sampleLatticeKernel(DeviceMesh mesh, const float4* centerHalf, uint32_t count, float4* out, int stride) {
    const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= uint64_t(count) * 19) return;
    const uint32_t node = uint32_t(t / 19), s = uint32_t(t % 19);
    const float4 ch = centerHalf[node];
    const int L = cSampleLattice[s];
    const f3 rel = mk3(float(L % 3 - 1), float((L / 3) % 3 - 1), float(L / 9 - 1));
    out[t * stride] = samplePoint(mesh, mk3(ch.x, ch.y, ch.z) + rel * ch.w);
    if (stride == 2) out[t * 2 + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- sampling with exact de-duplication ---------------------------------------------------------------------
// Neighbouring nodes share lattice points, and the nearest-triangle query is a pure function of the sample
// position, so positions that are BIT-IDENTICAL need one traversal only. (The positions of one lattice point
// computed from different nodes differ by an ulp about 40 % of the time, because each node's centre went through its
// own chain of float additions; those stay separate queries, which keeps every value exactly what the node would
// have computed on its own.) On a uniformly refined level 59 % of the 19 N samples are distinct.
// Open-addressing table of sample indices: a slot holds the index t of its owner, whose position is recomputed from
// the node arrays for the comparison, so no key has to be published next to the claim.

__global__ void dedupeInsertKernel(const float4* __restrict__ centerHalf, uint32_t nSamples, uint32_t* table, uint32_t mask, uint32_t* slotOf) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nSamples) return;
    const f3 p = latticeSamplePosition(centerHalf, t);
    const uint32_t bx = __float_as_uint(p.x), by = __float_as_uint(p.y), bz = __float_as_uint(p.z);
    uint32_t h = bx * 0x9E3779B1u ^ by * 0x85EBCA77u ^ bz * 0xC2B2AE3Du;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    uint32_t slot = h & mask;
    for (;;) {
        const uint32_t old = atomicCAS(&table[slot], 0xFFFFFFFFu, t);
        if (old == 0xFFFFFFFFu) break;
        const f3 q = latticeSamplePosition(centerHalf, old);   // any member of the slot's group has the group's position
        if (__float_as_uint(q.x) == bx && __float_as_uint(q.y) == by && __float_as_uint(q.z) == bz) {
            atomicMin(&table[slot], t);   // the owner is the LOWEST sample index: identical on every rank of a collective build
            break;
        }
        slot = (slot + 1) & mask;
    }
    slotOf[t] = slot;
}

__global__ void dedupeResolveKernel(const uint32_t* __restrict__ table, uint32_t nSamples, uint32_t* rep /* in: slot, out: owner */, uint32_t* isOwner) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nSamples) return;
    const uint32_t owner = table[rep[t]];
    rep[t] = owner;
    isOwner[t] = owner == t ? 1u : 0u;
}

__global__ void dedupeOwnersKernel(const uint32_t* __restrict__ isOwner, const uint32_t* __restrict__ pos, uint32_t nSamples, uint32_t* owners) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nSamples && isOwner[t]) owners[pos[t]] = t;
}

inline bool sampleRefillEnabled() {
    // on by default since round 2 (NO_CONTINUITY levels of the C2 build 82.6 -> 74.4 ms, CONTINUITY 113.4 -> 110.0 ms, output
    // hash unchanged: profiles/r2_summary.md); SDFB200_SAMPLE_REFILL=0 selects the one-sample-per-thread kernel
    static const bool on = [] { const char* e = std::getenv("SDFB200_SAMPLE_REFILL"); return !(e && e[0] == '0'); }();
    return on;
}

// ---- start order of a level's traversals -------------------------------------------------------------------------------------
// A level's launch lasts as long as its longest traversals when those start late: samples far from the surface (deep inside the
// mesh, or in the corners of the box) visit thousands of nodes — milliseconds of dependent steps — while the bulk takes a few
// hundred. With the lane-refill schedule the samples are handed out in index order; here they are bucketed by the distance the
// node's corners already know (min |corner value|, relative to the box) and the far buckets are handed out first.
__global__ void nodeCostKernel(const float4* __restrict__ cornerValues, int perNode, int perCorner, uint32_t count, float invScale, float* __restrict__ cost) {
    const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= count) return;
    float m = INFINITY;
    for (int k = 0; k < 8; k++) m = fminf(m, fabsf(cornerValues[size_t(node) * perNode + size_t(k) * perCorner].x));
    cost[node] = m * invScale;
}
__device__ __forceinline__ uint32_t costBucket(float c) { return c > 0.25f ? 0u : (c > 0.12f ? 1u : (c > 0.05f ? 2u : 3u)); }
__global__ void scheduleCountKernel(const uint32_t* __restrict__ owners, uint32_t n, const float* __restrict__ nodeCost, uint32_t* counts) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = u < n ? costBucket(nodeCost[owners[u] / 19u]) : 4u;
    for (uint32_t k = 0; k < 4; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, b == k);
        if (m && (threadIdx.x & 31u) == uint32_t(__ffs(int(m)) - 1)) atomicAdd(counts + k, uint32_t(__popc(m)));
    }
}
__global__ void scheduleOffsetsKernel(uint32_t* counts) {   // counts[0..3] -> cursors[4..7] (exclusive prefix, far buckets first)
    uint32_t run = 0;
    for (int k = 0; k < 4; k++) { counts[4 + k] = run; run += counts[k]; }
}
__global__ void scheduleScatterKernel(const uint32_t* __restrict__ owners, uint32_t n, const float* __restrict__ nodeCost, uint32_t* counts, uint32_t* __restrict__ schedule) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = u < n ? costBucket(nodeCost[owners[u] / 19u]) : 4u;
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t k = 0; k < 4; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, b == k);
        if (!m) continue;
        const int leader = __ffs(int(m)) - 1;
        uint32_t base = 0;
        if (int(lane) == leader) base = atomicAdd(counts + 4 + k, uint32_t(__popc(m)));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (b == k) schedule[base + uint32_t(__popc(m & ((1u << lane) - 1u)))] = u;
    }
}
inline bool sampleScheduleEnabled() {   // A/B switch: SDFB200_SAMPLE_SCHEDULE=0 hands the samples out in index order
    static const bool on = [] { const char* e = std::getenv("SDFB200_SAMPLE_SCHEDULE"); return !(e && e[0] == '0'); }();
    return on;
}

inline int sampleLeafBatch() {   // lanes of a warp that must hold a leaf before the leaf branch runs (1 = take leaves as they come)
    static const int v = [] { const char* e = std::getenv("SDFB200_LEAF_BATCH"); const int x = e ? std::atoi(e) : 1; return x < 1 ? 1 : (x > 32 ? 32 : x); }();
    return v;
}

__global__ void dedupeScatterKernel(const uint32_t* __restrict__ rep, const uint32_t* __restrict__ pos, const float4* __restrict__ results,
                                    uint32_t nSamples, float4* out, int stride) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nSamples) return;
    out[size_t(t) * stride] = results[pos[rep[t]]];
    if (stride == 2) out[size_t(t) * 2 + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// The 19 mid-point samples of every node of a level into out[(node * 19 + s) * stride]; returns the number of BVH
// traversals actually run. Workspace buffers are grow-only and live as long as the builder.
struct LevelSampler {
    DevBuf<uint32_t> table, rep, isOwner, pos, owners;
    DevBuf<float4> results, slice;
    DevBuf<uint32_t> refillCounter;   // lane-refill schedule: next unassigned sample of the launch
    DevBuf<uint32_t> schedule, scheduleCounts;   // ... and the order in which the samples are started (far ones first)
    Scanner scanner;
    SampleExchange exchange;   // world > 1: every rank traverses the BVH for its slice of the distinct positions only
    cudaStream_t stream = nullptr;   // where this sampler's launches go (legacy default stream unless a pass runs on a side stream)

    // results[0 .. n) for n work items, item u computed by launch(first, count, dst) -> dst[0 .. count) = items first ..
    template <class Launch> const float4* shared(uint32_t n, Launch launch) {
        if (exchange.world <= 1) {
            results.ensure(std::max<uint32_t>(n, 1));
            if (n) launch(0u, n, results.p);
            return results.p;
        }
        const uint32_t per = (n + exchange.world - 1) / exchange.world;   // equal blocks, the tail of the last ones is padding
        results.ensure(size_t(std::max<uint32_t>(per, 1)) * exchange.world);
        slice.ensure(std::max<uint32_t>(per, 1));
        if (per == 0) return results.p;
        const uint32_t first = std::min<uint64_t>(uint64_t(exchange.rank) * per, n), count = std::min<uint32_t>(per, n - first);
        SDFB_CUDA(cudaMemsetAsync(slice.p, 0, size_t(per) * sizeof(float4), stream));
        if (count) launch(first, count, slice.p);
        if (exchange.allgather(exchange.user, slice.p, results.p, uint64_t(per) * sizeof(float4)) != 0)
            throw Error(SDFB200_ERR_CUDA, "the all-gather hook of the collective build failed");
        return results.p;   // item u sits at block u / per, offset u % per = index u
    }

    // nodeCost (optional): per node, min |corner distance| / box size — decides the start order of the traversals only
    uint32_t run(const DeviceMesh& mesh, const float4* centerHalf, uint32_t count, float4* out, int stride, const float* nodeCost = nullptr) {
        const uint64_t n64 = uint64_t(count) * 19;
        if (n64 >= (uint64_t(1) << 30)) {   // the table (2 n slots, 32-bit sample indices) would pass 2^31 slots: plain path
            if (exchange.world > 1) throw Error(SDFB200_ERR_INVALID, "more than 2^30 samples on one level of a collective build");
            sampleLatticeKernel<<<divUp(n64, kBvhThreads), kBvhThreads, bvhStackBytes(mesh), stream>>>(mesh, centerHalf, count, out, stride);
            return 0xFFFFFFFFu;
        }
        const uint32_t n = uint32_t(n64);
        uint32_t size = 1024;
        while (size < 2 * n) size <<= 1;
        table.ensure(size); rep.ensure(n); isOwner.ensure(n); pos.ensure(n);
        SDFB_CUDA(cudaMemsetAsync(table.p, 0xFF, size_t(size) * 4, stream));
        dedupeInsertKernel<<<divUp(n, 256), 256, 0, stream>>>(centerHalf, n, table.p, size - 1, rep.p);
        dedupeResolveKernel<<<divUp(n, 256), 256, 0, stream>>>(table.p, n, rep.p, isOwner.p);
        const uint32_t nUnique = scanner.run(isOwner.p, pos.p, n, false, stream);
        owners.ensure(nUnique);
        dedupeOwnersKernel<<<divUp(n, 256), 256, 0, stream>>>(isOwner.p, pos.p, n, owners.p);
        const uint32_t* ownersPtr = owners.p;
        const uint32_t* schedulePtr = nullptr;
        if (nodeCost && nUnique > 8192 && exchange.world <= 1 && sampleRefillEnabled() && sampleScheduleEnabled()) {
            schedule.ensure(nUnique); scheduleCounts.ensure(8);
            SDFB_CUDA(cudaMemsetAsync(scheduleCounts.p, 0, 8 * sizeof(uint32_t), stream));
            scheduleCountKernel<<<divUp(nUnique, 256), 256, 0, stream>>>(owners.p, nUnique, nodeCost, scheduleCounts.p);
            scheduleOffsetsKernel<<<1, 1, 0, stream>>>(scheduleCounts.p);
            scheduleScatterKernel<<<divUp(nUnique, 256), 256, 0, stream>>>(owners.p, nUnique, nodeCost, scheduleCounts.p, schedule.p);
            schedulePtr = schedule.p;
        }
        const float4* res = shared(nUnique, [&](uint32_t first, uint32_t cnt, float4* dst) {
            if (sampleRefillEnabled()) {   // see sampleOwnersRefillKernel
                refillCounter.ensure(1);
                SDFB_CUDA(cudaMemsetAsync(refillCounter.p, 0, sizeof(uint32_t), stream));
                const uint32_t blocks = std::min<uint32_t>(divUp(cnt, kBvhThreads), 148u * 8u);
                sampleOwnersRefillKernel<<<blocks, kBvhThreads, bvhStackBytes(mesh), stream>>>(mesh, centerHalf, ownersPtr, first, cnt, dst, refillCounter.p, sampleLeafBatch(), schedulePtr);
                finishOwnersKernel<<<divUp(cnt, 256), 256, 0, stream>>>(mesh, centerHalf, ownersPtr, first, cnt, dst);
                return;
            }
            sampleOwnersKernel<<<divUp(cnt, kBvhThreads), kBvhThreads, bvhStackBytes(mesh), stream>>>(mesh, centerHalf, ownersPtr, first, cnt, dst);
        });
        dedupeScatterKernel<<<divUp(n, 256), 256, 0, stream>>>(rep.p, pos.p, res, n, out, stride);
        return nUnique;
    }

    // explicit point list (fix-up pass of the CONTINUITY builder)
    const float4* runPoints(const DeviceMesh& mesh, const float4* points, uint32_t n);
};

// explicit point list (fix-up pass of the CONTINUITY builder)
__global__ void __launch_bounds__(kBvhThreads) samplePointsKernel(DeviceMesh mesh, const float4* points, uint32_t first, uint32_t n, float4* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float4 p = points[first + t];
    out[t] = samplePoint(mesh, mk3(p.x, p.y, p.z));
}

inline const float4* LevelSampler::runPoints(const DeviceMesh& mesh, const float4* points, uint32_t n) {
    return shared(n, [&](uint32_t first, uint32_t cnt, float4* dst) {
        samplePointsKernel<<<divUp(cnt, kBvhThreads), kBvhThreads, bvhStackBytes(mesh), stream>>>(mesh, points, first, cnt, dst);
    });
}

// interpolateValue, scalar branch: acc = 0 + sum_n ((c_n * x^i) * y^j) * z^k, n ascending, left to right.
__device__ __forceinline__ float polyValueExact(const float* c, float x, float y, float z) {
    float acc = 0.0f;
#pragma unroll
    for (int n = 0; n < 64; n++) {
        float t = c[n];
#pragma unroll
        for (int a = 0; a < (n & 3); a++) t *= x;
#pragma unroll
        for (int a = 0; a < ((n >> 2) & 3); a++) t *= y;
#pragma unroll
        for (int a = 0; a < (n >> 4); a++) t *= z;
        acc += t;
    }
    return acc;
}

// One Hermite row: left-to-right float sum of w * in[col] over the row's non-zero columns.
// in[col] = corner value scaled by nodeSize^order (slots 4..7 are the always-zero mixed derivatives).
__device__ __forceinline__ float hermiteRow(const HermiteTable& tab, int row, const float4* lattice, float nodeSize) {
    const int b = tab.rowStart[row], e = tab.rowStart[row + 1];
    float acc = 0.0f;
    for (int i = b; i < e; i++) {
        const int col = tab.col[i], c = col >> 3, s = col & 7;
        // corner c = (x,y,z) bits -> lattice point (2x, 2y, 2z)
        const float4 v = lattice[2 * (c & 1) + 6 * ((c >> 1) & 1) + 18 * (c >> 2)];
        float in;
        if (s == 0) in = v.x;
        else if (s == 1) in = v.y * nodeSize;
        else if (s == 2) in = v.z * nodeSize;
        else if (s == 3) in = v.w * nodeSize;
        else in = 0.0f;   // 0 * nodeSize^2 (or ^3) = +0
        const float term = float(int(tab.weight[i])) * in;
        acc = (i == b) ? term : acc + term;
    }
    return acc;
}

void uploadHermite() {
    static bool done[64] = {};
    int dev = 0;
    SDFB_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && done[dev]) return;
    const HermiteTable t = makeHermiteTable();
    SDFB_CUDA(cudaMemcpyToSymbol(cHermite, &t, sizeof(t)));
    if (dev < 64) done[dev] = true;
}

// host-side phases of the mesh preparation into the build statistics (mesh_device.cu measured them)
inline void meshStats(const PreparedMesh& pm, sdfb200_build_stats& st) {
    st.triangle_data_ms = pm.triangleDataMs;
    st.bvh_ms = pm.bvhMs;
    st.upload_ms = pm.uploadMs;
}

// Weight of a mid-point in the error integral: trapezoid / by-distance 2^k/64 (OctreeSdfUtils.h:60-138),
// Simpson 4^k/216 (:213-238), k = number of centred axes; the float constant is formed first, as in the
// reference expression `w / 64.0f * pow2(..)`.
__device__ __forceinline__ float errorWeight(int rule, int centred) {
    if (rule == SDFB200_RULE_SIMPSONS) return (centred == 1 ? 4.0f : (centred == 2 ? 16.0f : 64.0f)) / 216.0f;
    return (centred == 1 ? 2.0f : (centred == 2 ? 4.0f : 8.0f)) / 64.0f;
}

}  // namespace
}  // namespace sdfb200
