// OctreeSdf construction with C1 continuity across T-junctions (InitAlgorithm::CONTINUITY) on the GPU.
//
// Replaces, behind sdfb200_build_octree(initAlgorithm = CONTINUITY), the reference's breadth-first builder
//   OctreeSdf::initOctreeWithContinuityNoDelay<VHQueries<TriCubicInterpolation>>
//                                                   src/sdf/OctreeSdfBreadthFirstNoDelay.h:84-1224
//   getNeighboursVector / ...InUniformGrid          src/sdf/OctreeSdfBreadthFirst.h:47-89
//   TriCubicInterpolation::interpolateVertexValues  include/SdfLib/InterpolationMethods.h:457-497
//
// What the reference does per depth: Iter 1 (parallel) samples + fits + decides every node; Iter 2 (serial)
// finds, for every subdividing node, the mid-point samples lying on a face/edge shared with a coarser-or-equal
// LEAF, replaces them by the node's own polynomial where that is within the threshold (this is what makes the
// field C1 across T-junctions) and otherwise queues that leaf; the fix-up pass (serial) re-opens queued leaves.
//
// B200 design. The reference's 6 neighbour links per node are an incremental neighbour finder; what they
// compute is purely geometric (checked against the link machinery on the CPU oracle): "the same-depth cell at
// coords + dir is covered by a leaf / is a regular inner node". Here every node keeps its integer coordinates
// and a probe is a descent from the start grid through the node words already in HBM (<= depth - startDepth
// dependent L2 loads, 18 probes per node on 18 lanes). The only genuinely sequential state of the reference is
// the append cursor of mOctreeData, i.e. the ARRAY ORDER:
//   * Iter 2 order  = exclusive scan of (8 | 64) words over the level's nodes;
//   * fix-up order  = queue order of the first occurrence of each leaf, then breadth-first inside each
//                     re-opened leaf = (root rank, relative depth, child path) — reproduced with per-round
//                     scans and a (root x round) table of segment sizes, no serial pass.
// Fix-up topology is order independent: a node splits iff one of its 18 same-depth neighbours is a regular
// (unmarked) inner node, which no fix-up changes. Values depend only on the node's own corner values.
// Arithmetic is bit-faithful (-fmad=false); the history-dependent 32^3 vertex cache is not emulated
// (DESIGN.md "parity"): oracle(use_cache=0) is the bit-exact comparison.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "device_utils.cuh"
#include "octree_device.cuh"
#include "sdf_internal.h"

namespace sdfb200 {

namespace {

constexpr uint32_t kMark = 1u << 30;
constexpr uint32_t kFreshChild = ~(7u << 29);   // children words until their own Iter 1 (:525)
constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int kRefShift = 27;                   // leaf reference = store id << 27 | index in store
constexpr uint32_t kRefIndexMask = (1u << kRefShift) - 1;
constexpr int kPoolStoreBase = 11;              // stores 0..10: regular levels, 11 + d: fix-up leaves of the pass at depth d

// the 18 face / edge directions in neighbourMasks order (:139-176): entry = 4 * (dir - 1) + sign
__constant__ int cEntry[18] = {0, 1, 4, 5, 8, 9, 10, 11, 12, 13, 16, 17, 18, 19, 20, 21, 22, 23};
__constant__ uint32_t cFaceMask[24];

// A node store: SoA records, either one regular level or the leaves one fix-up pass created.
struct StoreView {
    const float4* centerHalf;
    const uint32_t* coord;
    const uint8_t* depth;      // null: uniformDepth
    const uint32_t* word;      // index of the node's word in the octree array
    const float4* values;      // 16 per node: corner c -> [2c] = (f, fx, fy, fz), [2c+1] = (fxy, fxz, fyz, fxyz)
    uint32_t uniformDepth;
    uint32_t count;
};

struct Grid { uint32_t G, startDepth; float boxMin[3]; float cellSize; };

__device__ __forceinline__ int axisSign(uint32_t dir, uint32_t sign, int axis) {
    int bit = 0;
    for (int a = 0; a < axis; a++) if (dir & (1u << a)) bit++;
    return ((sign >> bit) & 1u) ? 1 : -1;
}

// Same-depth neighbour of cell q at depth nd. 0: outside the grid; 1: covered by a leaf (w = its word);
// 2: inner node at depth nd (w = its word); 3 (fix-up probes only): closed by a re-opened (marked) node.
__device__ __forceinline__ int probeCell(const uint32_t* __restrict__ oct, const Grid& g, uint32_t nd, int qx, int qy, int qz, bool fix,
                                         uint32_t& w) {
    const int res = 1 << nd;
    if (qx < 0 || qy < 0 || qz < 0 || qx >= res || qy >= res || qz >= res) return 0;
    const uint32_t sh = nd - g.startDepth;
    w = uint32_t(qz >> sh) * g.G * g.G + uint32_t(qy >> sh) * g.G + uint32_t(qx >> sh);
    for (uint32_t b = sh; b > 0; b--) {
        const uint32_t v = oct[w];
        if (v & kLeafBit) return 1;
        if (fix && (v & kMark)) return 3;
        const uint32_t cid = ((uint32_t(qx) >> (b - 1)) & 1u) | (((uint32_t(qy) >> (b - 1)) & 1u) << 1) | (((uint32_t(qz) >> (b - 1)) & 1u) << 2);
        w = (v & kOctIndexMask) + cid;
    }
    const uint32_t v = oct[w];
    if (v & kLeafBit) return 1;
    if (fix && (v & kMark)) return 3;
    return 2;
}

__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// One Hermite row with all eight value slots live (mixed derivatives are non-zero once a corner value has
// been replaced by interpolateVertexValues). in[c][s] = value * nodeSize^order, then the row is the
// left-to-right float sum of w * in[col] (InterpolationMethods.h:292-378).
__device__ __forceinline__ float hermiteRow8(const HermiteTable& tab, int row, const float4 (*lattice)[2], float nodeSize) {
    const int b = tab.rowStart[row], e = tab.rowStart[row + 1];
    const float sq = nodeSize * nodeSize;
    const float cu = sq * nodeSize;
    float acc = 0.0f;
    for (int i = b; i < e; i++) {
        const int col = tab.col[i], c = col >> 3, s = col & 7;
        const int L = 2 * (c & 1) + 6 * ((c >> 1) & 1) + 18 * (c >> 2);
        const float v = comp(lattice[L][s >> 2], s & 3);
        const float in = (s == 0) ? v : (s < 4 ? v * nodeSize : (s < 7 ? v * sq : v * cu));
        const float term = float(int(tab.weight[i])) * in;
        acc = (i == b) ? term : acc + term;
    }
    return acc;
}

// d^(OX+OY+OZ)/dx^OX dy^OY dz^OZ of the polynomial: terms in increasing n, the integer factor is applied to
// the coefficient first, powers multiplied in x, y, z order, left-to-right sum starting with the first term
// (InterpolationMethods.h:442-497).
template <int OX, int OY, int OZ>
__device__ __forceinline__ float polyDerivExact(const float* c, float x, float y, float z) {
    float acc = 0.0f;
    bool first = true;
#pragma unroll
    for (int n = 0; n < 64; n++) {
        const int i = n & 3, j = (n >> 2) & 3, k = n >> 4;
        if (i < OX || j < OY || k < OZ) continue;
        int w = 1;
#pragma unroll
        for (int a = 0; a < OX; a++) w *= (i - a);
#pragma unroll
        for (int a = 0; a < OY; a++) w *= (j - a);
#pragma unroll
        for (int a = 0; a < OZ; a++) w *= (k - a);
        float t = float(w) * c[n];
#pragma unroll
        for (int a = 0; a < i - OX; a++) t *= x;
#pragma unroll
        for (int a = 0; a < j - OY; a++) t *= y;
#pragma unroll
        for (int a = 0; a < k - OZ; a++) t *= z;
        acc = first ? t : acc + t;
        first = false;
    }
    return acc;
}

// interpolateVertexValues (:457-497)
__device__ __forceinline__ void vertexValues(const float* c, float x, float y, float z, float nodeSize, float4& lo, float4& hi) {
    const float sq = nodeSize * nodeSize;
    lo.x = polyValueExact(c, x, y, z);
    lo.y = polyDerivExact<1, 0, 0>(c, x, y, z) / nodeSize;
    lo.z = polyDerivExact<0, 1, 0>(c, x, y, z) / nodeSize;
    lo.w = polyDerivExact<0, 0, 1>(c, x, y, z) / nodeSize;
    hi.x = polyDerivExact<1, 1, 0>(c, x, y, z) / sq;
    hi.y = polyDerivExact<1, 0, 1>(c, x, y, z) / sq;
    hi.z = polyDerivExact<0, 1, 1>(c, x, y, z) / sq;
    hi.w = polyDerivExact<1, 1, 1>(c, x, y, z) / (sq * nodeSize);
}

__device__ __forceinline__ void loadHermite(HermiteTable& tab) {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&cHermite);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&tab);
    for (uint32_t i = threadIdx.x; i < sizeof(HermiteTable) / 4; i += blockDim.x) dst[i] = src[i];
    __syncthreads();
}

__device__ __forceinline__ int cornerLattice(int c) { return 2 * (c & 1) + 6 * ((c >> 1) & 1) + 18 * (c >> 2); }

__device__ __forceinline__ void fitCoefficients(const HermiteTable& tab, const float4 (*lattice)[2], float nodeSize, float* coeff, int lane) {
    const int r0 = tab.order[lane], r1 = tab.order[32 + lane];
    coeff[r0] = hermiteRow8(tab, r0, lattice, nodeSize);
    coeff[r1] = hermiteRow8(tab, r1, lattice, nodeSize);
}

__device__ __forceinline__ uint32_t startSlotOf(const Grid& g, f3 center) {   // :282-286
    const f3 f = (center - mk3(g.boxMin[0], g.boxMin[1], g.boxMin[2])) / g.cellSize;
    const int x = int(floorf(f.x)), y = int(floorf(f.y)), z = int(floorf(f.z));
    return uint32_t(z * int(g.G * g.G) + y * int(g.G) + x);
}

// ---- regular levels --------------------------------------------------------------------------------------

struct LevelArrays {
    uint32_t count;
    float4* centerHalf;
    uint32_t* coord;
    uint32_t* word;
    float4* values;     // 16 per node
    uint8_t* terminal;
};

// geometry of the children of a level whose nodes all subdivide (virtual levels): as contEmitKernel computes it, no sample needed
__global__ void contChildGeometryKernel(uint32_t count, const float4* centerHalf, const uint32_t* coord, float4* nextCenterHalf, uint32_t* nextCoord) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count * 8u) return;
    const uint32_t node = e >> 3, lane = e & 7u;
    const float4 ch = centerHalf[node];
    const float h = 0.5f * ch.w;
    const f3 c = mk3(ch.x, ch.y, ch.z) + cornerDir(lane) * h;
    nextCenterHalf[e] = make_float4(c.x, c.y, c.z, h);
    const uint32_t pc = coord[node];
    const uint32_t ix = ((pc & 1023u) << 1) | (lane & 1u), iy = (((pc >> 10) & 1023u) << 1) | ((lane >> 1) & 1u), iz = (((pc >> 20) & 1023u) << 1) | (lane >> 2);
    nextCoord[e] = ix | (iy << 10) | (iz << 20);
}

// speculative samples of ALL children of a level -> the samples of the children that exist (sub / subScan of the parents' Iter 2)
__global__ void contGatherSpeculativeKernel(const uint32_t* __restrict__ sub, const uint32_t* __restrict__ subScan, uint32_t parents, const float4* __restrict__ spec,
                                            float4* __restrict__ mids) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    constexpr uint64_t perParent = 8 * 38;
    if (i >= uint64_t(parents) * perParent) return;
    const uint32_t p = uint32_t(i / perParent), r = uint32_t(i % perParent);
    if (sub[p]) mids[size_t(subScan[p]) * perParent + r] = spec[i];
}

// corner samples of the seed nodes (:197-221)
__global__ void contSeedKernel(DeviceMesh mesh, Grid g, LevelArrays lv, uint32_t depth) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lv.count * 8) return;
    const float4 ch = lv.centerHalf[i >> 3];
    const f3 c = mk3(ch.x, ch.y, ch.z);
    lv.values[size_t(i >> 3) * 16 + 2 * (i & 7u)] = samplePoint(mesh, c + cornerDir(i & 7u) * ch.w);
    lv.values[size_t(i >> 3) * 16 + 2 * (i & 7u) + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((i & 7u) == 0) lv.word[i >> 3] = (depth == g.startDepth) ? startSlotOf(g, c) : kNone;
}

// Iter 1 (:258-369): fit, error integral, provisional leaf bit. One warp per node; the 19 true samples were
// taken by sampleLatticeKernel (one thread per sample).
__global__ void __launch_bounds__(kWarpsPerCta * 32)
contDecideKernel(LevelArrays lv, const float4* mids, float* coeffOut, uint32_t* oct, int rule, float sqThreshold, float decay) {
    __shared__ HermiteTable tab;
    __shared__ float4 lattice[kWarpsPerCta][27][2];
    __shared__ float coeff[kWarpsPerCta][64];
    __shared__ float terms[kWarpsPerCta][19];
    loadHermite(tab);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t node = blockIdx.x * kWarpsPerCta + warp;
    if (node >= lv.count) return;
    const float4 ch = lv.centerHalf[node];
    if (lane < 16) lattice[warp][cornerLattice(lane >> 1)][lane & 1] = lv.values[size_t(node) * 16 + lane];
    if (lane < 19) {
        lattice[warp][cSampleLattice[lane]][0] = mids[size_t(node) * 38 + 2 * lane];
        lattice[warp][cSampleLattice[lane]][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    const float nodeSize = 2.0f * ch.w;
    fitCoefficients(tab, lattice[warp], nodeSize, coeff[warp], lane);
    __syncwarp();
    coeffOut[size_t(node) * 64 + lane] = coeff[warp][lane];
    coeffOut[size_t(node) * 64 + 32 + lane] = coeff[warp][32 + lane];
    if (lane < 19) {
        const int L = cSampleLattice[lane];
        const int lx = L % 3, ly = (L / 3) % 3, lz = L / 9;
        const float v = polyValueExact(coeff[warp], 0.5f * float(lx), 0.5f * float(ly), 0.5f * float(lz));
        const float w = errorWeight(rule, (lx == 1) + (ly == 1) + (lz == 1));
        const float truth = lattice[warp][L][0].x;
        float d;
        if (rule == SDFB200_RULE_BY_DISTANCE) d = gmax(gabs(truth - v) - decay * gabs(v), 0.0f);
        else d = truth - v;
        terms[warp][lane] = w * (d * d);
    }
    __syncwarp();
    if (lane == 0) {
        float value;
        if (rule == SDFB200_RULE_NONE) value = INFINITY;
        else {
            value = terms[warp][0];
            for (int s = 1; s < 19; s++) value += terms[warp][s];
        }
        const bool terminal = value < sqThreshold;
        lv.terminal[node] = terminal ? 1 : 0;
        oct[lv.word[node]] = (terminal ? kLeafBit : 0u) | kOctIndexMask;   // setValues(terminal, uint32 max) (:367)
    }
}

// Iter 2, first half (:399-513): T-junction samples of the subdividing nodes. One warp per node.
__global__ void __launch_bounds__(kWarpsPerCta * 32)
contJunctionKernel(Grid g, LevelArrays lv, uint32_t depth, float4* mids, const float* coeffIn, const uint32_t* oct, float sqThreshold,
                   uint8_t* candCount, uint32_t* candWords) {
    __shared__ float coeff[kWarpsPerCta][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t node = blockIdx.x * kWarpsPerCta + warp;
    if (node >= lv.count) return;
    if (lv.terminal[node]) { if (lane == 0) candCount[node] = 0; return; }
    const uint32_t pc = lv.coord[node];
    const int ix = int(pc & 1023u), iy = int((pc >> 10) & 1023u), iz = int(pc >> 20);
    bool leafFound = false;
    uint32_t nbWord = kNone, myMask = 0;
    if (lane < 18) {
        const int entry = cEntry[lane];
        const uint32_t dir = uint32_t(entry >> 2) + 1u, sign = uint32_t(entry & 3);
        const int qx = ix + ((dir & 1u) ? axisSign(dir, sign, 0) : 0), qy = iy + ((dir & 2u) ? axisSign(dir, sign, 1) : 0),
                  qz = iz + ((dir & 4u) ? axisSign(dir, sign, 2) : 0);
        leafFound = probeCell(oct, g, depth, qx, qy, qz, false, nbWord) == 1;
        myMask = cFaceMask[entry];
    }
    const uint32_t samplesMask = __reduce_or_sync(0xffffffffu, leafFound ? myMask : 0u);
    uint32_t myBit = 0;
    if (samplesMask) {   // warp-uniform
        coeff[warp][lane] = coeffIn[size_t(node) * 64 + lane];
        coeff[warp][32 + lane] = coeffIn[size_t(node) * 64 + 32 + lane];
        __syncwarp();
        if (lane < 19 && (samplesMask & (1u << (18 - lane)))) {
            const int L = cSampleLattice[lane];
            const float x = 0.5f * float(L % 3), y = 0.5f * float((L / 3) % 3), z = 0.5f * float(L / 9);
            const float inter = polyValueExact(coeff[warp], x, y, z);
            const float d = mids[size_t(node) * 38 + 2 * lane].x - inter;
            if (d * d > sqThreshold) myBit = 1u << (18 - lane);
            else {
                float4 lo, hi;
                vertexValues(coeff[warp], x, y, z, 2.0f * lv.centerHalf[node].w, lo, hi);
                mids[size_t(node) * 38 + 2 * lane] = lo;
                mids[size_t(node) * 38 + 2 * lane + 1] = hi;
            }
        }
    }
    const uint32_t subdivisionMask = __reduce_or_sync(0xffffffffu, myBit);
    const bool cand = leafFound && (subdivisionMask & myMask);
    const uint32_t ballot = __ballot_sync(0xffffffffu, cand);
    if (cand) candWords[size_t(node) * 18 + __popc(ballot & ((1u << lane) - 1u))] = nbWord;
    if (lane == 0) candCount[node] = uint8_t(__popc(ballot));
}

__global__ void contSizesKernel(LevelArrays lv, bool real, bool deepest, uint32_t* sizes, uint32_t* sub) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= lv.count) return;
    const bool leaf = deepest || lv.terminal[i];
    sizes[i] = real ? (leaf ? 64u : 8u) : 0u;
    sub[i] = leaf ? 0u : 1u;
}

// Iter 2, second half (:515-733): node word, leaf coefficients or the 8 children. One warp per node.
__global__ void __launch_bounds__(kWarpsPerCta * 32)
contEmitKernel(Grid g, LevelArrays lv, LevelArrays next, uint32_t depth, bool real, bool deepest, const float4* mids, const float* coeffIn,
               const uint32_t* sizeScan, const uint32_t* subScan, uint32_t levelBase, uint32_t* oct, uint32_t* leafRef,
               const uint8_t* candCount, const uint32_t* candWords, const uint32_t* candScan, uint32_t* candList, uint32_t* valueRangeBits) {
    __shared__ HermiteTable tab;
    __shared__ float4 lattice[kWarpsPerCta][27][2];
    __shared__ float coeff[kWarpsPerCta][64];
    if (deepest) loadHermite(tab);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t node = blockIdx.x * kWarpsPerCta + warp;
    if (node >= lv.count) return;
    const float4 ch = lv.centerHalf[node];
    const bool leaf = deepest || lv.terminal[node];
    const uint32_t at = levelBase + sizeScan[node];
    if (lane < 16) lattice[warp][cornerLattice(lane >> 1)][lane & 1] = lv.values[size_t(node) * 16 + lane];
    if (leaf) {
        __syncwarp();
        if (deepest) {   // coefficients were not fitted by Iter 1 (:716-719)
            fitCoefficients(tab, lattice[warp], 2.0f * ch.w, coeff[warp], lane);
            __syncwarp();
            reinterpret_cast<float*>(oct)[at + lane] = coeff[warp][lane];
            reinterpret_cast<float*>(oct)[at + 32 + lane] = coeff[warp][32 + lane];
        } else {
            reinterpret_cast<float*>(oct)[at + lane] = coeffIn[size_t(node) * 64 + lane];
            reinterpret_cast<float*>(oct)[at + 32 + lane] = coeffIn[size_t(node) * 64 + 32 + lane];
        }
        float cornerAbs = 0.0f;
        if (lane < 8) cornerAbs = gabs(lattice[warp][cornerLattice(lane)][0].x);
        for (int o = 4; o > 0; o >>= 1) cornerAbs = fmaxf(cornerAbs, __shfl_down_sync(0xffffffffu, cornerAbs, o));
        if (lane == 0) {
            oct[lv.word[node]] = (at & kOctIndexMask) | kLeafBit;
            leafRef[(at - g.G * g.G * g.G) >> 3] = (depth << kRefShift) | node;
            if (!isnan(cornerAbs)) atomicMax(valueRangeBits, __float_as_uint(cornerAbs));
        }
        return;
    }
    if (lane < 19) {
        lattice[warp][cSampleLattice[lane]][0] = mids[size_t(node) * 38 + 2 * lane];
        lattice[warp][cSampleLattice[lane]][1] = mids[size_t(node) * 38 + 2 * lane + 1];
    }
    __syncwarp();
    if (real) {
        if (lane == 0) oct[lv.word[node]] = at & kOctIndexMask;
        if (lane < 8) oct[at + lane] = kFreshChild;
    }
    const uint32_t base = subScan[node] * 8u;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int e = lane + 32 * r, c = e >> 4, k = (e >> 1) & 7, h = e & 1;
        const int L = ((c & 1) + (k & 1)) + 3 * (((c >> 1) & 1) + ((k >> 1) & 1)) + 9 * ((c >> 2) + (k >> 2));
        next.values[size_t(base) * 16 + e] = lattice[warp][L][h];
    }
    if (lane < 8) {
        const float h = 0.5f * ch.w;
        const f3 c = mk3(ch.x, ch.y, ch.z) + cornerDir(lane) * h;
        next.centerHalf[base + lane] = make_float4(c.x, c.y, c.z, h);
        const uint32_t pc = lv.coord[node];
        const uint32_t ix = ((pc & 1023u) << 1) | (lane & 1u), iy = (((pc >> 10) & 1023u) << 1) | ((lane >> 1) & 1u),
                       iz = (((pc >> 20) & 1023u) << 1) | (uint32_t(lane) >> 2);
        next.coord[base + lane] = ix | (iy << 10) | (iz << 20);
        next.word[base + lane] = real ? at + lane : ((depth + 1 == g.startDepth) ? startSlotOf(g, c) : kNone);
    }
    if (candCount) {
        const uint32_t n = candCount[node];
        if (lane < n) candList[candScan[node] + lane] = candWords[size_t(node) * 18 + lane];
    }
}

__global__ void widenCountsKernel(const uint8_t* in, uint32_t* out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// ---- fix-up pass (:741-1181) -----------------------------------------------------------------------------

__global__ void fixClaimKernel(const uint32_t* candList, uint32_t n, const uint32_t* oct, uint32_t G3, uint32_t* claim) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t slot = ((oct[candList[i]] & kOctIndexMask) - G3) >> 3;
    atomicMin(&claim[slot], i);
}
__global__ void fixRootFlagKernel(const uint32_t* candList, uint32_t n, const uint32_t* oct, uint32_t G3, const uint32_t* claim, uint32_t* isRoot) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t slot = ((oct[candList[i]] & kOctIndexMask) - G3) >> 3;
    isRoot[i] = claim[slot] == i ? 1u : 0u;
}

struct RoundArrays {
    uint32_t count;
    uint32_t* rootIdx;
    uint8_t* depth;
    uint8_t* childId;
    uint32_t* parent;       // index in the previous round
    uint32_t* coord;
    float4* centerHalf;
    float4* values;         // 16 per node
    uint32_t* split;        // 0 / 1
    uint32_t* nSamples;     // true samples the node needs (0 unless it splits)
    uint32_t* interpMask;
    uint32_t* size;
    uint32_t* word;
    uint32_t* block;
};

// round 0: one node per re-opened leaf, copied from the store that holds it. 16 threads per root.
__global__ void fixRootInitKernel(const uint32_t* candList, const uint32_t* isRoot, const uint32_t* rootPos, uint32_t n, const uint32_t* oct,
                                  uint32_t G3, uint32_t* claim, const uint32_t* leafRef, const StoreView* stores, RoundArrays r0,
                                  uint32_t* oldCoef) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = t >> 4, part = t & 15u;
    if (i >= n || !isRoot[i]) return;
    const uint32_t w = candList[i];
    const uint32_t coefIdx = oct[w] & kOctIndexMask;
    const uint32_t slot = (coefIdx - G3) >> 3;
    const uint32_t ref = leafRef[slot];
    const StoreView st = stores[ref >> kRefShift];
    const uint32_t src = ref & kRefIndexMask;
    const uint32_t k = rootPos[i];
    r0.values[size_t(k) * 16 + part] = st.values[size_t(src) * 16 + part];
    if (part == 0) {
        r0.rootIdx[k] = k;
        r0.depth[k] = st.depth ? st.depth[src] : uint8_t(st.uniformDepth);
        r0.childId[k] = 0;
        r0.parent[k] = kNone;
        r0.coord[k] = st.coord[src];
        r0.centerHalf[k] = st.centerHalf[src];
        r0.word[k] = w;
        oldCoef[k] = coefIdx;
        claim[slot] = kNone;   // leave the claim table clean for the next pass
    }
}

// split decision: a node at depth nd <= current depth splits iff one of its 18 same-depth neighbours is a
// regular (unmarked) inner node (:794-911). One warp per node.
__global__ void __launch_bounds__(kWarpsPerCta * 32)
fixProbeKernel(Grid g, RoundArrays rd, uint32_t round, uint32_t currentDepth, const uint32_t* oct, unsigned long long* firstLeaf) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t node = blockIdx.x * kWarpsPerCta + warp;
    if (node >= rd.count) return;
    const uint32_t nd = rd.depth[node];
    uint32_t mine = 0;
    if (nd <= currentDepth && lane < 18) {
        const uint32_t pc = rd.coord[node];
        const int ix = int(pc & 1023u), iy = int((pc >> 10) & 1023u), iz = int(pc >> 20);
        const int entry = cEntry[lane];
        const uint32_t dir = uint32_t(entry >> 2) + 1u, sign = uint32_t(entry & 3);
        const int qx = ix + ((dir & 1u) ? axisSign(dir, sign, 0) : 0), qy = iy + ((dir & 2u) ? axisSign(dir, sign, 1) : 0),
                  qz = iz + ((dir & 4u) ? axisSign(dir, sign, 2) : 0);
        uint32_t w;
        if (probeCell(oct, g, nd, qx, qy, qz, true, w) == 2) mine = cFaceMask[entry];
    }
    const uint32_t subdivided = __reduce_or_sync(0xffffffffu, mine);
    if (lane == 0) {
        rd.split[node] = subdivided ? 1u : 0u;
        rd.nSamples[node] = uint32_t(__popc(subdivided & 0x7FFFFu));   // true samples: the points on faces shared with subdivided neighbours
        rd.interpMask[node] = ~subdivided;
        if (!subdivided) atomicMin(&firstLeaf[rd.rootIdx[node]], (static_cast<unsigned long long>(round) << 32) | node);
    }
}

// positions of the true samples of the splitting nodes, in (node, sample) order
__global__ void fixSamplePointsKernel(RoundArrays rd, const uint32_t* sampleScan, float4* points) {
    const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= rd.count || !rd.nSamples[node]) return;
    const float4 ch = rd.centerHalf[node];
    const uint32_t need = ~rd.interpMask[node] & 0x7FFFFu;
    uint32_t at = sampleScan[node];
    for (int s = 0; s < 19; s++)
        if (need & (1u << (18 - s))) {
            const int L = cSampleLattice[s];
            const f3 rel = mk3(float(L % 3 - 1), float((L / 3) % 3 - 1), float(L / 9 - 1));
            const f3 p = mk3(ch.x, ch.y, ch.z) + rel * ch.w;
            points[at++] = make_float4(p.x, p.y, p.z, 0.f);
        }
}

// values of a splitting fix-up node (:913-939) and its 8 children (:955-1142). One warp per node.
__global__ void __launch_bounds__(kWarpsPerCta * 32)
fixValuesKernel(RoundArrays rd, RoundArrays next, const uint32_t* splitScan, const uint32_t* sampleScan, const float4* samples,
                float sqThreshold) {
    __shared__ HermiteTable tab;
    __shared__ float4 lattice[kWarpsPerCta][27][2];
    __shared__ float coeff[kWarpsPerCta][64];
    loadHermite(tab);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t node = blockIdx.x * kWarpsPerCta + warp;
    if (node >= rd.count || !rd.split[node]) return;
    const float4 ch = rd.centerHalf[node];
    const float nodeSize = 2.0f * ch.w;
    if (lane < 16) lattice[warp][cornerLattice(lane >> 1)][lane & 1] = rd.values[size_t(node) * 16 + lane];
    __syncwarp();
    fitCoefficients(tab, lattice[warp], nodeSize, coeff[warp], lane);
    __syncwarp();
    const uint32_t interpMask = rd.interpMask[node];
    if (lane < 19) {
        const int L = cSampleLattice[lane];
        const float x = 0.5f * float(L % 3), y = 0.5f * float((L / 3) % 3), z = 0.5f * float(L / 9);
        float4 lo, hi = make_float4(0.f, 0.f, 0.f, 0.f);
        bool interpolate = (interpMask & (1u << (18 - lane))) != 0;
        if (!interpolate) {   // on a face shared with a subdivided neighbour: the true sample, unless the polynomial is close enough
            const uint32_t need = ~interpMask & 0x7FFFFu;
            lo = samples[sampleScan[node] + __popc(need >> (19 - lane))];   // rank among the node's true samples (lane 0: shift by 19 = none)
            const float d = lo.x - polyValueExact(coeff[warp], x, y, z);
            interpolate = d * d < sqThreshold;
        }
        if (interpolate) vertexValues(coeff[warp], x, y, z, nodeSize, lo, hi);
        lattice[warp][L][0] = lo;
        lattice[warp][L][1] = hi;
    }
    __syncwarp();
    const uint32_t base = splitScan[node] * 8u;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int e = lane + 32 * r, c = e >> 4, k = (e >> 1) & 7, h = e & 1;
        const int L = ((c & 1) + (k & 1)) + 3 * (((c >> 1) & 1) + ((k >> 1) & 1)) + 9 * ((c >> 2) + (k >> 2));
        next.values[size_t(base) * 16 + e] = lattice[warp][L][h];
    }
    if (lane < 8) {
        const float h = 0.5f * ch.w;
        const f3 c = mk3(ch.x, ch.y, ch.z) + cornerDir(lane) * h;
        next.centerHalf[base + lane] = make_float4(c.x, c.y, c.z, h);
        const uint32_t pc = rd.coord[node];
        const uint32_t ix = ((pc & 1023u) << 1) | (lane & 1u), iy = (((pc >> 10) & 1023u) << 1) | ((lane >> 1) & 1u),
                       iz = (((pc >> 20) & 1023u) << 1) | (uint32_t(lane) >> 2);
        next.coord[base + lane] = ix | (iy << 10) | (iz << 20);
        next.rootIdx[base + lane] = rd.rootIdx[node];
        next.depth[base + lane] = uint8_t(rd.depth[node] + 1);
        next.childId[base + lane] = uint8_t(lane);
        next.parent[base + lane] = node;
    }
}

// words each fix-up node appends: 8 (split), 64 (leaf), 0 for the first leaf of its root in breadth-first
// order, which recycles the re-opened leaf's coefficient block (:1146-1157)
__global__ void fixSizeKernel(RoundArrays rd, uint32_t round, uint32_t rounds, uint32_t K, const unsigned long long* firstLeaf,
                              uint32_t* byRoot, uint32_t* byRound) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rd.count) return;
    const uint32_t k = rd.rootIdx[i];
    uint32_t s = 8u;
    if (!rd.split[i]) s = (firstLeaf[k] == ((static_cast<unsigned long long>(round) << 32) | i)) ? 0u : 64u;
    rd.size[i] = s;
    if (s) {
        atomicAdd(&byRoot[size_t(k) * rounds + round], s);
        atomicAdd(&byRound[size_t(round) * K + k], s);
    }
}

struct PoolArrays {
    float4* centerHalf;
    uint32_t* coord;
    uint8_t* depth;
    uint32_t* word;
    float4* values;
};

// final position of every fix-up node, its word, and for leaves the coefficient block + pool record.
// One warp per node; rounds are launched in order (a child's word is its parent's block + child id).
__global__ void __launch_bounds__(kWarpsPerCta * 32)
fixWriteKernel(RoundArrays rd, const uint32_t* prevBlock, uint32_t round, uint32_t rounds, uint32_t K, const uint32_t* sizeScan,
               const uint32_t* splitScan, const uint32_t* byRootScan, const uint32_t* byRoundScan, const uint32_t* oldCoef, uint32_t fixBase,
               uint32_t G3, uint32_t* oct, uint32_t* leafRef, PoolArrays pool, uint32_t poolBase, uint32_t poolStore, uint32_t* splitWords) {
    __shared__ HermiteTable tab;
    __shared__ float4 lattice[kWarpsPerCta][27][2];
    __shared__ float coeff[kWarpsPerCta][64];
    loadHermite(tab);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t node = blockIdx.x * kWarpsPerCta + warp;
    if (node >= rd.count) return;
    const uint32_t k = rd.rootIdx[node];
    const bool split = rd.split[node] != 0;
    const uint32_t size = rd.size[node];
    const uint32_t segStart = byRoundScan[size_t(round) * K + k] - byRoundScan[size_t(round) * K];
    const uint32_t block = (size == 0) ? oldCoef[k] : fixBase + byRootScan[size_t(k) * rounds + round] + (sizeScan[node] - segStart);
    const uint32_t word = (round == 0) ? rd.word[node] : prevBlock[rd.parent[node]] + rd.childId[node];
    if (lane == 0) {
        rd.block[node] = block;
        rd.word[node] = word;
    }
    if (split) {
        if (lane == 0) {
            oct[word] = (block & kOctIndexMask) | kMark;   // setValues(false, childIndex); markNode() (:948-949)
            splitWords[splitScan[node]] = word;
        }
        return;
    }
    const float4 ch = rd.centerHalf[node];
    if (lane < 16) lattice[warp][cornerLattice(lane >> 1)][lane & 1] = rd.values[size_t(node) * 16 + lane];
    __syncwarp();
    fitCoefficients(tab, lattice[warp], 2.0f * ch.w, coeff[warp], lane);
    __syncwarp();
    reinterpret_cast<float*>(oct)[block + lane] = coeff[warp][lane];
    reinterpret_cast<float*>(oct)[block + 32 + lane] = coeff[warp][32 + lane];
    const uint32_t p = poolBase + (node - splitScan[node]);   // rank among this round's leaves
    if (lane < 16) pool.values[size_t(p) * 16 + lane] = rd.values[size_t(node) * 16 + lane];
    if (lane == 0) {
        oct[word] = (block & kOctIndexMask) | kLeafBit;
        leafRef[(block - G3) >> 3] = (poolStore << kRefShift) | p;
        pool.centerHalf[p] = ch;
        pool.coord[p] = rd.coord[node];
        pool.depth[p] = rd.depth[node];
        pool.word[p] = word;
    }
}

__global__ void unmarkKernel(const uint32_t* words, uint32_t n, uint32_t* oct) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) oct[words[i]] &= ~kMark;
}

__device__ __forceinline__ uint32_t orderedFloat(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// computeMinBorderValue (src/sdf/OctreeSdf.cpp:155-230) over the nodes of one store that are leaves of the
// final tree: polynomial value at the leaf corners lying on the border of the unit cube. 8 threads per node.
__global__ void minBorderKernel(StoreView st, const uint32_t* oct, uint32_t* minBorderOrdered) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t i = t >> 3, corner = t & 7u;
    if (i >= st.count) return;
    const uint32_t word = st.word[i];
    if (word == kNone) return;
    const uint32_t v = oct[word];
    if (!(v & kLeafBit)) return;
    const uint32_t depth = st.depth ? st.depth[i] : st.uniformDepth;
    const uint32_t pc = st.coord[i];
    const uint32_t res = 1u << depth;
    const uint32_t ix = pc & 1023u, iy = (pc >> 10) & 1023u, iz = pc >> 20;
    if (!(ix == 0 || iy == 0 || iz == 0 || ix == res - 1 || iy == res - 1 || iz == res - 1)) return;
    const float half = 0.5f / float(res);
    const f3 pos = mk3((float(ix) + 0.5f) / float(res), (float(iy) + 0.5f) / float(res), (float(iz) + 0.5f) / float(res));
    const f3 sp = pos + half * cornerDir(corner);
    if (double(sp.x) < 1e-4 || double(sp.y) < 1e-4 || double(sp.z) < 1e-4 || double(sp.x) > double(1.0f) - 1e-4 ||
        double(sp.y) > double(1.0f) - 1e-4 || double(sp.z) > double(1.0f) - 1e-4) {
        const float* c = reinterpret_cast<const float*>(oct) + (v & kOctIndexMask);
        const float val = polyValueExact(c, float(corner & 1u), float((corner >> 1) & 1u), float(corner >> 2));
        if (!isnan(val)) atomicMin(minBorderOrdered, orderedFloat(val));
    }
}

__global__ void fillKernel(uint32_t* p, uint32_t value, uint64_t n) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = value;
}

// ---- host orchestration ------------------------------------------------------------------------------------

struct NodeLevel {
    uint32_t count = 0;
    DevBuf<float4> centerHalf, values;
    DevBuf<uint32_t> coord, word;
    DevBuf<uint8_t> terminal;
    void alloc(uint32_t n) {
        count = n;
        centerHalf.alloc(n); values.alloc(size_t(n) * 16); coord.alloc(n); word.alloc(n); terminal.alloc(n);
    }
    LevelArrays arrays() { return LevelArrays{count, centerHalf.p, coord.p, word.p, values.p, terminal.p}; }
    StoreView store(uint32_t depth) const { return StoreView{centerHalf.p, coord.p, nullptr, word.p, values.p, depth, count}; }
};

struct FixRound {
    uint32_t count = 0;
    DevBuf<uint32_t> rootIdx, parent, coord, split, nSamples, sampleScan, interpMask, size, word, block, sizeScan, splitScan;
    DevBuf<uint8_t> depth, childId;
    DevBuf<float4> centerHalf, values;
    uint32_t nSplit = 0;
    void alloc(uint32_t n) {
        count = n;
        rootIdx.alloc(n); parent.alloc(n); coord.alloc(n); split.alloc(n); nSamples.alloc(n); sampleScan.alloc(n); interpMask.alloc(n); size.alloc(n); word.alloc(n); block.alloc(n);
        sizeScan.alloc(n); splitScan.alloc(n); depth.alloc(n); childId.alloc(n); centerHalf.alloc(n); values.alloc(size_t(n) * 16);
    }
    RoundArrays arrays() {
        return RoundArrays{count, rootIdx.p, depth.p, childId.p, parent.p, coord.p, centerHalf.p, values.p, split.p, nSamples.p, interpMask.p, size.p, word.p, block.p};
    }
};

struct LeafPool {
    uint32_t count = 0;
    DevBuf<float4> centerHalf, values;
    DevBuf<uint32_t> coord, word;
    DevBuf<uint8_t> depth;
    void alloc(uint32_t n) { count = n; centerHalf.alloc(n); values.alloc(size_t(n) * 16); coord.alloc(n); word.alloc(n); depth.alloc(n); }
    PoolArrays arrays() { return PoolArrays{centerHalf.p, coord.p, depth.p, word.p, values.p}; }
    StoreView store() const { return StoreView{centerHalf.p, coord.p, depth.p, word.p, values.p, 0u, count}; }
};

// the octree array and its two side tables (one entry per 8 words above the start grid), grown by doubling
struct GrowingOctree {
    DevBuf<uint32_t> oct, leafRef, claim;
    uint64_t capacity = 0, G3 = 0;
    // `st`: the stream whose work uses the arrays at this point. On a side stream the old blocks may only go back to
    // the block cache (where the legacy-stream code would pick them up) once the copies have completed.
    void reserve(uint64_t words, cudaStream_t st = nullptr) {
        if (words <= capacity) return;
        if (words > uint64_t(kOctIndexMask)) throw Error(SDFB200_ERR_INVALID, "octree exceeds the 30-bit index space of OctreeNode");
        uint64_t cap = std::max<uint64_t>(capacity * 2, words + (words >> 2) + 4096);
        cap = std::min<uint64_t>(cap, uint64_t(kOctIndexMask) + 1);
        DevBuf<uint32_t> o(cap), r((cap - G3) / 8 + 2), c((cap - G3) / 8 + 2);
        const uint64_t oldSide = capacity ? (capacity - G3) / 8 + 2 : 0;
        if (capacity) {
            SDFB_CUDA(cudaMemcpyAsync(o.p, oct.p, capacity * 4, cudaMemcpyDeviceToDevice, st));
            SDFB_CUDA(cudaMemcpyAsync(r.p, leafRef.p, oldSide * 4, cudaMemcpyDeviceToDevice, st));
        }
        fillKernel<<<divUp(c.n, 256), 256, 0, st>>>(c.p, kNone, c.n);   // the claim table is clean between passes
        if (st) SDFB_CUDA(cudaStreamSynchronize(st));
        oct = std::move(o); leafRef = std::move(r); claim = std::move(c);
        capacity = cap;
    }
};

}  // namespace

void buildOctreeContinuityOnDevice(sdfb200_sdf& out, const PreparedMesh& mesh, const float* box6, uint32_t depth, uint32_t startDepth,
                                   int rule, float param0, float param1, const SampleExchange& exchange) {
    const auto tStart = std::chrono::steady_clock::now();
    out.stats = sdfb200_build_stats{};
    sdfb200_build_stats& st = out.stats;
    if (depth > 10) throw Error(SDFB200_ERR_INVALID, "octree depth > 10 is not supported (node coordinates are packed in 3x10 bits)");
    if (startDepth > depth) throw Error(SDFB200_ERR_INVALID, "startDepth must not exceed depth");
    out.format = SDFB200_FORMAT_OCTREE;
    out.maxDepth = depth;
    out.slotWords = 1;
    cubifyBox(out, box6, startDepth);
    SDFB_CUDA(cudaGetDevice(&out.device));
    DeviceCacheSettle settle(out.device);
    uploadHermite();
    {   // face/edge sample table (:139-176), derived: sample s is on the shared face iff rel[a] == side for every axis of dir
        static const int lat[19] = {1, 3, 4, 5, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 19, 21, 22, 23, 25};
        uint32_t table[24];
        for (uint32_t dir = 1; dir <= 6; dir++)
            for (uint32_t sign = 0; sign < 4; sign++) {
                uint32_t m = 0;
                const uint32_t nAxes = (dir & 1) + ((dir >> 1) & 1) + ((dir >> 2) & 1);
                if (sign < (1u << nAxes))
                    for (int s = 0; s < 19; s++) {
                        const int rel[3] = {lat[s] % 3 - 1, (lat[s] / 3) % 3 - 1, lat[s] / 9 - 1};
                        bool on = true;
                        int bit = 0;
                        for (int a = 0; a < 3; a++)
                            if (dir & (1u << a)) {
                                if (rel[a] != (((sign >> bit) & 1u) ? 1 : -1)) on = false;
                                bit++;
                            }
                        if (on) m |= 1u << (18 - s);
                    }
                table[4 * (dir - 1) + sign] = m;
            }
        SDFB_CUDA(cudaMemcpyToSymbol(cFaceMask, table, sizeof(table)));
    }

    if (!mesh.hasBvh) throw Error(SDFB200_ERR_INVALID, "OctreeSdf needs a mesh prepared with its BVH (SDFB200_MESH_BVH)");
    meshStats(mesh, st);
    const DeviceMesh dmesh = mesh.dev.view();

    NvtxRange nvtx("sdfb200:continuity:levels");
    auto t0 = std::chrono::steady_clock::now();
    const uint32_t G = uint32_t(out.startGridSize), G3 = G * G * G;
    Grid grid{G, startDepth, {out.boxMin[0], out.boxMin[1], out.boxMin[2]}, out.cellSize};
    const float sqThreshold = param0 * param0;
    GrowingOctree oc;
    oc.G3 = G3;
    oc.reserve(std::max<uint64_t>(uint64_t(G3) + 4096, uint64_t(32) << 20));   // 128 MB up front: growth (alloc + copy of three arrays) is the slow path
    SDFB_CUDA(cudaMemsetAsync(oc.oct.p, 0, size_t(G3) * 4));
    uint64_t words = G3;   // append cursor of mOctreeData

    std::vector<std::unique_ptr<NodeLevel>> levels(depth + 2);
    std::vector<std::unique_ptr<LeafPool>> pools;
    std::vector<StoreView> storeTable(32, StoreView{nullptr, nullptr, nullptr, nullptr, nullptr, 0u, 0u});
    DevBuf<StoreView> dStores(32);
    std::vector<std::unique_ptr<DevBuf<uint32_t>>> splitWordLists;
    std::vector<uint32_t> splitWordCounts;
    DevBuf<uint32_t> scalars(2);
    const uint32_t scalarInit[2] = {0u, 0xFFFFFFFFu};
    scalars.upload(scalarInit, 2);

    const uint32_t d0 = std::min(startDepth, 1u);
    // The levels above the start depth always subdivide: their nodes are known before any sample is, so their Iter-1 samples and
    // those of the start level are taken in ONE batch, with the seed corners on a side stream next to it (octree_build.cu has the
    // measurements: each of these launches lasts as long as its longest far-field traversal). Same positions: same bits.
    static const bool batchVirtual = [] { const char* e = std::getenv("SDFB200_BATCH_VIRTUAL_LEVELS"); return !(e && e[0] == '0'); }();
    const bool batched = batchVirtual && startDepth > d0;
    struct SeedStream {
        cudaStream_t s = nullptr; cudaEvent_t done = nullptr;
        ~SeedStream() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } if (done) cudaEventDestroy(done); }
    } seed;
    {
        const float boxSize = out.boxMax[0] - out.boxMin[0];
        const float h0 = float(0.5f * boxSize * std::pow(0.5f, d0));
        const f3 boxMin = mk3(out.boxMin[0], out.boxMin[1], out.boxMin[2]);
        const f3 c0 = boxMin + mk3(h0, h0, h0);
        const uint32_t per = 1u << d0;
        std::vector<float4> ch;
        std::vector<uint32_t> coord;
        for (uint32_t k = 0; k < per; k++)
            for (uint32_t j = 0; j < per; j++)
                for (uint32_t i = 0; i < per; i++) {
                    const f3 c = c0 + (mk3(float(i), float(j), float(k)) * 2.0f) * h0;
                    ch.push_back(make_float4(c.x, c.y, c.z, h0));
                    coord.push_back(i | (j << 10) | (k << 20));
                }
        levels[d0].reset(new NodeLevel());
        NodeLevel& L = *levels[d0];
        L.alloc(uint32_t(ch.size()));
        L.centerHalf.upload(ch.data(), ch.size());
        L.coord.upload(coord.data(), coord.size());
        SDFB_CUDA(cudaStreamSynchronize(0));   // ch / coord are stack vectors
        if (batched) {   // seeds next to the batch of the virtual levels (below), not before it
            SDFB_CUDA(cudaStreamCreateWithFlags(&seed.s, cudaStreamNonBlocking));
            SDFB_CUDA(cudaEventCreateWithFlags(&seed.done, cudaEventDisableTiming));
        }
        contSeedKernel<<<divUp(L.count * 8, 64), 64, bvhStackBytes(dmesh, 64), seed.s>>>(dmesh, grid, L.arrays(), d0);
        if (batched) SDFB_CUDA(cudaEventRecord(seed.done, seed.s));
        st.kernel_launches++;
        st.samples_evaluated += L.count * 8;
    }

    static const bool timing = std::getenv("SDFB200_TIMING") != nullptr;
    auto tPhase = std::chrono::steady_clock::now();
    auto tick = [&](const char* what, uint32_t d) {
        if (!timing) return;
        cudaDeviceSynchronize();
        std::fprintf(stderr, "[sdfb200] continuity depth %u %-10s %8.2f ms\n", d, what, msSince(tPhase));
        tPhase = std::chrono::steady_clock::now();
    };
    DevBuf<float4> mids, fixPoints;
    LevelSampler levelSampler;
    levelSampler.exchange = exchange;
    LevelSampler fixSampler;   // own result buffers: the fix-up samples of a round are read while the level's are still live
    fixSampler.exchange = exchange;
    DevBuf<float> coeffs;
    DevBuf<uint32_t> sizes, sub, sizeScan, subScan, candCount32, candScan, candWords, candList, isRoot, rootPos;
    DevBuf<uint8_t> candCount;
    Scanner scanner;
    // ---- fix-up pass of depth d: re-open the queued leaves (:741-1181). On one rank it is deferred until the Iter-1
    // sampling of depth d + 1 has been launched and then runs on a side stream underneath it: its rounds are small,
    // latency-bound launches (a far-field BVH traversal chain is ~1.8 ms) that would otherwise leave the GPU idle, and
    // depth d + 1 needs its results only for the decision / junction kernels, which are issued after this returns.
    cudaStream_t fs = nullptr;
    Scanner fixScanner;
    auto runFixup = [&](uint32_t d, uint32_t nCand) {
        setDeviceBlockStream(fs);
        fixSampler.stream = fs;
        {
            SDFB_CUDA(cudaMemcpyAsync(dStores.p, storeTable.data(), sizeof(StoreView) * 32, cudaMemcpyHostToDevice, fs));
            isRoot.ensure(nCand); rootPos.ensure(nCand);
            fixClaimKernel<<<divUp(nCand, 256), 256, 0, fs>>>(candList.p, nCand, oc.oct.p, G3, oc.claim.p);
            fixRootFlagKernel<<<divUp(nCand, 256), 256, 0, fs>>>(candList.p, nCand, oc.oct.p, G3, oc.claim.p, isRoot.p);
            const uint32_t K = fixScanner.run(isRoot.p, rootPos.p, nCand, false, fs);
            std::vector<std::unique_ptr<FixRound>> rounds;
            rounds.emplace_back(new FixRound());
            rounds[0]->alloc(K);
            DevBuf<uint32_t> oldCoef(K);
            DevBuf<unsigned long long> firstLeaf(K);
            SDFB_CUDA(cudaMemsetAsync(firstLeaf.p, 0xFF, size_t(K) * 8, fs));
            fixRootInitKernel<<<divUp(uint64_t(nCand) * 16, 256), 256, 0, fs>>>(candList.p, isRoot.p, rootPos.p, nCand, oc.oct.p, G3, oc.claim.p, oc.leafRef.p,
                                                                          dStores.p, rounds[0]->arrays(), oldCoef.p);
            st.kernel_launches += 6;
            for (uint32_t r = 0;; r++) {
                FixRound& R = *rounds[r];
                const uint32_t g8 = divUp(R.count, kWarpsPerCta);
                fixProbeKernel<<<g8, kWarpsPerCta * 32, 0, fs>>>(grid, R.arrays(), r, d, oc.oct.p, firstLeaf.p);
                R.nSplit = fixScanner.run(R.split.p, R.splitScan.p, R.count, false, fs);
                st.kernel_launches += 4;
                if (R.nSplit == 0) break;
                rounds.emplace_back(new FixRound());
                FixRound& Nx = *rounds[r + 1];
                Nx.alloc(R.nSplit * 8);
                const uint32_t nTrue = fixScanner.run(R.nSamples.p, R.sampleScan.p, R.count, false, fs);
                fixPoints.ensure(std::max<uint32_t>(nTrue, 1));
                fixSamplePointsKernel<<<divUp(R.count, 128), 128, 0, fs>>>(R.arrays(), R.sampleScan.p, fixPoints.p);
                const float4* fixSamples = fixSampler.runPoints(dmesh, fixPoints.p, nTrue);
                fixValuesKernel<<<g8, kWarpsPerCta * 32, 0, fs>>>(R.arrays(), Nx.arrays(), R.splitScan.p, R.sampleScan.p, fixSamples, sqThreshold);
                st.kernel_launches += 6;
                st.samples_evaluated += nTrue;
                st.nodes_processed += R.count;
            }
            tick("fix rounds", d);
            const uint32_t nRounds = uint32_t(rounds.size());
            DevBuf<uint32_t> byRoot(size_t(K) * nRounds + 1), byRound(size_t(K) * nRounds + 1);
            SDFB_CUDA(cudaMemsetAsync(byRoot.p, 0, byRoot.n * 4, fs));
            SDFB_CUDA(cudaMemsetAsync(byRound.p, 0, byRound.n * 4, fs));
            uint32_t nLeaves = 0, nSplits = 0;
            for (uint32_t r = 0; r < nRounds; r++) {
                FixRound& R = *rounds[r];
                fixSizeKernel<<<divUp(R.count, 256), 256, 0, fs>>>(R.arrays(), r, nRounds, K, firstLeaf.p, byRoot.p, byRound.p);
                fixScanner.run(R.size.p, R.sizeScan.p, R.count, false, fs);
                nLeaves += R.count - R.nSplit;
                nSplits += R.nSplit;
                st.kernel_launches += 4;
            }
            const uint32_t fixWords = fixScanner.run(byRoot.p, byRoot.p, size_t(K) * nRounds, false, fs);
            fixScanner.run(byRound.p, byRound.p, size_t(K) * nRounds, false, fs);
            oc.reserve(words + fixWords, fs);
            pools.emplace_back(new LeafPool());
            LeafPool& P = *pools.back();
            P.alloc(nLeaves);
            if (nLeaves > kRefIndexMask) throw Error(SDFB200_ERR_INVALID, "more than 2^27 leaves in one fix-up pass");
            splitWordLists.emplace_back(new DevBuf<uint32_t>(std::max<uint32_t>(nSplits, 1)));
            splitWordCounts.push_back(nSplits);
            uint32_t poolBase = 0, splitBase = 0;
            for (uint32_t r = 0; r < nRounds; r++) {
                FixRound& R = *rounds[r];
                fixWriteKernel<<<divUp(R.count, kWarpsPerCta), kWarpsPerCta * 32, 0, fs>>>(
                    R.arrays(), r ? rounds[r - 1]->block.p : nullptr, r, nRounds, K, R.sizeScan.p, R.splitScan.p, byRoot.p, byRound.p, oldCoef.p,
                    uint32_t(words), G3, oc.oct.p, oc.leafRef.p, P.arrays(), poolBase, kPoolStoreBase + d, splitWordLists.back()->p + splitBase);
                poolBase += R.count - R.nSplit;
                splitBase += R.nSplit;
                st.kernel_launches++;
            }
            storeTable[kPoolStoreBase + d] = P.store();
            words += fixWords;
            tick("fix write", d);
            st.nodes_processed += rounds.back()->count;
            if (fs) SDFB_CUDA(cudaStreamSynchronize(fs));   // the blocks of this pass go back to the cache only when its work is done
        }
        setDeviceBlockStream(nullptr);
    };
    static const bool overlapOff = std::getenv("SDFB200_CONT_NO_OVERLAP") != nullptr;   // A/B switch (profiles/)
    const bool overlap = exchange.world <= 1 && !timing && !overlapOff;
    cudaEvent_t emitted = nullptr;
    struct StreamGuard {   // the side stream and its event live as long as this build
        cudaStream_t& s; cudaEvent_t& e;
        ~StreamGuard() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } if (e) cudaEventDestroy(e); setDeviceBlockStream(nullptr); }
    } streamGuard{fs, emitted};
    if (overlap) {
        SDFB_CUDA(cudaStreamCreateWithFlags(&fs, cudaStreamNonBlocking));
        SDFB_CUDA(cudaEventCreateWithFlags(&emitted, cudaEventDisableTiming));
    }
    uint32_t pendingDepth = 0, pendingCand = 0;
    DevBuf<float4> preMids, preCentres;
    DevBuf<float> nodeCost;
    DevBuf<uint32_t> specCoord;
    static const bool speculativeLevel = [] { const char* e = std::getenv("SDFB200_SPECULATIVE_LEVEL"); return !(e && e[0] == '0'); }();
    bool speculate = false;
    std::vector<uint32_t> preOffset(depth + 2, 0u);
    if (batched) {
        uint32_t total = 0;
        for (uint32_t d = d0; d <= startDepth; d++) {
            NodeLevel& L = *levels[d];
            preOffset[d] = total;
            total += L.count;
            if (d == startDepth) break;
            levels[d + 1].reset(new NodeLevel());
            NodeLevel& N = *levels[d + 1];
            N.alloc(L.count * 8);
            contChildGeometryKernel<<<divUp(uint64_t(L.count) * 8, 256), 256>>>(L.count, L.centerHalf.p, L.coord.p, N.centerHalf.p, N.coord.p);
            st.kernel_launches++;
        }
        // ... and, speculatively, of all eight children of every start node (octree_build.cu)
        speculate = speculativeLevel && exchange.world <= 1 && startDepth + 1 < depth;
        const uint32_t nStart = levels[startDepth]->count;
        preOffset[startDepth + 1] = total;
        preCentres.alloc(total + (speculate ? nStart * 8 : 0));
        for (uint32_t d = d0; d <= startDepth; d++)
            SDFB_CUDA(cudaMemcpyAsync(preCentres.p + preOffset[d], levels[d]->centerHalf.p, size_t(levels[d]->count) * sizeof(float4), cudaMemcpyDeviceToDevice));
        if (speculate) {
            specCoord.alloc(size_t(nStart) * 8);
            contChildGeometryKernel<<<divUp(uint64_t(nStart) * 8, 256), 256>>>(nStart, levels[startDepth]->centerHalf.p, levels[startDepth]->coord.p, preCentres.p + total, specCoord.p);
            total += nStart * 8;
        }
        preMids.alloc(size_t(total) * 38);
        const uint32_t ran = levelSampler.run(dmesh, preCentres.p, total, preMids.p, 2);
        st.leaves += ran == 0xFFFFFFFFu ? uint64_t(total) * 19 : ran;
        SDFB_CUDA(cudaStreamWaitEvent(0, seed.done, 0));   // corner values of the seeds, before the first children inherit them
    }
    for (uint32_t d = d0; d <= depth; d++) {
        NodeLevel& L = *levels[d];
        const bool real = d >= startDepth, deepest = d == depth;
        if (L.count > kRefIndexMask) throw Error(SDFB200_ERR_INVALID, "more than 2^27 nodes in one level");
        const bool presampled = batched && d <= startDepth && !deepest;
        const bool childrenExist = batched && d < startDepth;   // geometry made for the batch
        if (!childrenExist) levels[d + 1].reset(new NodeLevel());
        if (L.count == 0) {
            if (pendingCand) { runFixup(pendingDepth, pendingCand); pendingCand = 0; }
            continue;
        }
        storeTable[d] = L.store(d);
        const uint32_t grid8 = divUp(L.count, kWarpsPerCta);
        // ---- Iter 1
        float4* midsPtr = presampled ? preMids.p + size_t(preOffset[d]) * 38 : mids.p;
        if (!deepest) {
            const bool gathered = speculate && d == startDepth + 1;
            if (gathered) {   // sub / subScan still hold the start level's Iter 2
                mids.ensure(size_t(L.count) * 38);
                midsPtr = mids.p;
                const uint32_t parents = levels[startDepth]->count;
                contGatherSpeculativeKernel<<<divUp(uint64_t(parents) * 8 * 38, 256), 256>>>(sub.p, subScan.p, parents, preMids.p + size_t(preOffset[d]) * 38, mids.p);
            } else if (!presampled) {
                mids.ensure(size_t(L.count) * 38);
                midsPtr = mids.p;
                nodeCost.ensure(L.count);   // start order of the traversals: nodes far from the surface first (LevelSampler::run)
                nodeCostKernel<<<divUp(L.count, 256), 256>>>(L.values.p, 16, 2, L.count, 1.0f / (out.boxMax[0] - out.boxMin[0]), nodeCost.p);
                const uint32_t ran = levelSampler.run(dmesh, L.centerHalf.p, L.count, mids.p, 2, nodeCost.p);
                st.leaves += ran == 0xFFFFFFFFu ? uint64_t(L.count) * 19 : ran;   // stats.leaves: BVH traversals run
            }
            if (pendingCand) { runFixup(pendingDepth, pendingCand); pendingCand = 0; }   // the previous depth's fix-up, under this sampling
            if (real) {
                coeffs.ensure(size_t(L.count) * 64);
                contDecideKernel<<<grid8, kWarpsPerCta * 32>>>(L.arrays(), midsPtr, coeffs.p, oc.oct.p, rule, sqThreshold, param1);
            } else {
                SDFB_CUDA(cudaMemsetAsync(L.terminal.p, 0, L.count));
            }
            st.kernel_launches += 2;
            st.samples_evaluated += uint64_t(L.count) * 19;
        }
        if (pendingCand) { runFixup(pendingDepth, pendingCand); pendingCand = 0; }   // deepest level: nothing to hide it under
        tick("iter1", d);
        // ---- Iter 2: T-junction samples
        uint32_t nCand = 0;
        const bool junctions = real && !deepest;
        if (junctions) {
            candCount.ensure(L.count);
            candWords.ensure(size_t(L.count) * 18);
            contJunctionKernel<<<grid8, kWarpsPerCta * 32>>>(grid, L.arrays(), d, midsPtr, coeffs.p, oc.oct.p, sqThreshold, candCount.p, candWords.p);
            candCount32.ensure(L.count);
            candScan.ensure(L.count);
            widenCountsKernel<<<divUp(L.count, 256), 256>>>(candCount.p, candCount32.p, L.count);
            nCand = scanner.run(candCount32.p, candScan.p, L.count);
            candList.ensure(std::max<uint32_t>(nCand, 1));
            st.kernel_launches += 5;
        }
        // ---- Iter 2: layout of the level, words, leaves, children
        sizes.ensure(L.count); sub.ensure(L.count); sizeScan.ensure(L.count); subScan.ensure(L.count);
        contSizesKernel<<<divUp(L.count, 256), 256>>>(L.arrays(), real, deepest, sizes.p, sub.p);
        const uint32_t levelWords = scanner.run(sizes.p, sizeScan.p, L.count);
        const uint32_t nSub = scanner.run(sub.p, subScan.p, L.count);
        oc.reserve(words + levelWords);
        NodeLevel& N = *levels[d + 1];
        if (!childrenExist) N.alloc(nSub * 8);
        contEmitKernel<<<grid8, kWarpsPerCta * 32>>>(grid, L.arrays(), N.arrays(), d, real, deepest, midsPtr, coeffs.p, sizeScan.p, subScan.p,
                                                     uint32_t(words), oc.oct.p, oc.leafRef.p, junctions ? candCount.p : nullptr, candWords.p,
                                                     candScan.p, candList.p, scalars.p);
        st.kernel_launches += 8;
        st.nodes_processed += L.count;
        words += levelWords;
        tick("iter2", d);
        if (nCand == 0) continue;

        if (nCand == 0) continue;
        if (overlap && !deepest) {   // run it under the next depth's sampling
            SDFB_CUDA(cudaEventRecord(emitted, cudaStream_t(0)));
            SDFB_CUDA(cudaStreamWaitEvent(fs, emitted, 0));
            pendingDepth = d;
            pendingCand = nCand;
        } else runFixup(d, nCand);
    }
    if (pendingCand) runFixup(pendingDepth, pendingCand);
    // final un-mark (:1191-1217), border minimum over the leaves of the final tree
    for (size_t i = 0; i < splitWordLists.size(); i++)
        if (splitWordCounts[i]) unmarkKernel<<<divUp(splitWordCounts[i], 256), 256>>>(splitWordLists[i]->p, splitWordCounts[i], oc.oct.p);
    for (const StoreView& sv : storeTable)
        if (sv.count) { minBorderKernel<<<divUp(uint64_t(sv.count) * 8, 256), 256>>>(sv, oc.oct.p, scalars.p + 1); st.kernel_launches++; }
    scalars.download(out.shardScalars, 2);
    SDFB_CUDA(cudaDeviceSynchronize());
    st.levels_ms = msSince(t0);
    finalizeOctreeScalars(out);

    nvtx.next("sdfb200:continuity:finish");
    t0 = std::chrono::steady_clock::now();
    levels.clear(); pools.clear();
    // keep exactly `words` entries on the device for the query kernels
    auto finishStep = [&](const char* what) { if (timing) std::fprintf(stderr, "[sdfb200] continuity finish %-14s %8.2f ms\n", what, msSince(t0)); };
    finishStep("free levels");
    out.dOctree.alloc(words);
    SDFB_CUDA(cudaMemcpyAsync(out.dOctree.p, oc.oct.p, words * 4, cudaMemcpyDeviceToDevice));
    finishStep("device copy");
    prepareOctreeQuery(out);
    out.nOctree = words;
    out.hostMirror = false;
    if (exchange.world == 1) ensureHostMirror(out);   // the ranks of a collective build fetch their mirror when a getter asks
    finishStep("download");
    st.download_ms = msSince(t0);
    out.isShard = false;
    out.plan = RootPlan();
    out.stats.total_ms = msSince(tStart);
}

}  // namespace sdfb200
