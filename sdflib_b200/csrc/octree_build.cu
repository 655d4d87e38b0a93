// OctreeSdf construction on the GPU (hot path 1, tri-cubic variant).
//
// Replaces, behind sdfb200_build_octree, the reference's depth-first builder
//   OctreeSdf::initOctree<VHQueries<TriCubicInterpolation>>   src/sdf/OctreeSdfDepthFirst.h:31-527
//   VHQueries::calculateVerticesInfo                           include/SdfLib/TrianglesInfluence.h:952-996
//   tmd::TriangleMeshDistance::_query                          libs/InteractiveComputerGraphics/.../TriangleMeshDistance.h:492-540
//   TriCubicInterpolation::calculateCoefficients / interpolateValue  include/SdfLib/InterpolationMethods.h:292-439
//   estimateErrorFunctionIntegralByTrapezoidRule               include/SdfLib/OctreeSdfUtils.h:60-85
//   OctreeSdf::computeMinBorderValue                           src/sdf/OctreeSdf.cpp:155-230
//
// B200 design (not a translation of the CPU stack machine):
//   * level-synchronous: one launch processes every candidate node of a depth; a warp owns a node,
//     lanes 0..18 own its 19 sample points (BVH descent in float64 + signed distance/gradient in
//     float32), then the whole warp fits the 64 Hermite coefficients and integrates the error;
//   * subdividing nodes are compacted with an exclusive scan, children are written as 8 coalesced
//     records (centre, 8 corner value quadruples) — no per-node heap, no stack;
//   * the reference's array ORDER (a by-product of its stack discipline) is rebuilt afterwards from
//     subtree sizes: bottom-up size pass, top-down offset pass, one emit pass that writes node words
//     and recomputes leaf coefficients straight into their final position.
// Arithmetic is bit-faithful to the CPU build (this TU is compiled with -fmad=false); the only
// deliberate difference is that the reference's history-dependent 32^3 vertex cache is not emulated
// (DESIGN.md "parity"), i.e. every sample asks the BVH afresh.
#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <memory>

#include "device_utils.cuh"
#include "octree_device.cuh"
#include "sdf_internal.h"

namespace sdfb200 {

namespace {

struct LevelView {
    uint32_t count;
    const float4* centerHalf;   // xyz centre, w half size
    const float4* corners;      // 8 per node: (f, gx, gy, gz)
    const uint32_t* coord;      // ix | iy << 10 | iz << 20 at this depth
};

// ---- kernels ---------------------------------------------------------------------------------------

// Corner samples of the seed nodes (calculateVerticesInfo<8>, OctreeSdfDepthFirst.h:113-135)
__global__ void seedCornersKernel(DeviceMesh mesh, const float4* centerHalf, float4* corners, uint32_t nSeeds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nSeeds * 8) return;
    const float4 ch = centerHalf[i >> 3];
    const f3 p = mk3(ch.x, ch.y, ch.z) + cornerDir(i & 7u) * ch.w;
    corners[i] = samplePoint(mesh, p);
}

// One warp per candidate node of a level below maxDepth: fit, error integral, decision. The 19 samples were
// taken by sampleLatticeKernel (one thread per sample, octree_device.cuh).
__global__ void __launch_bounds__(kWarpsPerCta * 32)
levelDecideKernel(LevelView lv, const float4* mids, uint32_t* subdivide, int rule, float sqThreshold, float decay) {
    __shared__ HermiteTable tab;
    __shared__ float4 lattice[kWarpsPerCta][27];
    __shared__ float coeff[kWarpsPerCta][64];
    __shared__ float terms[kWarpsPerCta][19];
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&cHermite);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&tab);
        for (uint32_t i = threadIdx.x; i < sizeof(HermiteTable) / 4; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t node = blockIdx.x * kWarpsPerCta + warp;
    if (node >= lv.count) return;
    const float4 ch = lv.centerHalf[node];
    if (lane < 8) lattice[warp][2 * (lane & 1) + 6 * ((lane >> 1) & 1) + 18 * (lane >> 2)] = lv.corners[size_t(node) * 8 + lane];
    if (lane < 19) lattice[warp][cSampleLattice[lane]] = mids[size_t(node) * 19 + lane];
    __syncwarp();
    const float nodeSize = 2.0f * ch.w;
    coeff[warp][tab.order[lane]] = hermiteRow(tab, tab.order[lane], lattice[warp], nodeSize);
    coeff[warp][tab.order[32 + lane]] = hermiteRow(tab, tab.order[32 + lane], lattice[warp], nodeSize);
    __syncwarp();
    if (lane < 19) {
        const int L = cSampleLattice[lane];
        const int lx = L % 3, ly = (L / 3) % 3, lz = L / 9;
        const float v = polyValueExact(coeff[warp], 0.5f * float(lx), 0.5f * float(ly), 0.5f * float(lz));
        const int centred = (lx == 1) + (ly == 1) + (lz == 1);
        const float w = errorWeight(rule, centred);
        const float truth = lattice[warp][L].x;
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
            for (int s = 1; s < 19; s++) value += terms[warp][s];   // left-to-right, as the reference expression
        }
        subdivide[node] = (value < sqThreshold) ? 0u : 1u;
    }
}

__global__ void fillOnesKernel(uint32_t* p, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 1u;
}

// Geometry of the children of a level whose nodes all subdivide (the virtual levels above the start depth): centres, half
// sizes and packed coordinates, exactly as emitChildrenKernel computes them — they do not depend on any sample.
__global__ void childGeometryKernel(LevelView lv, float4* nextCenterHalf, uint32_t* nextCoord) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= lv.count * 8u) return;
    const uint32_t node = e >> 3, lane = e & 7u;
    const float4 ch = lv.centerHalf[node];
    const float h = 0.5f * ch.w;
    const f3 c = mk3(ch.x, ch.y, ch.z) + cornerDir(lane) * h;
    nextCenterHalf[e] = make_float4(c.x, c.y, c.z, h);
    const uint32_t pc = lv.coord[node];
    const uint32_t ix = ((pc & 1023u) << 1) | (lane & 1u), iy = (((pc >> 10) & 1023u) << 1) | ((lane >> 1) & 1u), iz = (((pc >> 20) & 1023u) << 1) | (lane >> 2);
    nextCoord[e] = ix | (iy << 10) | (iz << 20);
}

// speculative samples of ALL children of a level -> the samples of the children that exist (childOf: first child or kNoChild)
__global__ void gatherSpeculativeKernel(const uint32_t* __restrict__ childOf, uint32_t parents, const float4* __restrict__ spec, int perNode, float4* __restrict__ mids) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t perParent = uint64_t(8) * perNode;
    if (i >= uint64_t(parents) * perParent) return;
    const uint32_t p = uint32_t(i / perParent), r = uint32_t(i % perParent);
    const uint32_t base = childOf[p];
    if (base != kNoChild) mids[size_t(base) * perNode + r] = spec[i];
}

// Children of the subdividing nodes: 8 records each, corner values inherited from the 27-point lattice
// (child c, corner k  <-  lattice point (c+k) per axis; OctreeSdfDepthFirst.h:225-336).
__global__ void __launch_bounds__(kWarpsPerCta * 32)
emitChildrenKernel(LevelView lv, const float4* mids, const uint32_t* subdivide, const uint32_t* scan, uint32_t* childOf,
                   float4* nextCenterHalf, float4* nextCorners, uint32_t* nextCoord) {
    __shared__ float4 lattice[kWarpsPerCta][27];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t node = blockIdx.x * kWarpsPerCta + warp;
    if (node >= lv.count) return;
    if (!subdivide[node]) { if (lane == 0) childOf[node] = kNoChild; return; }
    const uint32_t base = scan[node] * 8u;
    if (lane == 0) childOf[node] = base;
    if (lane < 8) lattice[warp][2 * (lane & 1) + 6 * ((lane >> 1) & 1) + 18 * (lane >> 2)] = lv.corners[size_t(node) * 8 + lane];
    if (lane < 19) lattice[warp][cSampleLattice[lane]] = mids[size_t(node) * 19 + lane];
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int e = lane + 32 * r, c = e >> 3, k = e & 7;
        const int L = ((c & 1) + (k & 1)) + 3 * (((c >> 1) & 1) + ((k >> 1) & 1)) + 9 * ((c >> 2) + (k >> 2));
        nextCorners[size_t(base) * 8 + e] = lattice[warp][L];
    }
    if (lane < 8) {
        const float4 ch = lv.centerHalf[node];
        const float h = 0.5f * ch.w;
        const f3 c = mk3(ch.x, ch.y, ch.z) + cornerDir(lane) * h;
        nextCenterHalf[base + lane] = make_float4(c.x, c.y, c.z, h);
        const uint32_t pc = lv.coord[node];
        const uint32_t ix = ((pc & 1023u) << 1) | (lane & 1u), iy = (((pc >> 10) & 1023u) << 1) | ((lane >> 1) & 1u),
                       iz = (((pc >> 20) & 1023u) << 1) | (uint32_t(lane) >> 2);
        nextCoord[base + lane] = ix | (iy << 10) | (iz << 20);
    }
}

// ---- layout: subtree sizes bottom-up, block offsets top-down ---------------------------------------
__global__ void leafSizesKernel(uint32_t* words, uint32_t* childOf, uint32_t n) {   // deepest level: every node is a leaf
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { words[i] = 64u; childOf[i] = kNoChild; }
}
__global__ void subtreeSizesKernel(const uint32_t* childOf, const uint32_t* nextWords, uint32_t* words, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = childOf[i];
    uint32_t w = 64u;
    if (c != kNoChild) {
        w = 8u;
        for (int k = 0; k < 8; k++) w += nextWords[c + k];
    }
    words[i] = w;
}
// children are laid out in the order the reference's stack pops them: 7 first
__global__ void childOffsetsKernel(const uint32_t* childOf, const uint32_t* block, const uint32_t* nextWords,
                                   uint32_t* nextSlot, uint32_t* nextBlock, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = childOf[i];
    if (c == kNoChild) return;
    const uint32_t b = block[i];
    if (b == kNoChild) {   // subtree not owned by this rank
        for (int k = 0; k < 8; k++) nextBlock[c + k] = kNoChild;
        return;
    }
    uint32_t running = b + 8u;
    for (int k = 7; k >= 0; k--) {
        nextSlot[c + k] = b + uint32_t(k);
        nextBlock[c + k] = running;
        running += nextWords[c + k];
    }
}

__device__ __forceinline__ uint32_t orderedFloat(float f) {   // monotone float -> uint mapping for atomicMin
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Final pass: node words, leaf coefficient blocks, value range and border minimum.
// `out` is addressed relative to outBase (a shard writes only its own words).
__global__ void __launch_bounds__(kWarpsPerCta * 32)
emitWordsKernel(LevelView lv, const uint32_t* childOf, const uint32_t* slot, const uint32_t* block, uint32_t* out,
                uint32_t depth, uint32_t* valueRangeBits, uint32_t* minBorderOrdered) {
    __shared__ HermiteTable tab;
    __shared__ float4 lattice[kWarpsPerCta][27];
    __shared__ float coeff[kWarpsPerCta][64];
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&cHermite);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&tab);
        for (uint32_t i = threadIdx.x; i < sizeof(HermiteTable) / 4; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t node = blockIdx.x * kWarpsPerCta + warp;
    if (node >= lv.count) return;
    const uint32_t b = block[node];
    if (b == kNoChild) return;   // not owned
    const uint32_t c = childOf[node];
    if (c != kNoChild) { if (lane == 0) out[slot[node]] = b & kOctIndexMask; return; }
    if (lane == 0) out[slot[node]] = (b & kOctIndexMask) | kLeafBit;
    const float4 ch = lv.centerHalf[node];
    float cornerAbs = 0.0f;
    if (lane < 8) {
        const float4 v = lv.corners[size_t(node) * 8 + lane];
        lattice[warp][2 * (lane & 1) + 6 * ((lane >> 1) & 1) + 18 * (lane >> 2)] = v;
        cornerAbs = gabs(v.x);
    }
    __syncwarp();
    const float nodeSize = 2.0f * ch.w;
    const int r0 = tab.order[lane], r1 = tab.order[32 + lane];
    const float c0 = hermiteRow(tab, r0, lattice[warp], nodeSize), c1 = hermiteRow(tab, r1, lattice[warp], nodeSize);
    coeff[warp][r0] = c0;
    coeff[warp][r1] = c1;
    __syncwarp();
    reinterpret_cast<float*>(out)[b + lane] = coeff[warp][lane];
    reinterpret_cast<float*>(out)[b + 32 + lane] = coeff[warp][32 + lane];
    // mValueRange = max |corner distance| over leaves (OctreeSdfDepthFirst.h:358-361)
    for (int o = 4; o > 0; o >>= 1) cornerAbs = fmaxf(cornerAbs, __shfl_down_sync(0xffffffffu, cornerAbs, o));
    if (lane == 0 && !isnan(cornerAbs)) atomicMax(valueRangeBits, __float_as_uint(cornerAbs));
    // computeMinBorderValue: polynomial at the leaf corners that lie on the unit-cube border
    const uint32_t pc = lv.coord[node];
    const uint32_t res = 1u << depth;
    const uint32_t ix = pc & 1023u, iy = (pc >> 10) & 1023u, iz = pc >> 20;
    const bool touches = ix == 0 || iy == 0 || iz == 0 || ix == res - 1 || iy == res - 1 || iz == res - 1;
    if (touches && lane < 8) {
        const float half = 0.5f / float(res);
        const f3 pos = mk3((float(ix) + 0.5f) / float(res), (float(iy) + 0.5f) / float(res), (float(iz) + 0.5f) / float(res));
        const f3 sp = pos + half * cornerDir(lane);
        if (double(sp.x) < 1e-4 || double(sp.y) < 1e-4 || double(sp.z) < 1e-4 || double(sp.x) > double(1.0f) - 1e-4 ||
            double(sp.y) > double(1.0f) - 1e-4 || double(sp.z) > double(1.0f) - 1e-4) {
            const float v = polyValueExact(coeff[warp], float(lane & 1), float((lane >> 1) & 1), float(lane >> 2));
            if (!isnan(v)) atomicMin(minBorderOrdered, orderedFloat(v));
        }
    }
}

__global__ void nearestKernel(DeviceMesh mesh, const f3* pts, uint64_t n, uint32_t* out) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = bvhNearest(mesh, pts[i]);
}

__global__ void pointTriangleKernel(const TriData* tri, const f3* w, const f3* pts, uint64_t n, int mode, float* outDist, f3* outGrad) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f3 g = mk3(0.f, 0.f, 0.f);
    float d;
    if (mode == 0) d = signedDistPointTriangle(pts[i], *tri);
    else if (mode == 1) d = signedDistGradMesh(pts[i], *tri, w[0], w[1], w[2], g);
    else if (mode == 2) d = signedDistGradSelf(pts[i], *tri, g);
    else d = sqDistPointTriangle(pts[i], frameOf(*tri));
    outDist[i] = d;
    if (outGrad) outGrad[i] = g;
}

// ---- host orchestration --------------------------------------------------------------------------

struct Level {
    uint32_t count = 0;
    DevBuf<float4> centerHalf, corners;
    DevBuf<uint32_t> coord, childOf, words, slot, block;
    void alloc(uint32_t n) {
        count = n;
        centerHalf.alloc(n); corners.alloc(size_t(n) * 8); coord.alloc(n); childOf.alloc(n); words.alloc(n); slot.alloc(n); block.alloc(n);
    }
    LevelView view() const { return LevelView{count, centerHalf.p, corners.p, coord.p}; }
};


}  // namespace

void cubifyBox(sdfb200_sdf& s, const float* box6, uint32_t startDepth) {   // OctreeSdf.cpp:43-51
    const f3 mn = mk3(box6[0], box6[1], box6[2]), mx = mk3(box6[3], box6[4], box6[5]);
    const f3 size = mx - mn;
    const float maxSize = gmax(gmax(size.x, size.y), size.z);
    const f3 center = mn + 0.5f * size;
    const f3 lo = center - mk3(0.5f * maxSize, 0.5f * maxSize, 0.5f * maxSize);
    const f3 hi = center + mk3(0.5f * maxSize, 0.5f * maxSize, 0.5f * maxSize);
    s.boxMin[0] = lo.x; s.boxMin[1] = lo.y; s.boxMin[2] = lo.z;
    s.boxMax[0] = hi.x; s.boxMax[1] = hi.y; s.boxMax[2] = hi.z;
    s.startGridSize = 1 << startDepth;
    s.cellSize = maxSize / float(s.startGridSize);
}

RootPlan makeRootPlan(const float4* rootCH, const uint32_t* rootCoord, uint32_t G, uint32_t startDepth, const float* boxMin3,
                      float cellSize, uint32_t numThreads, uint32_t rank, uint32_t world, const uint32_t* weight) {
    RootPlan plan;
    const uint32_t G3 = G * G * G;
    plan.G3 = G3; plan.world = world; plan.rank = rank;
    plan.rootSlot.resize(G3); plan.order.resize(G3); plan.owned.assign(G3, 0); plan.ownerOf.resize(G3);
    const f3 boxMin = mk3(boxMin3[0], boxMin3[1], boxMin3[2]);
    std::vector<uint64_t> key(G3);
    for (uint32_t r = 0; r < G3; r++) {
        const f3 f = (mk3(rootCH[r].x, rootCH[r].y, rootCH[r].z) - boxMin) / cellSize;   // OctreeSdfDepthFirst.h:408-409
        const int x = int(std::floor(f.x)), y = int(std::floor(f.y)), z = int(std::floor(f.z));
        plan.rootSlot[r] = uint32_t(z * int(G * G) + y * int(G) + x);
        if (numThreads >= 2) key[r] = plan.rootSlot[r];   // sub-octrees concatenated in start-grid order (:471-503)
        else {                                            // one global stack, virtual levels popped 7-first (:397-416)
            const uint32_t ix = rootCoord[r] & 1023u, iy = (rootCoord[r] >> 10) & 1023u, iz = rootCoord[r] >> 20;
            uint64_t k = 0;
            for (int b = int(startDepth) - 1; b >= 0; b--) {
                const uint32_t c = ((ix >> b) & 1u) | (((iy >> b) & 1u) << 1) | (((iz >> b) & 1u) << 2);
                k = (k << 3) | (7u - c);
            }
            key[r] = k;
        }
        plan.order[r] = r;
    }
    std::sort(plan.order.begin(), plan.order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
    // Owners by estimated work: longest-processing-time greedy over the roots' weights — heaviest root
    // first (ties in layout order), each to the rank with the least load so far (ties to the lowest rank). Every rank
    // computes the same weights from the replicated levels above the start depth, hence the same plan.
    std::vector<uint32_t> byWeight(plan.order);
    if (weight) std::stable_sort(byWeight.begin(), byWeight.end(), [&](uint32_t a, uint32_t b) { return weight[a] > weight[b]; });
    std::vector<uint64_t> load(world, 0);
    for (uint32_t i = 0; i < G3; i++) {
        const uint32_t r = byWeight[i];
        uint32_t q = 0;
        if (weight) { for (uint32_t k = 1; k < world; k++) if (load[k] < load[q]) q = k; }
        else q = i % world;
        load[q] += weight ? uint64_t(weight[r]) + 1 : 1;
        plan.ownerOf[r] = q;
        plan.owned[r] = q == rank;
    }
    return plan;
}

namespace {

__global__ void maskUnownedKernel(const uint8_t* owned, uint32_t* subdivide, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !owned[i]) subdivide[i] = 0u;
}

struct OctreeBuildState : BuildState {
    std::vector<std::unique_ptr<Level>> levels;
    uint32_t depth = 0, startDepth = 0;
    uint32_t numStreams() const override { return 1; }

    // phase 1: candidate levels top-down (only the subtrees of the own roots below the start depth), subtree sizes
    void buildLevels(sdfb200_sdf& out, const PreparedMesh& mesh, int rule, float param0, float param1, uint32_t numThreads,
                     uint32_t rank, uint32_t world) {
        sdfb200_build_stats& st = out.stats;
        if (!mesh.hasBvh) throw Error(SDFB200_ERR_INVALID, "OctreeSdf needs a mesh prepared with its BVH (SDFB200_MESH_BVH)");
        meshStats(mesh, st);
        const DeviceMesh dmesh = mesh.dev.view();

        NvtxRange nvtx("sdfb200:octree:levels");
        auto t0 = std::chrono::steady_clock::now();
        const uint32_t d0 = std::min(startDepth, 1u);
        const f3 boxMin = mk3(out.boxMin[0], out.boxMin[1], out.boxMin[2]);
        const float boxSize = out.boxMax[0] - out.boxMin[0];
        levels.resize(depth + 1);
        {   // seeds at depth d0 (OctreeSdfDepthFirst.h:113-135); centres in the reference's float arithmetic
            const float h0 = float(0.5f * boxSize * std::pow(0.5f, d0));
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
            levels[d0].reset(new Level());
            Level& L = *levels[d0];
            L.alloc(uint32_t(ch.size()));
            L.centerHalf.upload(ch.data(), ch.size());
            L.coord.upload(coord.data(), coord.size());
            st.kernel_launches++;
            st.samples_evaluated += L.count * 8;
        }
        DevBuf<float4> mids;
        LevelSampler levelSampler;
        DevBuf<uint32_t> flags, scan;
        DevBuf<uint8_t> dOwned;
        Scanner scanner;
        // The levels above the start depth always subdivide (OctreeSdfDepthFirst.h:397-416), so their nodes are known before any
        // sample is: the mid-point samples of ALL of them and of the start level are taken in ONE batch, with the seed corners on a
        // side stream next to it. Each of these launches is a handful of CTAs that lasts as long as its longest far-field
        // traversal (2.5 - 4.7 ms each on the C2 mesh: seeds, depth 1, 2, 3 = 13.3 ms in a row, profiles/r2_summary.md); together
        // they last as long as the longest one. Same positions, same kernels per sample: same bits.
        static const bool batchVirtual = [] { const char* e = std::getenv("SDFB200_BATCH_VIRTUAL_LEVELS"); return !(e && e[0] == '0'); }();
        const bool batched = batchVirtual && startDepth > d0;
        struct SeedStream {
            cudaStream_t s = nullptr; cudaEvent_t ready = nullptr, done = nullptr;
            ~SeedStream() { if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); } if (ready) cudaEventDestroy(ready); if (done) cudaEventDestroy(done); }
        } seed;
        {
            Level& L = *levels[d0];
            if (batched) {
                SDFB_CUDA(cudaStreamCreateWithFlags(&seed.s, cudaStreamNonBlocking));
                SDFB_CUDA(cudaEventCreateWithFlags(&seed.ready, cudaEventDisableTiming));
                SDFB_CUDA(cudaEventCreateWithFlags(&seed.done, cudaEventDisableTiming));
                SDFB_CUDA(cudaEventRecord(seed.ready, 0));                     // the uploads above
                SDFB_CUDA(cudaStreamWaitEvent(seed.s, seed.ready, 0));
            }
            seedCornersKernel<<<divUp(L.count * 8, 64), 64, bvhStackBytes(dmesh, 64), seed.s>>>(dmesh, L.centerHalf.p, L.corners.p, L.count);
            if (batched) SDFB_CUDA(cudaEventRecord(seed.done, seed.s));
        }
        DevBuf<float4> preMids, preCentres;
        DevBuf<float> nodeCost;
        DevBuf<uint32_t> specCoord;
        static const bool speculativeLevel = [] { const char* e = std::getenv("SDFB200_SPECULATIVE_LEVEL"); return !(e && e[0] == '0'); }();
        bool speculate = false;
        std::vector<uint32_t> preOffset(depth + 2, 0u);
        if (batched) {
            uint32_t total = 0;
            for (uint32_t d = d0; d <= startDepth; d++) {
                Level& L = *levels[d];
                preOffset[d] = total;
                total += L.count;
                if (d == startDepth) break;
                levels[d + 1].reset(new Level());
                Level& N = *levels[d + 1];
                N.alloc(L.count * 8);
                childGeometryKernel<<<divUp(uint64_t(L.count) * 8, 256), 256>>>(L.view(), N.centerHalf.p, N.coord.p);
                st.kernel_launches++;
            }
            // ... and, speculatively, of all eight children of every start node: most start nodes subdivide (438 of 512 on the C2 mesh),
            // and the level below the start depth is still a few hundred CTAs that wait for their far-field traversals
            speculate = speculativeLevel && world == 1 && startDepth + 1 < depth;
            const uint32_t nStart = levels[startDepth]->count;
            preOffset[startDepth + 1] = total;
            preCentres.alloc(total + (speculate ? nStart * 8 : 0));
            for (uint32_t d = d0; d <= startDepth; d++)
                SDFB_CUDA(cudaMemcpyAsync(preCentres.p + preOffset[d], levels[d]->centerHalf.p, size_t(levels[d]->count) * sizeof(float4), cudaMemcpyDeviceToDevice));
            if (speculate) {
                specCoord.alloc(size_t(nStart) * 8);
                childGeometryKernel<<<divUp(uint64_t(nStart) * 8, 256), 256>>>(levels[startDepth]->view(), preCentres.p + total, specCoord.p);
                total += nStart * 8;
            }
            preMids.alloc(size_t(total) * 19);
            const uint32_t ran = levelSampler.run(dmesh, preCentres.p, total, preMids.p, 1);
            st.leaves += ran == 0xFFFFFFFFu ? uint64_t(total) * 19 : ran;
            SDFB_CUDA(cudaStreamWaitEvent(0, seed.done, 0));                    // corner values of the seeds, before the first children inherit them
        }
        for (uint32_t d = d0; d <= depth; d++) {
            Level& L = *levels[d];
            if (d == startDepth) makePlan(out, L, numThreads, rank, world);
            if (d == depth) break;
            if (L.count == 0) { levels[d + 1].reset(new Level()); continue; }
            const bool presampled = batched && d <= startDepth;
            const bool gathered = speculate && d == startDepth + 1;
            if (!presampled) mids.alloc(size_t(L.count) * 19);
            if (gathered) {
                const Level& P = *levels[startDepth];
                gatherSpeculativeKernel<<<divUp(uint64_t(P.count) * 8 * 19, 256), 256>>>(P.childOf.p, P.count, preMids.p + size_t(preOffset[d]) * 19, 19, mids.p);
                st.leaves += 0;
            }
            const float4* midsPtr = presampled ? preMids.p + size_t(preOffset[d]) * 19 : mids.p;
            flags.alloc(L.count);
            scan.alloc(L.count);
            const uint32_t grid = divUp(L.count, kWarpsPerCta);
            if (!presampled && !gathered) {
                nodeCost.ensure(L.count);   // start order of the traversals: nodes far from the surface first (LevelSampler::run)
                nodeCostKernel<<<divUp(L.count, 256), 256>>>(L.corners.p, 8, 1, L.count, 1.0f / boxSize, nodeCost.p);
                const uint32_t ran = levelSampler.run(dmesh, L.centerHalf.p, L.count, mids.p, 1, nodeCost.p);
                st.leaves += ran == 0xFFFFFFFFu ? uint64_t(L.count) * 19 : ran;   // stats.leaves: BVH traversals run
            }
            if (d >= startDepth)
                levelDecideKernel<<<grid, kWarpsPerCta * 32>>>(L.view(), midsPtr, flags.p, rule, param0 * param0, param1);
            else
                fillOnesKernel<<<divUp(L.count, 256), 256>>>(flags.p, L.count);   // virtual levels always subdivide
            st.kernel_launches++;
            if (d == startDepth && world > 1) {   // the roots of other ranks are not refined here
                dOwned.alloc(L.count);
                dOwned.upload(out.plan.owned.data(), L.count);
                maskUnownedKernel<<<divUp(L.count, 256), 256>>>(dOwned.p, flags.p, L.count);
            }
            const uint32_t nSubdivide = scanner.run(flags.p, scan.p, L.count);
            st.kernel_launches += 4;
            st.samples_evaluated += uint64_t(L.count) * 19;
            if (!(batched && d < startDepth)) {   // (the children of a virtual level exist already: their geometry was needed for the batch)
                levels[d + 1].reset(new Level());
                levels[d + 1]->alloc(nSubdivide * 8);
            }
            Level& N = *levels[d + 1];
            emitChildrenKernel<<<grid, kWarpsPerCta * 32>>>(L.view(), midsPtr, flags.p, scan.p, L.childOf.p, N.centerHalf.p,
                                                            N.corners.p, N.coord.p);
            st.kernel_launches++;
            st.nodes_processed += L.count;
        }
        SDFB_CUDA(cudaDeviceSynchronize());
        st.levels_ms = msSince(t0);

        // subtree sizes, bottom-up
        nvtx.next("sdfb200:octree:layout");
        t0 = std::chrono::steady_clock::now();
        {
            Level& D = *levels[depth];
            if (D.count) { leafSizesKernel<<<divUp(D.count, 256), 256>>>(D.words.p, D.childOf.p, D.count); st.kernel_launches++; }
            st.nodes_processed += D.count;
        }
        for (int d = int(depth) - 1; d >= int(startDepth); d--) {
            Level& L = *levels[size_t(d)];
            if (!L.count) continue;
            subtreeSizesKernel<<<divUp(L.count, 256), 256>>>(L.childOf.p, levels[size_t(d) + 1]->words.p, L.words.p, L.count);
            st.kernel_launches++;
        }
        Level& R = *levels[startDepth];
        const uint32_t G3 = out.plan.G3;
        std::vector<uint32_t> rootWords(G3);
        R.words.download(rootWords.data(), G3);
        SDFB_CUDA(cudaDeviceSynchronize());
        out.shardSizes.assign(G3, 0u);
        for (uint32_t r = 0; r < G3; r++)
            if (out.plan.owned[r]) out.shardSizes[out.plan.rootSlot[r]] = rootWords[r];
        st.layout_ms = msSince(t0);
    }

    void makePlan(sdfb200_sdf& out, Level& R, uint32_t numThreads, uint32_t rank, uint32_t world) {
        const uint32_t G = uint32_t(out.startGridSize), G3 = G * G * G;
        if (R.count != G3) throw Error(SDFB200_ERR_INVALID, "internal: start level is not a full grid");
        std::vector<float4> rootCH(G3);
        std::vector<uint32_t> rootCoord(G3);
        std::vector<float4> corners(size_t(G3) * 8);
        R.centerHalf.download(rootCH.data(), G3);
        R.coord.download(rootCoord.data(), G3);
        R.corners.download(corners.data(), corners.size());
        SDFB_CUDA(cudaDeviceSynchronize());
        // work estimate of a start voxel: it is refined down to the leaves where the surface passes through or next to it
        // (some corner closer than the voxel's diagonal), and ends after a few levels elsewhere
        std::vector<uint32_t> weight(G3);
        for (uint32_t r = 0; r < G3; r++) {
            float nearest = INFINITY;
            for (int c = 0; c < 8; c++) nearest = std::min(nearest, std::fabs(corners[size_t(r) * 8 + c].x));
            weight[r] = nearest < 3.4641f * rootCH[r].w ? 64u : 1u;   // 2 * half * sqrt(3)
        }
        out.plan = makeRootPlan(rootCH.data(), rootCoord.data(), G, startDepth, out.boxMin, out.cellSize, numThreads, rank, world,
                                world > 1 ? weight.data() : nullptr);
    }

    // phase 2: global offsets, node words + leaf blocks of the own roots at their final positions
    void finish(sdfb200_sdf& out, const uint32_t* allSizesBySlot) override {
        sdfb200_build_stats& st = out.stats;
        NvtxRange nvtx("sdfb200:octree:emit");
        auto t0 = std::chrono::steady_clock::now();
        const RootPlan& plan = out.plan;
        const uint32_t G3 = plan.G3;
        Level& R = *levels[startDepth];
        std::vector<uint32_t> rootBlock(G3);
        out.streams.assign(1, ShardStream());
        ShardStream& S = out.streams[0];
        S.elemBytes = 4; S.rootBase.resize(G3); S.rootSize.resize(G3);
        uint64_t running = G3;
        for (uint32_t i = 0; i < G3; i++) {
            const uint32_t r = plan.order[i];
            S.rootBase[r] = running;
            S.rootSize[r] = allSizesBySlot[plan.rootSlot[r]];
            rootBlock[r] = plan.owned[r] ? uint32_t(running) : kNoChild;
            running += S.rootSize[r];
        }
        if (running > uint64_t(kOctIndexMask)) throw Error(SDFB200_ERR_INVALID, "octree exceeds the 30-bit index space of OctreeNode");
        const uint64_t totalWords = running;
        R.slot.upload(plan.rootSlot.data(), G3);
        R.block.upload(rootBlock.data(), G3);
        for (uint32_t d = startDepth; d < depth; d++) {
            Level& L = *levels[d];
            Level& N = *levels[d + 1];
            if (!L.count || !N.count) continue;
            childOffsetsKernel<<<divUp(L.count, 256), 256>>>(L.childOf.p, L.block.p, N.words.p, N.slot.p, N.block.p, L.count);
            st.kernel_launches++;
        }
        out.dOctree.alloc(totalWords);
        out.nOctree = totalWords;
        out.hostMirror = false;
        S.dBase = reinterpret_cast<uint8_t*>(out.dOctree.p);
        SDFB_CUDA(cudaMemsetAsync(out.dOctree.p, 0, totalWords * sizeof(uint32_t)));
        DevBuf<uint32_t> scalars(2);
        const uint32_t scalarInit[2] = {0u, 0xFFFFFFFFu};
        scalars.upload(scalarInit, 2);
        for (uint32_t d = startDepth; d <= depth; d++) {
            Level& L = *levels[d];
            if (!L.count) continue;
            emitWordsKernel<<<divUp(L.count, kWarpsPerCta), kWarpsPerCta * 32>>>(L.view(), L.childOf.p, L.slot.p, L.block.p, out.dOctree.p, d,
                                                                                  scalars.p, scalars.p + 1);
            st.kernel_launches++;
        }
        scalars.download(out.shardScalars, 2);
        SDFB_CUDA(cudaDeviceSynchronize());
        st.layout_ms += msSince(t0);
        levels.clear();
        if (plan.world == 1) {
            finalizeOctreeScalars(out);
            nvtx.next("sdfb200:octree:query_index_and_download");
            t0 = std::chrono::steady_clock::now();
            prepareOctreeQuery(out);
            ensureHostMirror(out);
            st.download_ms = msSince(t0);
            out.isShard = false;
        }
    }
};

}  // namespace

void finalizeOctreeScalars(sdfb200_sdf& s) {
    float vr;
    std::memcpy(&vr, &s.shardScalars[0], 4);
    s.valueRange = vr;
    const uint32_t o = s.shardScalars[1];
    const uint32_t bits = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    float mb;
    std::memcpy(&mb, &bits, 4);
    s.minBorderValue = (o == 0xFFFFFFFFu) ? INFINITY : mb;
}

void buildOctreeOnDevice(sdfb200_sdf& out, const PreparedMesh& mesh, const float* box6, uint32_t depth, uint32_t startDepth,
                         int rule, float param0, float param1, uint32_t numThreads, uint32_t rank, uint32_t world) {
    const auto tStart = std::chrono::steady_clock::now();
    out.stats = sdfb200_build_stats{};
    if (depth > 10) throw Error(SDFB200_ERR_INVALID, "octree depth > 10 is not supported (node coordinates are packed in 3x10 bits)");
    if (startDepth > depth) throw Error(SDFB200_ERR_INVALID, "startDepth must not exceed depth");
    out.format = SDFB200_FORMAT_OCTREE;
    out.maxDepth = depth;
    out.slotWords = 1;
    cubifyBox(out, box6, startDepth);
    SDFB_CUDA(cudaGetDevice(&out.device));
    uploadHermite();
    DeviceCacheSettle settle(out.device);
    std::unique_ptr<OctreeBuildState> state(new OctreeBuildState());
    state->depth = depth;
    state->startDepth = startDepth;
    out.isShard = true;
    state->buildLevels(out, mesh, rule, param0, param1, numThreads, rank, world);
    if (world == 1) state->finish(out, out.shardSizes.data());
    else out.build = std::move(state);
    out.stats.total_ms = msSince(tStart);
}

void nearestTriangleOnDevice(const HostMesh& mesh, const float* xyz, uint64_t n, uint32_t* outTri) {
    const std::shared_ptr<PreparedMesh> pm = prepareMesh(mesh, true, false);
    const MeshOnDevice& dm = pm->dev;
    DevBuf<f3> pts(n);
    DevBuf<uint32_t> out(n);
    pts.upload(reinterpret_cast<const f3*>(xyz), n);
    if (n) nearestKernel<<<divUp(n, kBvhThreads), kBvhThreads, bvhStackBytes(dm.view())>>>(dm.view(), pts.p, n, out.p);
    out.download(outTri, n);
    SDFB_CUDA(cudaDeviceSynchronize());
}

void pointTriangleOnDevice(const float* tri37, const float* v123, const float* xyz, uint64_t n, int mode, float* outDist,
                           float* outGrad) {
    DevBuf<TriData> tri(1);
    DevBuf<f3> w(3), pts(n), grad(outGrad ? n : 0);
    DevBuf<float> dist(n);
    tri.upload(reinterpret_cast<const TriData*>(tri37), 1);
    const float zeros[9] = {0};
    w.upload(reinterpret_cast<const f3*>(v123 ? v123 : zeros), 3);
    pts.upload(reinterpret_cast<const f3*>(xyz), n);
    if (n) pointTriangleKernel<<<divUp(n, 128), 128>>>(tri.p, w.p, pts.p, n, mode, dist.p, outGrad ? grad.p : nullptr);
    dist.download(outDist, n);
    if (outGrad) grad.download(reinterpret_cast<f3*>(outGrad), n);
    SDFB_CUDA(cudaDeviceSynchronize());
}

}  // namespace sdfb200
