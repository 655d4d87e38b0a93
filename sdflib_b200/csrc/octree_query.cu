// Bulk OctreeSdf::getDistance / getDistance(p, grad) (hot path 2).
//
// Reference: src/sdf/OctreeSdf.cpp:93-152 (traversal), include/SdfLib/InterpolationMethods.h:432-455
// (64-term polynomial and its analytic gradient), include/SdfLib/utils/Mesh.h:42-63 (out-of-grid
// points fall back to the box distance + mMinBorderValue).
//
// This file is compiled twice (build.py):
//   default              -> launchOctreeQueryFast : FMA Horner evaluation (63 FMA for the value)
//   -DSDFB_QUERY_EXACT   -> launchOctreeQueryExact: the reference's literal operation order, no FMA
//                           (-fmad=false), bit-identical to the CPU reference
// In BOTH variants the cell selection ((p - min) / cell, floor, the fract(2f) descent) uses exactly
// the reference's IEEE operations, so both visit the same leaf; they differ only in how the leaf
// polynomial is summed (<= a few ulp of the largest term).
//
// Kernels (octree_query_kernels.cuh):
//   octreeQueryTileKernel  the default of the FMA path: persistent warps over 32-query tiles (next tile's points loaded
//                          ahead), dense top index instead of the first dependent gathers, division-free cell selection,
//                          quad-cooperative evaluation (see the comment on the kernel for what each of these removes)
//   octreeQueryKernel      one query per thread: the reference-order (bit-exact) object, and the fallback of the FMA
//                          object for arrays the tile kernel's preconditions exclude (leaf blocks not 16-byte aligned,
//                          block offsets not of the 8-word form, start grid not a power of two)
// Measured and dropped (profiles/r2_summary.md): a dense leaf index down to the deepest level for the one-query-per-
// thread kernel (+3 %: it was bound by the coefficient fill, not the descent) and a variant staging distinct leaves in
// shared memory (slower: the register fill from shared memory costs the same wavefronts).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "sdf_internal.h"

namespace sdfb200 {

namespace {
#include "octree_query_kernels.cuh"
}  // namespace

#ifndef SDFB_QUERY_EXACT
// Builds s.dTopIndex (topIndexKernel): levels = as deep as the tree goes below the start grid, capped so that the index
// stays within 2^21 cells (8 MB: L2-resident next to the structure). Leaves s.topLevels = -1 when the array does not
// have the 8-word block structure the packing relies on (a hand-made .bin) or the start grid is not a power of two;
// the one-query-per-thread kernel serves those. Runs on the legacy default stream and synchronises (build / load time).
void prepareOctreeQuery(sdfb200_sdf& s) {
    s.topLevels = -1;
    const char* plain = std::getenv("SDFB200_QUERY_PLAIN");   // read once per structure, not per query
    s.forcePlainQuery = plain && plain[0] == '1';
    const char* packed = std::getenv("SDFB200_QUERY_PACKED");   // A/B switch: 0 = scalar FFMA / FADD in the tile kernel
    s.packedQuery = !(packed && packed[0] == '0');
    if (s.format != SDFB200_FORMAT_OCTREE || !s.dOctree.p) return;
    int startDepth = 0;
    while ((1 << startDepth) < s.startGridSize) startDepth++;
    if ((1 << startDepth) != s.startGridSize || 3 * startDepth > 21) return;
    int levels = std::min(std::max(0, int(s.maxDepth) - startDepth), 15);
    while (levels > 0 && 3 * (startDepth + levels) > 21) levels--;
    const uint64_t cells = uint64_t(1) << (3 * (startDepth + levels));
    s.dTopIndex.alloc(cells + 1);
    uint32_t* bad = s.dTopIndex.p + cells;
    SDFB_CUDA(cudaMemsetAsync(bad, 0, sizeof(uint32_t)));
    topIndexKernel<<<uint32_t((cells + 255) / 256), 256>>>(s.dOctree.p, s.startGridSize, levels, s.dTopIndex.p, bad);
    SDFB_CUDA(cudaGetLastError());
    uint32_t hostBad = 1;
    SDFB_CUDA(cudaMemcpy(&hostBad, bad, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (hostBad) { s.dTopIndex.release(); return; }
    s.topLevels = levels;
    s.gridShift = startDepth;
}

namespace {
int smCount(int device) {
    static int cached[64] = {0};
    if (device < 0 || device >= 64) device = 0;
    if (!cached[device]) {
        int n = 0;
        SDFB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device));
        cached[device] = n > 0 ? n : 1;
    }
    return cached[device];
}
}  // namespace
#endif

#ifdef SDFB_QUERY_EXACT
void launchOctreeQueryExact(
#else
void launchOctreeQueryFast(
#endif
    const sdfb200_sdf& s, const float* dXyz, uint64_t n, float* dDist, float* dGrad, cudaStream_t st
#ifndef SDFB_QUERY_EXACT
    , bool hostMapped
#endif
    ) {
    if (n == 0) return;
    QueryParams q;
    q.minx = s.boxMin[0]; q.miny = s.boxMin[1]; q.minz = s.boxMin[2];
    q.maxx = s.boxMax[0]; q.maxy = s.boxMax[1]; q.maxz = s.boxMax[2];
    q.cell = s.cellSize;
    q.grid = s.startGridSize;
    q.minBorder = s.minBorderValue;
    const uint32_t grid = uint32_t((n + 255) / 256);
    // leaf blocks are 16-byte aligned iff the start grid has a multiple of 4 slots (all blocks are 8 or 64 words)
    const bool vec = (uint64_t(s.startGridSize) * s.startGridSize * s.startGridSize) % 4 == 0 && s.leafBlocksAligned;
    // 16 path bits: built structures are at most depth 10 (builders), loaded ones were measured by validateStructure
    if (s.relativeDepth > 16) throw Error(SDFB200_ERR_INVALID, "octree deeper than 16 levels below its start grid");
#ifndef SDFB_QUERY_EXACT
    // Markstein's division needs a normal divisor whose significand is not all ones; 1 / cell must be normal as well
    uint32_t cellBits;
    std::memcpy(&cellBits, &s.cellSize, 4);
    const bool cellOk = s.cellSize > 1e-30f && s.cellSize < 1e30f && (cellBits & 0x7FFFFFu) != 0x7FFFFFu;
    if (vec && s.topLevels >= 0 && cellOk && !s.forcePlainQuery && n < (uint64_t(1) << 36)) {
        TileQuery tq;
        const int L = s.topLevels;
        tq.cellL = std::ldexp(s.cellSize, -L);
        tq.rcellL = 1.0f / tq.cellL;
        const float limit = float(uint32_t(s.startGridSize) << L);
        std::memcpy(&tq.limitBits, &limit, 4);
        tq.shiftN = s.gridShift + L;
        tq.topLevels = L;
        tq.G3 = uint32_t(s.startGridSize) * uint32_t(s.startGridSize) * uint32_t(s.startGridSize);
        (void)hostMapped;
        const uint64_t tiles = (n + 31) / 32;
        const uint32_t ctas = uint32_t(std::min<uint64_t>((tiles + kTileWarps - 1) / kTileWarps, uint64_t(smCount(s.device)) * 6));   // persistent: 6 CTAs per SM
        // value only: the packed form (FFMA2 / FMUL2 / FADD2: 0.1618 -> 0.1576 ms on the 256^3 grid, same bits); with gradients it
        // spills at 40 registers and loses on random points (0.616 -> 0.657 ms), so that kernel stays scalar
        if (dGrad) octreeQueryTileKernel<true, false><<<ctas, kTileWarps * 32, 0, st>>>(s.dOctree.p, s.dTopIndex.p, q, tq, dXyz, n, dDist, dGrad);
        else if (s.packedQuery) octreeQueryTileKernel<false, true><<<ctas, kTileWarps * 32, 0, st>>>(s.dOctree.p, s.dTopIndex.p, q, tq, dXyz, n, dDist, nullptr);
        else octreeQueryTileKernel<false, false><<<ctas, kTileWarps * 32, 0, st>>>(s.dOctree.p, s.dTopIndex.p, q, tq, dXyz, n, dDist, nullptr);
        SDFB_CUDA(cudaGetLastError());
        return;
    }
#endif
    if (dGrad) {
        if (vec) octreeQueryKernel<true, true><<<grid, 256, 0, st>>>(s.dOctree.p, q, dXyz, n, dDist, dGrad);
        else octreeQueryKernel<true, false><<<grid, 256, 0, st>>>(s.dOctree.p, q, dXyz, n, dDist, dGrad);
    } else {
        if (vec) octreeQueryKernel<false, true><<<grid, 256, 0, st>>>(s.dOctree.p, q, dXyz, n, dDist, nullptr);
        else octreeQueryKernel<false, false><<<grid, 256, 0, st>>>(s.dOctree.p, q, dXyz, n, dDist, nullptr);
    }
    SDFB_CUDA(cudaGetLastError());
}


// Sphere tracing over an OCTREE structure (device pointers; see octreeTraceKernel)
#ifdef SDFB_QUERY_EXACT
void launchOctreeTraceExact(
#else
void launchOctreeTraceFast(
#endif
    const sdfb200_sdf& s, const float* dOrigin, const float* dDirection, uint64_t n, float epsilon, float farDistance, uint32_t maxIterations,
    float* dHit, float* dTravelled, uint32_t* dIterations, cudaStream_t st) {
    if (n == 0) return;
    QueryParams q;
    q.minx = s.boxMin[0]; q.miny = s.boxMin[1]; q.minz = s.boxMin[2];
    q.maxx = s.boxMax[0]; q.maxy = s.boxMax[1]; q.maxz = s.boxMax[2];
    q.cell = s.cellSize;
    q.grid = s.startGridSize;
    q.minBorder = s.minBorderValue;
    if (s.relativeDepth > 16) throw Error(SDFB200_ERR_INVALID, "octree deeper than 16 levels below its start grid");
    const bool vec = (uint64_t(s.startGridSize) * s.startGridSize * s.startGridSize) % 4 == 0 && s.leafBlocksAligned;
    const TraceParams tp{epsilon, farDistance, maxIterations};
    const uint32_t grid = uint32_t((n + 127) / 128);
    if (vec) octreeTraceKernel<true><<<grid, 128, 0, st>>>(s.dOctree.p, q, tp, dOrigin, dDirection, n, dHit, dTravelled, dIterations);
    else octreeTraceKernel<false><<<grid, 128, 0, st>>>(s.dOctree.p, q, tp, dOrigin, dDirection, n, dHit, dTravelled, dIterations);
    SDFB_CUDA(cudaGetLastError());
}

}  // namespace sdfb200
