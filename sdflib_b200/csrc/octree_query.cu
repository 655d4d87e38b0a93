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
// Kernel shape: one query per thread, 256-thread CTAs, grid sized by the batch. The descent is a
// chain of dependent 4-byte gathers (L2-resident after the first touch: the whole structure of the
// headline config is 82 MB < 126 MB L2); the leaf block is 64 consecutive floats.
// What bounds it (profiles/r1_summary.md): not HBM but the L1/LSU data pipe — every query must bring its
// own 64 coefficients (256 B) into registers: each of the 16 vector loads costs one wavefront per distinct leaf
// among the warp's lanes (6.2 on the 256^3 workload, tests/model_query_wavefronts.py: ~99 of ~115 wavefronts per
// warp); ncu shows l1tex__data_pipe_lsu_wavefronts at 77 % of peak with DRAM at 15 %. A warp-cooperative variant (distinct
// leaves staged once per warp in shared memory, evaluated from there) was measured and is slower
// (0.36 ms vs 0.26 ms on the 256^3 grid): it removes the tag-stage replays but keeps the same
// register-fill traffic and adds match/shuffle/shared-store work. It was dropped.
#include <algorithm>

#include "sdf_internal.h"

namespace sdfb200 {

namespace {
#include "octree_query_kernels.cuh"
}  // namespace

#ifndef SDFB_QUERY_EXACT
// Builds s.dLeafIndex (EXPERIMENTAL, see leafIndexKernel). levels = as deep as the tree goes, capped so that the index
// stays within 2^27 cells (512 MB); leaves s.leafIndexLevels = -1 when the array does not have the 8-word block
// structure the packing relies on (a hand-made .bin), in which case the plain kernel keeps being used.
void buildLeafIndex(sdfb200_sdf& s, cudaStream_t st) {
    int startDepth = 0;
    while ((1 << startDepth) < s.startGridSize) startDepth++;
    int levels = std::max(0, int(s.maxDepth) - startDepth);
    while (levels > 0 && 3 * (startDepth + levels) > 27) levels--;
    s.leafIndexLevels = -1;
    if ((1 << startDepth) != s.startGridSize || 3 * startDepth > 27) return;
    const uint64_t cells = uint64_t(1) << (3 * (startDepth + levels));
    s.dLeafIndex.alloc(cells + 1);
    SDFB_CUDA(cudaStreamSynchronize(cudaStream_t(0)));   // allocations are ordered on the default stream, the kernel runs on `st`
    uint32_t* bad = s.dLeafIndex.p + cells;
    SDFB_CUDA(cudaMemsetAsync(bad, 0, sizeof(uint32_t), st));
    leafIndexKernel<<<uint32_t((cells + 255) / 256), 256, 0, st>>>(s.dOctree.p, s.startGridSize, levels, s.dLeafIndex.p, bad);
    SDFB_CUDA(cudaGetLastError());
    uint32_t hostBad = 1;
    SDFB_CUDA(cudaMemcpyAsync(&hostBad, bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SDFB_CUDA(cudaStreamSynchronize(st));
    if (hostBad) { s.dLeafIndex.release(); return; }
    s.leafIndexLevels = levels;
}
#endif

#ifdef SDFB_QUERY_EXACT
void launchOctreeQueryExact(
#else
void launchOctreeQueryFast(
#endif
    const sdfb200_sdf& s, const float* dXyz, uint64_t n, float* dDist, float* dGrad, cudaStream_t st) {
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
    if (s.maxDepth > 16) throw Error(SDFB200_ERR_INVALID, "octree deeper than 16 levels");
#ifndef SDFB_QUERY_EXACT
    if (s.useCoopQuery && vec) {   // EXPERIMENTAL, set by sdfb200_query under SDFB200_QUERY_COOP=1
        if (dGrad) octreeQueryCoopKernel<true><<<grid, 256, 0, st>>>(s.dOctree.p, q, dXyz, n, dDist, dGrad);
        else octreeQueryCoopKernel<false><<<grid, 256, 0, st>>>(s.dOctree.p, q, dXyz, n, dDist, nullptr);
        SDFB_CUDA(cudaGetLastError());
        return;
    }
#endif
    if (s.useLeafIndex && s.leafIndexLevels >= 0) {   // EXPERIMENTAL, set by sdfb200_query under SDFB200_QUERY_INDEX=1
        const uint32_t* ix = s.dLeafIndex.p;
        const int lv = s.leafIndexLevels;
        if (dGrad) {
            if (vec) octreeQueryIndexedKernel<true, true><<<grid, 256, 0, st>>>(s.dOctree.p, ix, lv, q, dXyz, n, dDist, dGrad);
            else octreeQueryIndexedKernel<true, false><<<grid, 256, 0, st>>>(s.dOctree.p, ix, lv, q, dXyz, n, dDist, dGrad);
        } else {
            if (vec) octreeQueryIndexedKernel<false, true><<<grid, 256, 0, st>>>(s.dOctree.p, ix, lv, q, dXyz, n, dDist, nullptr);
            else octreeQueryIndexedKernel<false, false><<<grid, 256, 0, st>>>(s.dOctree.p, ix, lv, q, dXyz, n, dDist, nullptr);
        }
        SDFB_CUDA(cudaGetLastError());
        return;
    }
    if (dGrad) {
        if (vec) octreeQueryKernel<true, true><<<grid, 256, 0, st>>>(s.dOctree.p, q, dXyz, n, dDist, dGrad);
        else octreeQueryKernel<true, false><<<grid, 256, 0, st>>>(s.dOctree.p, q, dXyz, n, dDist, dGrad);
    } else {
        if (vec) octreeQueryKernel<false, true><<<grid, 256, 0, st>>>(s.dOctree.p, q, dXyz, n, dDist, nullptr);
        else octreeQueryKernel<false, false><<<grid, 256, 0, st>>>(s.dOctree.p, q, dXyz, n, dDist, nullptr);
    }
    SDFB_CUDA(cudaGetLastError());
}

}  // namespace sdfb200
