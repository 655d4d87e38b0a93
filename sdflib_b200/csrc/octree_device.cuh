// Device code shared by the OctreeSdf builders (octree_build.cu: NO_CONTINUITY, octree_cont.cu: CONTINUITY):
// float64 sphere-BVH nearest-triangle descent, point sampling, the 64x64 Hermite map and the exact-order
// polynomial evaluation. Everything has internal linkage: each translation unit gets its own copy of the
// __constant__ tables (no relocatable device code).
#pragma once
#include <cfloat>
#include <cmath>

#include "device_utils.cuh"
#include "sdf_internal.h"

namespace sdfb200 {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr uint32_t kNoChild = 0xFFFFFFFFu;

// lattice index L = x + 3y + 9z of the 19 mid-points, in the reference's sample order
__constant__ int cSampleLattice[19] = {1, 3, 4, 5, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 19, 21, 22, 23, 25};

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

// ---- device helpers ------------------------------------------------------------------------------

__device__ __forceinline__ d3 vertexD(const DeviceMesh& m, uint32_t v) {
    const f3 p = m.verts[v];
    return mkd(double(p.x), double(p.y), double(p.z));
}

// Nearest triangle id with the reference's traversal order: near child first, the far child is
// re-tested against the running best when the near subtree is done, leaves replace the best only on
// strict '<' against the re-squared running distance.
__device__ uint32_t bvhNearest(const DeviceMesh& m, f3 pf) {
    const d3 p = mkd(double(pf.x), double(pf.y), double(pf.z));
    double best = DBL_MAX;
    int bestTri = -1;
    int stackNode[48];
    double stackDist[48];
    int sp = 0;
    int cur = 0;
    for (;;) {
        const BvhNode nd = m.bvh[cur];
        bool descend = false;
        if (nd.left < 0) {
            const uint32_t t = uint32_t(nd.right);
            const double d2 = eberlySqDist(p, vertexD(m, m.idx[3 * t]), vertexD(m, m.idx[3 * t + 1]), vertexD(m, m.idx[3 * t + 2]));
            if (d2 < best * best) { best = sqrt(d2); bestTri = nd.right; }
        } else {
            const d3 dl3 = p - mkd(nd.lc[0], nd.lc[1], nd.lc[2]);
            const d3 dr3 = p - mkd(nd.rc[0], nd.rc[1], nd.rc[2]);
            const double dl = sqrt(ddot(dl3, dl3)) - nd.lr;
            const double dr = sqrt(ddot(dr3, dr3)) - nd.rr;
            const bool leftFirst = dl < dr;
            const int first = leftFirst ? nd.left : nd.right, second = leftFirst ? nd.right : nd.left;
            const double dFirst = leftFirst ? dl : dr, dSecond = leftFirst ? dr : dl;
            stackNode[sp] = second;
            stackDist[sp] = dSecond;
            sp++;
            if (dFirst < best) { cur = first; descend = true; }
        }
        if (descend) continue;
        bool found = false;
        while (sp > 0) {
            sp--;
            if (stackDist[sp] < best) { cur = stackNode[sp]; found = true; break; }
        }
        if (!found) break;
    }
    return uint32_t(bestTri);
}

// TriCubicInterpolation::calculatePointValues: (signed distance, unit gradient) of the nearest triangle
__device__ __forceinline__ float4 samplePoint(const DeviceMesh& m, f3 p) {
    const uint32_t t = bvhNearest(m, p);
    f3 g;
    const float d = signedDistGradMesh(p, m.tris[t], m.verts[m.idx[3 * t]], m.verts[m.idx[3 * t + 1]], m.verts[m.idx[3 * t + 2]], g);
    return make_float4(d, g.x, g.y, g.z);
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

inline void uploadMesh(MeshOnDevice& m, const HostMesh& mesh, const TriVec& tris, const RawVec<BvhNode>* bvh) {
    m.numTriangles = mesh.numTriangles();
    m.verts.alloc(mesh.nVerts); m.verts.upload(mesh.verts, mesh.nVerts);
    m.idx.alloc(mesh.nIdx); m.idx.upload(mesh.idx, mesh.nIdx);
    m.tris.alloc(tris.size()); m.tris.upload(tris.data(), tris.size());
    if (bvh) { m.bvh.alloc(bvh->size()); m.bvh.upload(bvh->data(), bvh->size()); }
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
