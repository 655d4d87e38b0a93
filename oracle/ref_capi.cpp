// TEST INFRASTRUCTURE — not product code, never linked into libsdfb200.so.
//
// extern "C" driver around the UNMODIFIED reference sources (compiled from where they lie
// under /root/reference by oracle/Makefile into oracle/_ref/libsdfref.so). It lets the Python
// tests and bench.py's `--impl reference` arm call the reference's own classes:
//   sdflib::OctreeSdf / ExactOctreeSdf ctors      (include/SdfLib/OctreeSdf.h:156, ExactOctreeSdf.h:91)
//   SdfFunction::getDistance / saveToFile / loadFromFile (include/SdfLib/SdfFunction.h:29-57)
//   TriangleUtils::* point-triangle kernels          (include/SdfLib/utils/TriangleUtils.h:76-376)
//   TriCubicInterpolation::*                         (include/SdfLib/InterpolationMethods.h:267-498)
//   GJK::IsNearMinimize                              (src/utils/GJK.cpp:830-866)
//   tmd::TriangleMeshDistance (nearest triangle)     (libs/InteractiveComputerGraphics/.../TriangleMeshDistance.h)
//   PrimitivesFactory::getIsosphere                  (src/utils/PrimitivesFactory.cpp:19-104)
// Nothing here restates reference logic; it only forwards.
#include "SdfLib/OctreeSdf.h"
#include "SdfLib/ExactOctreeSdf.h"
#include "SdfLib/TrianglesInfluence.h"
#include "SdfLib/InterpolationMethods.h"
#include "SdfLib/OctreeSdfUtils.h"
#include "SdfLib/utils/PrimitivesFactory.h"
#include "SdfLib/utils/GJK.h"
#include "SdfLib/utils/Timer.h"
#include <omp.h>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

using namespace sdflib;

namespace {
Mesh makeMesh(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx) {
    Mesh m(reinterpret_cast<glm::vec3*>(const_cast<float*>(verts)), nVerts, const_cast<uint32_t*>(idx), nIdx);
    m.computeBoundingBox();
    return m;
}
}  // namespace

extern "C" {

// ---- mesh fixtures -----------------------------------------------------------------------
// Returns counts; copies when the out pointers are non-null.
void ref_isosphere(uint32_t subdivisions, float* outVerts, uint32_t* outIdx, uint32_t* nVerts, uint32_t* nIdx) {
    auto mesh = PrimitivesFactory::getIsosphere(subdivisions);
    *nVerts = uint32_t(mesh->getVertices().size());
    *nIdx = uint32_t(mesh->getIndices().size());
    if (outVerts) std::memcpy(outVerts, mesh->getVertices().data(), sizeof(float) * 3 * *nVerts);
    if (outIdx) std::memcpy(outIdx, mesh->getIndices().data(), sizeof(uint32_t) * *nIdx);
}

// ---- kernel-level entry points -----------------------------------------------------------
static_assert(sizeof(TriangleUtils::TriangleData) == 37 * sizeof(float), "TriangleData layout");

void ref_triangle_data(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, float* out37) {
    Mesh m = makeMesh(verts, nVerts, idx, nIdx);
    auto td = TriangleUtils::calculateMeshTriangleData(m);
    std::memcpy(out37, td.data(), td.size() * sizeof(TriangleUtils::TriangleData));
}

void ref_sq_dist(const float* tri37, const float* pts, uint64_t n, float* out) {
    const auto& td = *reinterpret_cast<const TriangleUtils::TriangleData*>(tri37);
    for (uint64_t i = 0; i < n; i++)
        out[i] = TriangleUtils::getSqDistPointAndTriangle(glm::vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), td);
}

// mode 0: signed distance only (TriangleUtils.h:137); mode 1: + gradient with mesh vertices (:198);
// mode 2: + gradient self-contained (:292).
void ref_signed_dist(const float* tri37, const float* v123, const float* pts, uint64_t n, int mode, float* outDist,
                     float* outGrad) {
    const auto& td = *reinterpret_cast<const TriangleUtils::TriangleData*>(tri37);
    for (uint64_t i = 0; i < n; i++) {
        glm::vec3 p(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
        glm::vec3 g(0.0f);
        if (mode == 0) outDist[i] = TriangleUtils::getSignedDistPointAndTriangle(p, td);
        else if (mode == 1)
            outDist[i] = TriangleUtils::getSignedDistPointAndTriangle(
                p, td, glm::vec3(v123[0], v123[1], v123[2]), glm::vec3(v123[3], v123[4], v123[5]),
                glm::vec3(v123[6], v123[7], v123[8]), g);
        else outDist[i] = TriangleUtils::getSignedDistPointAndTriangle(p, td, g);
        if (outGrad) { outGrad[3 * i] = g.x; outGrad[3 * i + 1] = g.y; outGrad[3 * i + 2] = g.z; }
    }
}

void ref_tricubic_coefficients(const float* values8x8, float nodeSize, float* out64) {
    std::array<std::array<float, 8>, 8> in;
    std::memcpy(in.data(), values8x8, sizeof(in));
    std::array<float, 64> out;
    std::vector<uint32_t> noTris;
    Mesh noMesh;
    std::vector<TriangleUtils::TriangleData> noData;
    TriCubicInterpolation::calculateCoefficients(in, nodeSize, noTris, noMesh, noData, out);
    std::memcpy(out64, out.data(), sizeof(out));
}

void ref_tricubic_eval(const float* coeff64, const float* frac, uint64_t n, float* outValue, float* outGrad,
                       float* outVertexValues, float nodeSize) {
    std::array<float, 64> c;
    std::memcpy(c.data(), coeff64, sizeof(c));
    for (uint64_t i = 0; i < n; i++) {
        glm::vec3 f(frac[3 * i], frac[3 * i + 1], frac[3 * i + 2]);
        if (outValue) outValue[i] = TriCubicInterpolation::interpolateValue(c, f);
        if (outGrad) {
            glm::vec3 g = TriCubicInterpolation::interpolateGradient(c, f);
            outGrad[3 * i] = g.x; outGrad[3 * i + 1] = g.y; outGrad[3 * i + 2] = g.z;
        }
        if (outVertexValues) {
            std::array<float, 8> vv;
            TriCubicInterpolation::interpolateVertexValues(c, f, nodeSize, vv);
            std::memcpy(outVertexValues + 8 * i, vv.data(), sizeof(vv));
        }
    }
}

// rule: 1 trapezoid, 2 simpson, 3 by-distance (OctreeSdf.h:100-106)
float ref_error_estimate(const float* coeff64, const float* mid19x8, int rule, float decay) {
    std::array<float, 64> c;
    std::memcpy(c.data(), coeff64, sizeof(c));
    std::array<std::array<float, 8>, 19> mid;
    std::memcpy(mid.data(), mid19x8, sizeof(mid));
    if (rule == 2) return estimateErrorFunctionIntegralBySimpsonsRule<TriCubicInterpolation>(c, mid);
    if (rule == 3) return estimateDecayErrorFunctionIntegralByTrapezoidRule<TriCubicInterpolation>(c, mid, decay);
    return estimateErrorFunctionIntegralByTrapezoidRule<TriCubicInterpolation>(c, mid);
}

int ref_is_near_minimize(float halfNodeSize, const float* vertRadius8, const float* tri9, float distThreshold,
                         uint32_t* outIter) {
    std::array<float, 8> r;
    std::memcpy(r.data(), vertRadius8, sizeof(r));
    std::array<glm::vec3, 3> t = {glm::vec3(tri9[0], tri9[1], tri9[2]), glm::vec3(tri9[3], tri9[4], tri9[5]),
                                  glm::vec3(tri9[6], tri9[7], tri9[8])};
    uint32_t iter = 0;
    bool res = GJK::IsNearMinimize(halfNodeSize, r, t, distThreshold, &iter);
    if (outIter) *outIter = iter;
    return res ? 1 : 0;
}

// PerNodeRegionTrianglesInfluence::filterTriangles (TrianglesInfluence.h:767-860)
uint32_t ref_filter_triangles(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx,
                              const float* center3, float halfSize, const uint32_t* inTris, uint32_t nIn,
                              const uint32_t* cornerTris8, uint32_t* outTris) {
    Mesh m = makeMesh(verts, nVerts, idx, nIdx);
    auto td = TriangleUtils::calculateMeshTriangleData(m);
    PerNodeRegionTrianglesInfluence<NoneInterpolation> infl;
    std::vector<uint32_t> in(inTris, inTris + nIn), out;
    std::array<std::array<float, 0>, 8> vv;
    std::array<uint32_t, 8> vi;
    std::memcpy(vi.data(), cornerTris8, sizeof(vi));
    infl.filterTriangles(glm::vec3(center3[0], center3[1], center3[2]), halfSize, in, out, vv, vi, m, td);
    std::memcpy(outTris, out.data(), out.size() * sizeof(uint32_t));
    return uint32_t(out.size());
}

// Nearest triangle through the vendored double-precision BVH, exactly as VHQueries calls it
// (TrianglesInfluence.h:884-924).
void ref_nearest_triangle(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, const float* pts,
                          uint64_t n, uint32_t* outTri) {
    Mesh m = makeMesh(verts, nVerts, idx, nIdx);
    ICG icg(m);
    for (uint64_t i = 0; i < n; i++)
        outTri[i] = icg.getNearestTriangle(glm::vec3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
}

// ---- whole-structure entry points --------------------------------------------------------
SdfFunction* ref_build_octree(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx,
                              const float* box6, uint32_t depth, uint32_t startDepth, int terminationRule,
                              float param0, float param1, int algorithm, uint32_t numThreads, double* outSeconds) {
    Mesh m = makeMesh(verts, nVerts, idx, nIdx);
    BoundingBox box(glm::vec3(box6[0], box6[1], box6[2]), glm::vec3(box6[3], box6[4], box6[5]));
    OctreeSdf::TerminationRuleParams params;
    params.params = {param0, param1};
    auto t0 = std::chrono::steady_clock::now();
    auto* sdf = new OctreeSdf(m, box, depth, startDepth, OctreeSdf::TerminationRule(terminationRule), params,
                              OctreeSdf::InitAlgorithm(algorithm), numThreads);
    if (outSeconds) *outSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return sdf;
}

SdfFunction* ref_build_exact(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx,
                             const float* box6, uint32_t maxDepth, uint32_t startDepth, uint32_t minTrianglesPerNode,
                             uint32_t numThreads, double* outSeconds) {
    Mesh m = makeMesh(verts, nVerts, idx, nIdx);
    BoundingBox box(glm::vec3(box6[0], box6[1], box6[2]), glm::vec3(box6[3], box6[4], box6[5]));
    auto t0 = std::chrono::steady_clock::now();
    auto* sdf = new ExactOctreeSdf(m, box, maxDepth, startDepth, minTrianglesPerNode, numThreads);
    if (outSeconds) *outSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return sdf;
}

void ref_delete(SdfFunction* sdf) { delete sdf; }  // virtual dtor (SdfFunction.h:24)

int ref_format(SdfFunction* sdf) { return int(sdf->getFormat()); }

int ref_save(SdfFunction* sdf, const char* path) { return sdf->saveToFile(path) ? 1 : 0; }

SdfFunction* ref_load(const char* path) { return SdfFunction::loadFromFile(path).release(); }

void ref_sample_area(SdfFunction* sdf, float* out6) {
    BoundingBox b = sdf->getSampleArea();
    out6[0] = b.min.x; out6[1] = b.min.y; out6[2] = b.min.z;
    out6[3] = b.max.x; out6[4] = b.max.y; out6[5] = b.max.z;
}

// OctreeSdf getters (OctreeSdf.h:177-208)
uint64_t ref_octree_data_size(SdfFunction* sdf) {
    if (auto* o = dynamic_cast<OctreeSdf*>(sdf)) return o->getOctreeData().size();
    if (auto* e = dynamic_cast<ExactOctreeSdf*>(sdf)) return e->getOctreeData().size();
    return 0;
}
void ref_octree_data(SdfFunction* sdf, uint32_t* out) {
    if (auto* o = dynamic_cast<OctreeSdf*>(sdf))
        std::memcpy(out, o->getOctreeData().data(), o->getOctreeData().size() * sizeof(uint32_t));
    else if (auto* e = dynamic_cast<ExactOctreeSdf*>(sdf))
        std::memcpy(out, e->getOctreeData().data(), e->getOctreeData().size() * 2 * sizeof(uint32_t));
}
void ref_octree_header(SdfFunction* sdf, int* startGridSize, uint32_t* maxDepth, float* valueRange,
                       float* minBorderValue) {
    if (auto* o = dynamic_cast<OctreeSdf*>(sdf)) {
        *startGridSize = o->getStartGridSize().x;
        *maxDepth = o->getOctreeMaxDepth();
        *valueRange = o->getOctreeValueRange();
        *minBorderValue = o->getOctreeMinBorderValue();
    } else if (auto* e = dynamic_cast<ExactOctreeSdf*>(sdf)) {
        *startGridSize = e->getStartGridSize().x;
        *maxDepth = e->getOctreeMaxDepth();
        *valueRange = float(e->getMinTrianglesInLeafs());
        *minBorderValue = float(e->getMaxTrianglesInLeafs());
    }
}

// Bulk query. numThreads > 1 uses an external `omp parallel for` — legal for OctreeSdf (const,
// re-entrant); ExactOctreeSdf::getDistance writes a mutable scratch (ExactOctreeSdf.h:178), so
// every thread queries its own copy round-tripped through saveToFile/loadFromFile.
// Returns elapsed seconds of the query loop.
double ref_query(SdfFunction* sdf, const float* pts, uint64_t n, float* outDist, float* outGrad, int numThreads) {
    std::vector<std::unique_ptr<SdfFunction>> copies;
    std::vector<SdfFunction*> perThread(size_t(numThreads > 1 ? numThreads : 1), sdf);
    if (numThreads > 1 && sdf->getFormat() == SdfFunction::EXACT_OCTREE) {
        char path[128];
        std::snprintf(path, sizeof(path), "/tmp/sdfref_clone_%d.bin", int(omp_get_wtime() * 1e6) & 0xffffff);
        sdf->saveToFile(path);
        for (int t = 1; t < numThreads; t++) {
            copies.push_back(SdfFunction::loadFromFile(path));
            perThread[size_t(t)] = copies.back().get();
        }
        std::remove(path);
    }
    auto t0 = std::chrono::steady_clock::now();
    if (numThreads > 1) {
#pragma omp parallel for schedule(static) num_threads(numThreads)
        for (int64_t i = 0; i < int64_t(n); i++) {
            SdfFunction* s = perThread[size_t(omp_get_thread_num())];
            glm::vec3 p(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
            if (outGrad) {
                glm::vec3 g(0.0f);
                outDist[i] = s->getDistance(p, g);
                outGrad[3 * i] = g.x; outGrad[3 * i + 1] = g.y; outGrad[3 * i + 2] = g.z;
            } else outDist[i] = s->getDistance(p);
        }
    } else {
        for (uint64_t i = 0; i < n; i++) {
            glm::vec3 p(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
            if (outGrad) {
                glm::vec3 g(0.0f);
                outDist[i] = sdf->getDistance(p, g);
                outGrad[3 * i] = g.x; outGrad[3 * i + 1] = g.y; outGrad[3 * i + 2] = g.z;
            } else outDist[i] = sdf->getDistance(p);
        }
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int ref_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
