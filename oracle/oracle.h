/* TEST INFRASTRUCTURE — CPU restatement ("port") of the SdfLib hot paths. Not product code.
 *
 * Plain sequential C++ written from the behaviour of the reference (file:line cited at every
 * function in oracle.cpp). It is pinned against oracle/_ref/libsdfref.so — the unmodified
 * reference compiled here — by tests/test_oracle.py (bit-exact on every entry point),
 * and against the committed fixtures under tests/golden/ where the reference is not present.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may load it.
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcSdf OrcSdf;

/* mesh fixtures (src/utils/PrimitivesFactory.cpp:19-104) */
void orc_isosphere(uint32_t subdivisions, float* outVerts, uint32_t* outIdx, uint32_t* nVerts, uint32_t* nIdx);

/* kernels */
void orc_triangle_data(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, float* out37);
void orc_sq_dist(const float* tri37, const float* pts, uint64_t n, float* out);
void orc_signed_dist(const float* tri37, const float* v123, const float* pts, uint64_t n, int mode, float* outDist,
                     float* outGrad);
void orc_tricubic_coefficients(const float* values8x8, float nodeSize, float* out64);
void orc_tricubic_eval(const float* coeff64, const float* frac, uint64_t n, float* outValue, float* outGrad,
                       float* outVertexValues, float nodeSize);
float orc_error_estimate(const float* coeff64, const float* mid19x8, int rule, float decay);
int orc_is_near_minimize(float halfNodeSize, const float* vertRadius8, const float* tri9, float distThreshold,
                         uint32_t* outIter);
uint32_t orc_filter_triangles(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx,
                              const float* center3, float halfSize, const uint32_t* inTris, uint32_t nIn,
                              const uint32_t* cornerTris8, uint32_t* outTris);
void orc_nearest_triangle(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, const float* pts,
                          uint64_t n, uint32_t* outTri);

void orc_nearest_triangle_visits(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, const float* pts,
                                 uint64_t n, uint32_t* outTri, uint32_t* outVisits2);
void orc_nearest_triangle_visits_seeded(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, const float* pts,
                                        uint64_t n, const double* seeds, uint32_t* outTri, uint32_t* outVisits2);

/* whole structures. useCache=1 emulates the reference's 32^3 direct-mapped vertex cache
 * (TrianglesInfluence.h:934-991) — required for bit-identity with the reference's single-thread
 * build; useCache=0 is the history-free variant the level-synchronous GPU build is compared with.
 * numThreads keeps its layout-selecting meaning (<2 single DFS layout, >=2 per-start-voxel layout). */
OrcSdf* orc_build_octree(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, const float* box6,
                         uint32_t depth, uint32_t startDepth, int terminationRule, float param0, float param1,
                         int algorithm, uint32_t numThreads, int useCache);
OrcSdf* orc_build_exact(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, const float* box6,
                        uint32_t maxDepth, uint32_t startDepth, uint32_t minTrianglesPerNode, uint32_t numThreads,
                        int useCache);
void orc_delete(OrcSdf* sdf);
int orc_format(const OrcSdf* sdf);
int orc_save(const OrcSdf* sdf, const char* path);
OrcSdf* orc_load(const char* path);
void orc_sample_area(const OrcSdf* sdf, float* out6);
uint64_t orc_octree_data_size(const OrcSdf* sdf);
void orc_octree_data(const OrcSdf* sdf, uint32_t* out);
void orc_octree_header(const OrcSdf* sdf, int* startGridSize, uint32_t* maxDepth, float* valueRange,
                       float* minBorderValue);
/* exact-octree side arrays */
uint64_t orc_exact_sizes(const OrcSdf* sdf, uint64_t* nSets, uint64_t* nMasks, uint64_t* nTris);
void orc_exact_arrays(const OrcSdf* sdf, uint32_t* sets, uint8_t* masks, float* tris37);
void orc_exact_header(const OrcSdf* sdf, uint32_t* out8);
double orc_query(const OrcSdf* sdf, const float* pts, uint64_t n, float* outDist, float* outGrad, int numThreads);

#ifdef __cplusplus
}
#endif
