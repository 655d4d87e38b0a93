/* sdfb200 — C-ABI of the B200-native SdfLib hot paths (octree construction + bulk getDistance).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types. Everything behind it
 * is hand-written CUDA for sm_100a (sdflib_b200/csrc). The C++ mirror of the reference classes
 * (include/SdfLib/*.h) and the Python host (sdflib_b200/) both sit on top of exactly these symbols.
 *
 * Reference interfaces replaced (all paths relative to the reference tree):
 *   - sdflib::OctreeSdf ctor            include/SdfLib/OctreeSdf.h:156-172, src/sdf/OctreeSdf.cpp:18-86
 *   - sdflib::ExactOctreeSdf ctor       include/SdfLib/ExactOctreeSdf.h:91-93, src/sdf/ExactOctreeSdf.cpp:7-31
 *   - SdfFunction::getDistance (x2)     include/SdfLib/SdfFunction.h:29-36, src/sdf/OctreeSdf.cpp:93-152,
 *                                       src/sdf/ExactOctreeSdf.cpp:38-320
 *   - SdfFunction::saveToFile/loadFromFile  src/sdf/SdfFunction.cpp:9-79 (same .bin bytes)
 *   - the Unity C exports               src/tools/SdfLibUnity/SdfExportFunc.h:16-58 (see sdfb200_unity.h)
 *
 * Error behaviour: every call returns SDFB200_OK (0) or a negative code and records a thread-local
 * message retrievable with sdfb200_last_error(); the library never calls exit() (the reference's BVH
 * does on an empty mesh) and never falls back to the CPU: without a CUDA device every compute entry
 * point fails with SDFB200_ERR_CUDA.
 */
#ifndef SDFB200_H
#define SDFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDFB200_VERSION 100

enum {
    SDFB200_OK = 0,
    SDFB200_ERR_INVALID = -1,      /* bad argument (null pointer, depth out of range, empty mesh, ...) */
    SDFB200_ERR_CUDA = -2,         /* CUDA runtime error / no device */
    SDFB200_ERR_IO = -3,           /* file cannot be opened / truncated / unknown format */
    SDFB200_ERR_UNSUPPORTED = -4   /* option of the reference API that is not built yet */
};

/* SdfFunction::SdfFormat (include/SdfLib/SdfFunction.h:16-22) */
enum { SDFB200_FORMAT_GRID = 0, SDFB200_FORMAT_OCTREE = 1, SDFB200_FORMAT_EXACT_OCTREE = 2 };
/* OctreeSdf::InitAlgorithm (include/SdfLib/OctreeSdf.h:23-28) */
enum { SDFB200_ALG_UNIFORM = 0, SDFB200_ALG_NO_CONTINUITY = 1, SDFB200_ALG_CONTINUITY = 2 };
/* OctreeSdf::TerminationRule (include/SdfLib/OctreeSdf.h:100-106) */
enum { SDFB200_RULE_NONE = 0, SDFB200_RULE_TRAPEZOIDAL = 1, SDFB200_RULE_SIMPSONS = 2, SDFB200_RULE_BY_DISTANCE = 3 };

/* sdfb200_query flags */
enum {
    SDFB200_QUERY_DEVICE_POINTERS = 1, /* xyz / dist / grad are device pointers on the handle's GPU */
    SDFB200_QUERY_EXACT_ORDER = 2      /* evaluate the leaf polynomial in the reference's literal operation
                                          order without FMA (bit-identical to the CPU reference); default is
                                          the FMA Horner form, within 1e-5 of it */
};

typedef struct sdfb200_sdf sdfb200_sdf; /* opaque: host mirrors + device buffers of one SdfFunction */

typedef struct sdfb200_info {
    int32_t format;               /* SDFB200_FORMAT_* */
    float box_min[3], box_max[3]; /* getGridBoundingBox() == getSampleArea() (cubified input box) */
    int32_t start_grid_size;      /* getStartGridSize().x */
    uint32_t max_depth;           /* getOctreeMaxDepth() */
    /* OCTREE */
    float value_range;            /* getOctreeValueRange() */
    float min_border_value;       /* getOctreeMinBorderValue() */
    /* EXACT_OCTREE */
    uint32_t start_depth, min_triangles_in_leafs, max_triangles_in_leafs, max_triangles_encoded_in_leafs,
        bit_encoding_start_depth, bits_per_index;
    /* array sizes */
    uint64_t octree_words;        /* OCTREE: #uint32 of mOctreeData; EXACT: #nodes (2 uint32 each) */
    uint64_t triangle_sets_words; /* EXACT: #uint32 of mTrianglesSets */
    uint64_t triangle_masks_bytes;/* EXACT: #bytes of mTrianglesMasks */
    uint64_t num_triangles;       /* EXACT: #TriangleData (37 floats each) */
    int32_t device;               /* CUDA device the structure lives on */
} sdfb200_info;

/* Timings of the last build on this handle, milliseconds (host wall clock around synchronised phases). */
typedef struct sdfb200_build_stats {
    double total_ms, triangle_data_ms, bvh_ms, upload_ms, levels_ms, layout_ms, download_ms;
    uint64_t nodes_processed, leaves, samples_evaluated, kernel_launches;
} sdfb200_build_stats;

const char* sdfb200_last_error(void);
int sdfb200_version(void);
int sdfb200_device_count(void);        /* 0 when no CUDA device is visible */
int sdfb200_set_device(int device);
/* The library recycles device blocks (size-class free lists over the stream-ordered allocator) and pinned host blocks
 * between calls so that repeated builds do not pay for allocation. This returns everything that is cached and not in
 * use by a live handle to the driver (device side: after a device synchronisation). */
int sdfb200_release_cached_memory(void);    /* device used by subsequent build/load calls of this thread */

/* ---- construction (hot path 1) ----------------------------------------------------------------
 * vertices: numVertices * 3 floats; indices: numIndices uint32 (3 per triangle); box6 = min xyz, max xyz.
 * Same argument meaning as the reference constructors; numThreads keeps its layout-selecting meaning
 * (< 2: single depth-first layout, >= 2: per-start-voxel layout, src/sdf/OctreeSdfDepthFirst.h:395-503). */
int sdfb200_build_octree(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                         const float* box6, uint32_t depth, uint32_t startDepth, int terminationRule, float param0,
                         float param1, int initAlgorithm, uint32_t numThreads, sdfb200_sdf** out);
int sdfb200_build_exact(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                        const float* box6, uint32_t maxDepth, uint32_t startDepth, uint32_t minTrianglesPerNode,
                        uint32_t numThreads, sdfb200_sdf** out);

/* Sharded construction (one process per GPU). The start-depth voxels ("roots") are the reference's
 * own task decomposition (src/sdf/OctreeSdfDepthFirst.h:433-469, include/SdfLib/ExactOctreeSdfDepthFirst.h:534-574);
 * a root's owner is chosen by estimated work (longest-processing-time greedy over per-root weights that every rank
 * derives from the replicated levels above the start depth). Protocol, identical on every rank:
 *   1. sdfb200_build_*_shard      levels + subtree sizes of the own roots (GPU)
 *   2. sdfb200_shard_sizes        K values per start-grid slot (K = 1 OCTREE: words; K = 3 EXACT: node records, set
 *                                 words, mask bytes), 0 for roots of other ranks  -> caller all-reduces (SUM)
 *   3. sdfb200_shard_finish       global offsets from all sizes; own blocks emitted with their FINAL indices (GPU)
 *   4. sdfb200_shard_words/export flat uint32 payload in a caller-owned DEVICE buffer -> caller all-gathers (NCCL),
 *                                 every rank's payload at q * strideWords
 *   5. sdfb200_assemble           block copies into the complete arrays; the handle is then a normal SdfFunction
 * The collectives themselves stay with the caller (torch.distributed / NCCL in sdflib_b200/sharded.py). */
int sdfb200_build_octree_shard(const float* vertices, uint32_t numVertices, const uint32_t* indices,
                               uint32_t numIndices, const float* box6, uint32_t depth, uint32_t startDepth,
                               int terminationRule, float param0, float param1, int initAlgorithm,
                               uint32_t numThreads, uint32_t rank, uint32_t worldSize, sdfb200_sdf** out);
/* InitAlgorithm::CONTINUITY over several ranks (second row). Its neighbour probes cross start-voxel
 * boundaries, so the octree logic is REPLICATED on every rank (deterministic, a few percent of the build) and the
 * nearest-triangle sampling — 94 % of the kernel time — is what the ranks share: per depth every rank traverses the BVH
 * for its slice of the level's distinct sample positions and the slices are all-gathered (16 bytes per sample).
 * The collective stays with the caller: `allgather` must gather `bytesPerRank` bytes from every rank's dSend into dRecv
 * in rank order (device pointers owned by the library, stream-ordered on the legacy default stream or complete on
 * return) and return 0. Every rank returns the complete, identical structure (no assemble step). */
typedef int (*sdfb200_allgather_fn)(void* user, const void* dSend, void* dRecv, uint64_t bytesPerRank);
int sdfb200_build_octree_collective(const float* vertices, uint32_t numVertices, const uint32_t* indices,
                                    uint32_t numIndices, const float* box6, uint32_t depth, uint32_t startDepth,
                                    int terminationRule, float param0, float param1, int initAlgorithm,
                                    uint32_t numThreads, uint32_t rank, uint32_t worldSize,
                                    sdfb200_allgather_fn allgather, void* user, sdfb200_sdf** out);
int sdfb200_build_exact_shard(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                              const float* box6, uint32_t maxDepth, uint32_t startDepth, uint32_t minTrianglesPerNode,
                              uint32_t numThreads, uint32_t rank, uint32_t worldSize, sdfb200_sdf** out);

/* ---- prepared meshes (mesh ingestion on the device; reference: calculateMeshTriangleData, src/utils/TriangleUtils.cpp:7-428,
 * and tmd::TriangleMeshDistance's BVH, libs/InteractiveComputerGraphics/.../TriangleMeshDistance.h:421-490).
 * A prepared mesh holds, on the current device, everything the builders read: vertices, indices, TriangleData (computed on
 * the GPU, same bits as the reference's host loop), with SDFB200_MESH_BVH the nearest-triangle BVH (built on the GPU
 * with std::sort's exact tie order, bvh_device.cu; SDFB200_HOST_BVH=1 selects the host builder) and with SDFB200_MESH_EXACT the ExactOctreeSdf side arrays. One mesh serves any
 * number of builds; sdfb200_mesh_export / sdfb200_mesh_import replicate it on other ranks (broadcast the blob). */
typedef struct sdfb200_mesh sdfb200_mesh;
enum { SDFB200_MESH_BVH = 1, SDFB200_MESH_EXACT = 2,
       SDFB200_MESH_ALL_HOST_THREADS = 4 /* the host part may use every core although other ranks share the node */ };
int sdfb200_mesh_create(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices, int parts,
                        sdfb200_mesh** out);
void sdfb200_mesh_free(sdfb200_mesh* mesh);
int sdfb200_mesh_blob_bytes(const sdfb200_mesh* mesh, uint64_t* outBytes);
int sdfb200_mesh_export(const sdfb200_mesh* mesh, void* devicePtr, uint64_t capacityBytes);   /* device buffer on the mesh's GPU */
int sdfb200_mesh_import(const void* devicePtr, uint64_t bytes, sdfb200_mesh** out);          /* device buffer on the current GPU */
int sdfb200_mesh_stats(const sdfb200_mesh* mesh, double* triangleDataMs, double* bvhMs, double* uploadMs);
/* The builders of this header with the mesh given as a prepared mesh (same arguments otherwise). */
int sdfb200_build_octree_from_mesh(const sdfb200_mesh* mesh, const float* box6, uint32_t depth, uint32_t startDepth,
                                   int terminationRule, float param0, float param1, int initAlgorithm, uint32_t numThreads,
                                   uint32_t rank, uint32_t worldSize, sdfb200_sdf** out);
int sdfb200_build_octree_collective_from_mesh(const sdfb200_mesh* mesh, const float* box6, uint32_t depth, uint32_t startDepth,
                                              int terminationRule, float param0, float param1, uint32_t rank, uint32_t worldSize,
                                              sdfb200_allgather_fn allgather, void* user, sdfb200_sdf** out);
int sdfb200_build_exact_from_mesh(const sdfb200_mesh* mesh, const float* box6, uint32_t maxDepth, uint32_t startDepth,
                                  uint32_t minTrianglesPerNode, uint32_t numThreads, uint32_t rank, uint32_t worldSize,
                                  sdfb200_sdf** out);

/* ---- one process, several GPUs of one box: what the C++ drop-in classes call when more than one device
 * is requested. Host set-up once, the mesh replicated by peer copies, one host thread per device building the
 * start-depth voxels it owns (assigned by estimated work), ONE ncclAllGather over NVLink (libnccl.so.2 is loaded at run
 * time; peer copies when it is absent) assembling the complete structure on EVERY listed device. CONTINUITY shares the
 * per-depth BVH sampling instead (see sdfb200_build_octree_collective). outHandles receives nDevices handles,
 * outHandles[k] living on devices[k]; each is a complete SdfFunction (free every one with sdfb200_free). */
int sdfb200_build_octree_multi(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                               const float* box6, uint32_t depth, uint32_t startDepth, int terminationRule, float param0,
                               float param1, int initAlgorithm, uint32_t numThreads, const int* devices, uint32_t nDevices,
                               sdfb200_sdf** outHandles);
int sdfb200_build_exact_multi(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                              const float* box6, uint32_t maxDepth, uint32_t startDepth, uint32_t minTrianglesPerNode,
                              uint32_t numThreads, const int* devices, uint32_t nDevices, sdfb200_sdf** outHandles);
int sdfb200_nccl_available(void);   /* 1 when libnccl.so.2 could be loaded (SDFB200_NCCL_LIB overrides the name) */

int sdfb200_shard_sizes(const sdfb200_sdf* shard, uint32_t* outSizes, uint64_t capacity, uint64_t* outCount);
int sdfb200_shard_finish(sdfb200_sdf* shard, const uint32_t* allSizes, uint64_t count);
int sdfb200_shard_words(const sdfb200_sdf* shard, uint64_t* outWords);
int sdfb200_shard_export(const sdfb200_sdf* shard, uint32_t* devicePtr, uint64_t capacityWords);
int sdfb200_assemble(sdfb200_sdf* shard, const uint32_t* gatheredDevicePtr, const uint64_t* wordsPerRank,
                     uint64_t strideWords, uint32_t worldSize);

/* ---- persistence (.bin, cereal PortableBinary layout of the reference) ------------------------ */
int sdfb200_save(const sdfb200_sdf* sdf, const char* path);
int sdfb200_load(const char* path, sdfb200_sdf** out);
void sdfb200_free(sdfb200_sdf* sdf);

/* ---- getters ---------------------------------------------------------------------------------- */
int sdfb200_get_info(const sdfb200_sdf* sdf, sdfb200_info* out);
int sdfb200_get_build_stats(const sdfb200_sdf* sdf, sdfb200_build_stats* out);
/* OCTREE: capacity in uint32 words >= octree_words. EXACT: 2 * octree_words. */
int sdfb200_get_octree_data(const sdfb200_sdf* sdf, uint32_t* out, uint64_t capacityWords);
int sdfb200_get_exact_arrays(const sdfb200_sdf* sdf, uint32_t* triangleSets, uint8_t* triangleMasks,
                             float* triangleData37);
/* Device pointer of the structure arrays (OCTREE: mOctreeData) for zero-copy consumers. */
int sdfb200_get_device_octree(const sdfb200_sdf* sdf, const uint32_t** outDevicePtr);

/* ---- bulk getDistance (hot path 2) ------------------------------------------------------------
 * xyz: n packed float3. dist: n floats. grad: NULL or n packed float3 (getDistance(p, grad)).
 * Host pointers by default: the call returns when the results are in the caller's memory. Pinned (cudaHostAlloc /
 * cudaHostRegister) buffers are copied directly in 2 M-query chunks alternating between two streams; pageable buffers
 * go through a pinned ring owned by the handle (filled by all host threads); batches of up to 2048 queries — the
 * scalar getDistance(p) of the reference API — use one mapped pinned slot and cost a kernel launch and a stream
 * synchronisation. With SDFB200_QUERY_DEVICE_POINTERS the pointers are device pointers and the call only enqueues the
 * kernel on `cudaStream` (a cudaStream_t, NULL = default stream) without synchronising.
 * Threading: device-pointer calls are re-entrant (they share no mutable state). Host-pointer calls on ONE handle are
 * serialised by a mutex inside the handle — correct from any number of threads (the reference's getDistance is a const
 * read its users call from OpenMP loops), but they do not run concurrently; use one bulk call, or device pointers. */
int sdfb200_query(sdfb200_sdf* sdf, const float* xyz, uint64_t n, float* dist, float* grad, int flags,
                  void* cudaStream);

/* Sphere tracing over an OctreeSdf: the consumer of getDistance in the reference's viewer (raycast() of
 * src/render_engine/shaders/sdfOctreeRender.comp:392-410), one ray per GPU thread:
 *     while (last > epsilon && travelled < farDistance && it < maxIterations) { hit = pos; last = getDistance(pos);
 *                                                                            step = max(last, 0); travelled += step; pos += dir * step; it++; }
 * origins / directions: n packed float3 (directions as given: normalise them for distances in world units).
 * outHit: n packed float3, the last position evaluated; outTravelled: n floats, the distance marched when the surface was
 * reached (last < epsilon), -1 otherwise; outIterations: NULL or n uint32. flags: SDFB200_QUERY_DEVICE_POINTERS and
 * SDFB200_QUERY_EXACT_ORDER as for sdfb200_query (with the latter the march equals a CPU loop over the reference's
 * getDistance bit for bit). The shader's constants are epsilon = 1e-5, maxIterations = 1024. */
int sdfb200_sphere_trace(sdfb200_sdf* sdf, const float* origins, const float* directions, uint64_t n, float epsilon,
                         float farDistance, uint32_t maxIterations, float* outHit, float* outTravelled,
                         uint32_t* outIterations, int flags, void* cudaStream);

/* Kernel-level entry points (used by the parity tests; each runs the same device function the
 * builders use). All pointers are HOST pointers. */
int sdfb200_triangle_data(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                          float* out37);
int sdfb200_nearest_triangle(const float* vertices, uint32_t numVertices, const uint32_t* indices,
                             uint32_t numIndices, const float* xyz, uint64_t n, uint32_t* outTriangle);
/* The nearest-triangle BVH (reference: tmd::TriangleMeshDistance::_build_tree, TriangleMeshDistance.h:421-490) as the
 * traversal kernels read it: 2 * triangles - 1 nodes of 80 bytes in the reference's push order — float64 left sphere
 * (centre xyz, radius), float64 right sphere, int32 left / right links (>= 0: inner child, < 0: ~triangleId), int32 leaf
 * flag, int32 pad. sdfb200_mesh_bvh copies out the tree a prepared mesh holds (built on the GPU), sdfb200_bvh_host runs the
 * host builder (libstdc++'s own std::sort routines); the parity tests compare the two node for node. */
int sdfb200_mesh_bvh(const sdfb200_mesh* mesh, void* outNodes, uint64_t capacityNodes);
int sdfb200_bvh_host(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                     void* outNodes, uint64_t capacityNodes);
int sdfb200_point_triangle(const float* tri37, const float* v123, const float* xyz, uint64_t n, int mode,
                           float* outDist, float* outGrad);

/* Fixture generator of the benchmark configs: PrimitivesFactory::getIsosphere
 * (src/utils/PrimitivesFactory.cpp:19-104), same vertex/triangle order. Pass NULL outputs to query sizes. */
int sdfb200_make_isosphere(uint32_t subdivisions, float* outVertices, uint32_t* outIndices, uint32_t* numVertices,
                           uint32_t* numIndices);

#ifdef __cplusplus
}
#endif
#endif /* SDFB200_H */
