// The C-ABI of libsdfb200.so (include/sdfb200.h): argument checking, error translation, host<->device
// staging. No compute happens here; every entry point that needs the GPU fails with SDFB200_ERR_CUDA
// when no device is present — there is no CPU fallback.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <thread>
#include <vector>

#include "sdf_internal.h"

namespace sdfb200 {
namespace {
thread_local std::string gLastError;
thread_local int gDevice = 0;
}
void setLastError(const std::string& msg) { gLastError = msg; }

namespace {
template <class F> int guarded(F&& f) {
    try {
        f();
        return SDFB200_OK;
    } catch (const Error& e) {
        setLastError(e.what());
        return e.code;
    } catch (const std::bad_alloc&) {
        setLastError("out of host memory");
        return SDFB200_ERR_INVALID;
    } catch (const std::exception& e) {
        setLastError(e.what());
        return SDFB200_ERR_INVALID;
    }
}

template <class F> int guarded(const char* range, F&& f) {   // the same inside an NVTX range named after the entry point
    NvtxRange nvtx(range);
    return guarded(static_cast<F&&>(f));
}

void requireDevice() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        throw Error(SDFB200_ERR_CUDA, "no CUDA device available (sdfb200 has no CPU fallback)");
    }
    SDFB_CUDA(cudaSetDevice(gDevice));
    configureDevicePool(gDevice);
}

HostMesh checkedMesh(const float* v, uint32_t nv, const uint32_t* idx, uint32_t ni) {
    if (!v || !idx) throw Error(SDFB200_ERR_INVALID, "null mesh pointer");
    if (nv == 0 || ni < 3 || ni % 3 != 0) throw Error(SDFB200_ERR_INVALID, "empty mesh or index count not a multiple of 3");
    for (uint32_t i = 0; i < ni; i++)
        if (idx[i] >= nv) throw Error(SDFB200_ERR_INVALID, "triangle index out of range");
    return HostMesh{reinterpret_cast<const f3*>(v), nv, idx, ni};
}

void checkBox(const float* b) {
    if (!b) throw Error(SDFB200_ERR_INVALID, "null bounding box");
    for (int i = 0; i < 3; i++)
        if (!(b[i + 3] > b[i])) throw Error(SDFB200_ERR_INVALID, "bounding box max must exceed min on every axis");
}
}  // namespace
}  // namespace sdfb200

using namespace sdfb200;

struct sdfb200_mesh { std::shared_ptr<PreparedMesh> pm; };

namespace {
void checkOctreeOptions(int initAlgorithm, int terminationRule) {
    if (initAlgorithm != SDFB200_ALG_NO_CONTINUITY && initAlgorithm != SDFB200_ALG_CONTINUITY)
        throw Error(SDFB200_ERR_UNSUPPORTED, "InitAlgorithm::UNIFORM (the reference's testing variant) is not built: see DESIGN.md");
    if (terminationRule < SDFB200_RULE_NONE || terminationRule > SDFB200_RULE_BY_DISTANCE)
        throw Error(SDFB200_ERR_INVALID, "unknown termination rule");
}
const PreparedMesh& meshOnCurrentDevice(const sdfb200_mesh* m) {
    if (!m || !m->pm) throw Error(SDFB200_ERR_INVALID, "null mesh handle");
    requireDevice();
    SDFB_CUDA(cudaSetDevice(m->pm->device));
    return *m->pm;
}
std::vector<int> checkedDevices(const int* devices, uint32_t nDevices) {
    if (!devices || nDevices == 0) throw Error(SDFB200_ERR_INVALID, "empty device list");
    const int n = sdfb200_device_count();
    if (n == 0) throw Error(SDFB200_ERR_CUDA, "no CUDA device available (sdfb200 has no CPU fallback)");
    std::vector<int> d(devices, devices + nDevices);
    // test switch: several "ranks" on one device exercise the thread choreography on a single-GPU box (peer copies, no NCCL)
    const char* dup = std::getenv("SDFB200_ALLOW_DUPLICATE_DEVICES");
    for (size_t i = 0; i < d.size(); i++) {
        if (d[i] < 0 || d[i] >= n) throw Error(SDFB200_ERR_INVALID, "device index out of range");
        for (size_t j = 0; j < i; j++)
            if (d[j] == d[i] && !(dup && dup[0] == '1')) throw Error(SDFB200_ERR_INVALID, "device listed twice");
    }
    return d;
}
void runMulti(const HostMesh& mesh, const MultiBuildRequest& req, const int* devices, uint32_t nDevices, sdfb200_sdf** outHandles) {
    if (!outHandles) throw Error(SDFB200_ERR_INVALID, "null output handles");
    for (uint32_t k = 0; k < nDevices; k++) outHandles[k] = nullptr;
    const std::vector<int> dev = checkedDevices(devices, nDevices);
    std::vector<std::unique_ptr<sdfb200_sdf>> built;
    buildMulti(mesh, req, dev, built);
    for (uint32_t k = 0; k < nDevices; k++) outHandles[k] = built[k].release();
}
}  // namespace

extern "C" {

// ---- prepared meshes ---------------------------------------------------------------------------------------------
int sdfb200_mesh_create(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices, int parts,
                        sdfb200_mesh** out) {
    return guarded("sdfb200:mesh_create", [&] {
        if (!out) throw Error(SDFB200_ERR_INVALID, "null output handle");
        *out = nullptr;
        HostMesh mesh = checkedMesh(vertices, numVertices, indices, numIndices);
        requireDevice();
        if (parts & SDFB200_MESH_ALL_HOST_THREADS) setHostThreadsForThisThread(int(std::thread::hardware_concurrency()));
        std::unique_ptr<sdfb200_mesh> m(new sdfb200_mesh());
        try { m->pm = prepareMesh(mesh, (parts & SDFB200_MESH_BVH) != 0, (parts & SDFB200_MESH_EXACT) != 0); }
        catch (...) { setHostThreadsForThisThread(0); throw; }
        setHostThreadsForThisThread(0);
        *out = m.release();
    });
}

void sdfb200_mesh_free(sdfb200_mesh* mesh) {
    if (!mesh) return;
    if (mesh->pm) { cudaSetDevice(mesh->pm->device); cudaDeviceSynchronize(); }
    delete mesh;
}

int sdfb200_mesh_blob_bytes(const sdfb200_mesh* mesh, uint64_t* outBytes) {
    return guarded([&] {
        if (!mesh || !mesh->pm || !outBytes) throw Error(SDFB200_ERR_INVALID, "null argument");
        *outBytes = meshBlobBytes(*mesh->pm);
    });
}

int sdfb200_mesh_export(const sdfb200_mesh* mesh, void* devicePtr, uint64_t capacityBytes) {
    return guarded("sdfb200:mesh_export", [&] {
        if (!devicePtr) throw Error(SDFB200_ERR_INVALID, "null argument");
        meshBlobExport(meshOnCurrentDevice(mesh), devicePtr, capacityBytes, cudaStream_t(0));
    });
}

int sdfb200_mesh_import(const void* devicePtr, uint64_t bytes, sdfb200_mesh** out) {
    return guarded("sdfb200:mesh_import", [&] {
        if (!devicePtr || !out) throw Error(SDFB200_ERR_INVALID, "null argument");
        *out = nullptr;
        requireDevice();
        std::unique_ptr<sdfb200_mesh> m(new sdfb200_mesh());
        m->pm = meshBlobImport(devicePtr, bytes);
        *out = m.release();
    });
}

int sdfb200_mesh_stats(const sdfb200_mesh* mesh, double* triangleDataMs, double* bvhMs, double* uploadMs) {
    return guarded([&] {
        if (!mesh || !mesh->pm) throw Error(SDFB200_ERR_INVALID, "null mesh handle");
        if (triangleDataMs) *triangleDataMs = mesh->pm->triangleDataMs;
        if (bvhMs) *bvhMs = mesh->pm->bvhMs;
        if (uploadMs) *uploadMs = mesh->pm->uploadMs;
    });
}

int sdfb200_build_octree_from_mesh(const sdfb200_mesh* mesh, const float* box6, uint32_t depth, uint32_t startDepth, int terminationRule,
                                   float param0, float param1, int initAlgorithm, uint32_t numThreads, uint32_t rank, uint32_t worldSize,
                                   sdfb200_sdf** out) {
    return guarded("sdfb200:build_octree_from_mesh", [&] {
        if (!out) throw Error(SDFB200_ERR_INVALID, "null output handle");
        *out = nullptr;
        checkBox(box6);
        if (worldSize == 0 || rank >= worldSize) throw Error(SDFB200_ERR_INVALID, "rank/worldSize out of range");
        checkOctreeOptions(initAlgorithm, terminationRule);
        if (initAlgorithm == SDFB200_ALG_CONTINUITY && worldSize > 1)
            throw Error(SDFB200_ERR_UNSUPPORTED, "CONTINUITY does not shard by start voxels (its neighbour probes cross them): use sdfb200_build_octree_collective_from_mesh");
        const PreparedMesh& pm = meshOnCurrentDevice(mesh);
        std::unique_ptr<sdfb200_sdf> s(new sdfb200_sdf());
        if (initAlgorithm == SDFB200_ALG_CONTINUITY) buildOctreeContinuityOnDevice(*s, pm, box6, depth, startDepth, terminationRule, param0, param1);
        else buildOctreeOnDevice(*s, pm, box6, depth, startDepth, terminationRule, param0, param1, numThreads, rank, worldSize);
        *out = s.release();
    });
}

int sdfb200_build_octree_collective_from_mesh(const sdfb200_mesh* mesh, const float* box6, uint32_t depth, uint32_t startDepth,
                                              int terminationRule, float param0, float param1, uint32_t rank, uint32_t worldSize,
                                              sdfb200_allgather_fn allgather, void* user, sdfb200_sdf** out) {
    return guarded("sdfb200:build_octree_collective_from_mesh", [&] {
        if (!out) throw Error(SDFB200_ERR_INVALID, "null output handle");
        *out = nullptr;
        checkBox(box6);
        if (worldSize == 0 || rank >= worldSize) throw Error(SDFB200_ERR_INVALID, "rank/worldSize out of range");
        checkOctreeOptions(SDFB200_ALG_CONTINUITY, terminationRule);
        if (worldSize > 1 && !allgather) throw Error(SDFB200_ERR_INVALID, "worldSize > 1 needs an allgather hook");
        const PreparedMesh& pm = meshOnCurrentDevice(mesh);
        std::unique_ptr<sdfb200_sdf> s(new sdfb200_sdf());
        SampleExchange ex;
        ex.rank = rank; ex.world = worldSize; ex.allgather = allgather; ex.user = user;
        buildOctreeContinuityOnDevice(*s, pm, box6, depth, startDepth, terminationRule, param0, param1, ex);
        *out = s.release();
    });
}

int sdfb200_build_exact_from_mesh(const sdfb200_mesh* mesh, const float* box6, uint32_t maxDepth, uint32_t startDepth,
                                  uint32_t minTrianglesPerNode, uint32_t numThreads, uint32_t rank, uint32_t worldSize, sdfb200_sdf** out) {
    return guarded("sdfb200:build_exact_from_mesh", [&] {
        if (!out) throw Error(SDFB200_ERR_INVALID, "null output handle");
        *out = nullptr;
        checkBox(box6);
        if (worldSize == 0 || rank >= worldSize) throw Error(SDFB200_ERR_INVALID, "rank/worldSize out of range");
        meshOnCurrentDevice(mesh);
        std::unique_ptr<sdfb200_sdf> s(new sdfb200_sdf());
        buildExactOnDevice(*s, mesh->pm, box6, maxDepth, startDepth, minTrianglesPerNode, numThreads, rank, worldSize);
        *out = s.release();
    });
}

// ---- one process, several devices ----------------------------------------------------------------------------------
int sdfb200_nccl_available(void) { return ncclAvailable() ? 1 : 0; }

int sdfb200_build_octree_multi(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                               const float* box6, uint32_t depth, uint32_t startDepth, int terminationRule, float param0, float param1,
                               int initAlgorithm, uint32_t numThreads, const int* devices, uint32_t nDevices, sdfb200_sdf** outHandles) {
    return guarded("sdfb200:build_octree_multi", [&] {
        HostMesh mesh = checkedMesh(vertices, numVertices, indices, numIndices);
        checkBox(box6);
        checkOctreeOptions(initAlgorithm, terminationRule);
        MultiBuildRequest req{SDFB200_FORMAT_OCTREE, box6, depth, startDepth, numThreads, terminationRule, param0, param1, initAlgorithm, 0};
        runMulti(mesh, req, devices, nDevices, outHandles);
    });
}

int sdfb200_build_exact_multi(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                              const float* box6, uint32_t maxDepth, uint32_t startDepth, uint32_t minTrianglesPerNode, uint32_t numThreads,
                              const int* devices, uint32_t nDevices, sdfb200_sdf** outHandles) {
    return guarded("sdfb200:build_exact_multi", [&] {
        HostMesh mesh = checkedMesh(vertices, numVertices, indices, numIndices);
        checkBox(box6);
        MultiBuildRequest req{SDFB200_FORMAT_EXACT_OCTREE, box6, maxDepth, startDepth, numThreads, 0, 0.0f, 0.0f, 0, minTrianglesPerNode};
        runMulti(mesh, req, devices, nDevices, outHandles);
    });
}

const char* sdfb200_last_error(void) { return gLastError.c_str(); }
int sdfb200_version(void) { return SDFB200_VERSION; }

int sdfb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int sdfb200_set_device(int device) {
    return guarded([&] {
        int n = sdfb200_device_count();
        if (device < 0 || device >= n) throw Error(SDFB200_ERR_INVALID, "device index out of range");
        gDevice = device;
        SDFB_CUDA(cudaSetDevice(device));
    });
}

int sdfb200_release_cached_memory(void) {
    return guarded([&] { releaseCachedMemory(); });
}

int sdfb200_build_octree_shard(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                               const float* box6, uint32_t depth, uint32_t startDepth, int terminationRule, float param0,
                               float param1, int initAlgorithm, uint32_t numThreads, uint32_t rank, uint32_t worldSize,
                               sdfb200_sdf** out) {
    return guarded("sdfb200:build_octree_shard", [&] {
        if (!out) throw Error(SDFB200_ERR_INVALID, "null output handle");
        *out = nullptr;
        HostMesh mesh = checkedMesh(vertices, numVertices, indices, numIndices);
        checkBox(box6);
        if (worldSize == 0 || rank >= worldSize) throw Error(SDFB200_ERR_INVALID, "rank/worldSize out of range");
        if (initAlgorithm != SDFB200_ALG_NO_CONTINUITY && initAlgorithm != SDFB200_ALG_CONTINUITY)
            throw Error(SDFB200_ERR_UNSUPPORTED, "InitAlgorithm::UNIFORM (the reference's testing variant) is not built: see DESIGN.md");
        if (terminationRule < SDFB200_RULE_NONE || terminationRule > SDFB200_RULE_BY_DISTANCE)
            throw Error(SDFB200_ERR_INVALID, "unknown termination rule");
        if (initAlgorithm == SDFB200_ALG_CONTINUITY && worldSize > 1)
            throw Error(SDFB200_ERR_UNSUPPORTED, "CONTINUITY does not shard by start voxels (its neighbour probes cross them): use sdfb200_build_octree_collective");
        requireDevice();
        const auto tStart = std::chrono::steady_clock::now();
        const std::shared_ptr<PreparedMesh> pm = prepareMesh(mesh, true, false);
        const double prepMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tStart).count();
        std::unique_ptr<sdfb200_sdf> s(new sdfb200_sdf());
        if (initAlgorithm == SDFB200_ALG_CONTINUITY)
            buildOctreeContinuityOnDevice(*s, *pm, box6, depth, startDepth, terminationRule, param0, param1);
        else
            buildOctreeOnDevice(*s, *pm, box6, depth, startDepth, terminationRule, param0, param1, numThreads, rank, worldSize);
        s->stats.total_ms += prepMs;
        *out = s.release();
    });
}

int sdfb200_build_octree_collective(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                                    const float* box6, uint32_t depth, uint32_t startDepth, int terminationRule, float param0,
                                    float param1, int initAlgorithm, uint32_t numThreads, uint32_t rank, uint32_t worldSize,
                                    sdfb200_allgather_fn allgather, void* user, sdfb200_sdf** out) {
    (void)numThreads;   // the CONTINUITY layout does not depend on it
    return guarded("sdfb200:build_octree_collective", [&] {
        if (!out) throw Error(SDFB200_ERR_INVALID, "null output handle");
        *out = nullptr;
        HostMesh mesh = checkedMesh(vertices, numVertices, indices, numIndices);
        checkBox(box6);
        if (worldSize == 0 || rank >= worldSize) throw Error(SDFB200_ERR_INVALID, "rank/worldSize out of range");
        if (initAlgorithm != SDFB200_ALG_CONTINUITY)
            throw Error(SDFB200_ERR_INVALID, "sdfb200_build_octree_collective builds InitAlgorithm::CONTINUITY (NO_CONTINUITY shards by start voxels: sdfb200_build_octree_shard)");
        if (terminationRule < SDFB200_RULE_NONE || terminationRule > SDFB200_RULE_BY_DISTANCE)
            throw Error(SDFB200_ERR_INVALID, "unknown termination rule");
        if (worldSize > 1 && !allgather) throw Error(SDFB200_ERR_INVALID, "worldSize > 1 needs an allgather hook");
        requireDevice();
        const auto tStart = std::chrono::steady_clock::now();
        const std::shared_ptr<PreparedMesh> pm = prepareMesh(mesh, true, false);
        const double prepMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tStart).count();
        std::unique_ptr<sdfb200_sdf> s(new sdfb200_sdf());
        SampleExchange ex;
        ex.rank = rank; ex.world = worldSize; ex.allgather = allgather; ex.user = user;
        buildOctreeContinuityOnDevice(*s, *pm, box6, depth, startDepth, terminationRule, param0, param1, ex);
        s->stats.total_ms += prepMs;
        *out = s.release();
    });
}

int sdfb200_build_octree(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                         const float* box6, uint32_t depth, uint32_t startDepth, int terminationRule, float param0,
                         float param1, int initAlgorithm, uint32_t numThreads, sdfb200_sdf** out) {
    return sdfb200_build_octree_shard(vertices, numVertices, indices, numIndices, box6, depth, startDepth, terminationRule,
                                      param0, param1, initAlgorithm, numThreads, 0, 1, out);
}

int sdfb200_build_exact_shard(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                              const float* box6, uint32_t maxDepth, uint32_t startDepth, uint32_t minTrianglesPerNode,
                              uint32_t numThreads, uint32_t rank, uint32_t worldSize, sdfb200_sdf** out) {
    return guarded("sdfb200:build_exact_shard", [&] {
        if (!out) throw Error(SDFB200_ERR_INVALID, "null output handle");
        *out = nullptr;
        HostMesh mesh = checkedMesh(vertices, numVertices, indices, numIndices);
        checkBox(box6);
        if (worldSize == 0 || rank >= worldSize) throw Error(SDFB200_ERR_INVALID, "rank/worldSize out of range");
        requireDevice();
        const auto tStart = std::chrono::steady_clock::now();
        const std::shared_ptr<PreparedMesh> pm = prepareMesh(mesh, false, true);
        const double prepMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tStart).count();
        std::unique_ptr<sdfb200_sdf> s(new sdfb200_sdf());
        buildExactOnDevice(*s, pm, box6, maxDepth, startDepth, minTrianglesPerNode, numThreads, rank, worldSize);
        s->stats.total_ms += prepMs;
        *out = s.release();
    });
}

int sdfb200_build_exact(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                        const float* box6, uint32_t maxDepth, uint32_t startDepth, uint32_t minTrianglesPerNode,
                        uint32_t numThreads, sdfb200_sdf** out) {
    return sdfb200_build_exact_shard(vertices, numVertices, indices, numIndices, box6, maxDepth, startDepth, minTrianglesPerNode,
                                     numThreads, 0, 1, out);
}

int sdfb200_save(const sdfb200_sdf* sdf, const char* path) {
    return guarded("sdfb200:save", [&] {
        if (!sdf || !path) throw Error(SDFB200_ERR_INVALID, "null argument");
        if (sdf->isShard) throw Error(SDFB200_ERR_INVALID, "handle is an unassembled shard (sdfb200_assemble not called yet)");
        ensureHostMirror(*const_cast<sdfb200_sdf*>(sdf));
        saveBin(*sdf, path);
    });
}

int sdfb200_load(const char* path, sdfb200_sdf** out) {
    return guarded("sdfb200:load", [&] {
        if (!out || !path) throw Error(SDFB200_ERR_INVALID, "null argument");
        *out = nullptr;
        std::unique_ptr<sdfb200_sdf> s(new sdfb200_sdf());
        loadBin(*s, path);   // file errors are reported even without a GPU
        requireDevice();
        uploadStructure(*s);
        *out = s.release();
    });
}

void sdfb200_free(sdfb200_sdf* sdf) {
    if (!sdf) return;
    if (sdf->dOctree.p) { cudaSetDevice(sdf->device); cudaDeviceSynchronize(); }   // no kernel may still read the arrays
    delete sdf;
}

int sdfb200_get_info(const sdfb200_sdf* s, sdfb200_info* o) {
    return guarded([&] {
        if (!s || !o) throw Error(SDFB200_ERR_INVALID, "null argument");
        std::memset(o, 0, sizeof(*o));
        o->format = s->format;
        for (int i = 0; i < 3; i++) { o->box_min[i] = s->boxMin[i]; o->box_max[i] = s->boxMax[i]; }
        o->start_grid_size = s->startGridSize;
        o->max_depth = s->maxDepth;
        o->value_range = s->valueRange;
        o->min_border_value = s->minBorderValue;
        o->start_depth = s->startDepth;
        o->min_triangles_in_leafs = s->minTrisInLeafs;
        o->max_triangles_in_leafs = s->maxTrisInLeafs;
        o->max_triangles_encoded_in_leafs = s->maxTrisEncoded;
        o->bit_encoding_start_depth = s->bitEncodingStartDepth;
        o->bits_per_index = s->bitsPerIndex;
        o->octree_words = s->format == SDFB200_FORMAT_OCTREE ? s->nOctree : s->nOctree / 2;
        o->triangle_sets_words = s->nSets;
        o->triangle_masks_bytes = s->nMasks;
        o->num_triangles = s->format == SDFB200_FORMAT_EXACT_OCTREE ? s->numTris : 0;
        o->device = s->device;
    });
}

int sdfb200_get_build_stats(const sdfb200_sdf* s, sdfb200_build_stats* o) {
    return guarded([&] {
        if (!s || !o) throw Error(SDFB200_ERR_INVALID, "null argument");
        *o = s->stats;
    });
}

int sdfb200_get_octree_data(const sdfb200_sdf* s, uint32_t* out, uint64_t capacityWords) {
    return guarded([&] {
        if (!s || !out) throw Error(SDFB200_ERR_INVALID, "null argument");
        if (s->isShard) throw Error(SDFB200_ERR_INVALID, "handle is an unassembled shard (sdfb200_assemble not called yet)");
        if (capacityWords < s->nOctree) throw Error(SDFB200_ERR_INVALID, "output buffer too small");
        ensureHostMirror(*const_cast<sdfb200_sdf*>(s));
        std::memcpy(out, s->octree.data(), s->nOctree * sizeof(uint32_t));
    });
}

int sdfb200_get_exact_arrays(const sdfb200_sdf* s, uint32_t* sets, uint8_t* masks, float* tris37) {
    return guarded([&] {
        if (!s) throw Error(SDFB200_ERR_INVALID, "null argument");
        if (s->format != SDFB200_FORMAT_EXACT_OCTREE) throw Error(SDFB200_ERR_INVALID, "not an ExactOctreeSdf");
        if (sets || masks) ensureHostMirror(*const_cast<sdfb200_sdf*>(s));
        if (sets) std::memcpy(sets, s->sets.data(), s->nSets * 4);
        if (masks) std::memcpy(masks, s->masks.data(), s->nMasks);
        if (tris37) { const TriVec& t = const_cast<sdfb200_sdf*>(s)->hostTris(); std::memcpy(tris37, t.data(), t.size() * sizeof(TriData)); }
    });
}

int sdfb200_get_device_octree(const sdfb200_sdf* s, const uint32_t** outDevicePtr) {
    return guarded([&] {
        if (!s || !outDevicePtr) throw Error(SDFB200_ERR_INVALID, "null argument");
        *outDevicePtr = s->dOctree.p;
    });
}

int sdfb200_query(sdfb200_sdf* s, const float* xyz, uint64_t n, float* dist, float* grad, int flags, void* cudaStream) {
    return guarded("sdfb200:query", [&] {
        if (!s || (n && (!xyz || !dist))) throw Error(SDFB200_ERR_INVALID, "null argument");
        if (n == 0) return;
        if (!s->dOctree.p) throw Error(SDFB200_ERR_CUDA, "structure is not resident on a CUDA device");
        if (s->isShard) throw Error(SDFB200_ERR_INVALID, "handle is an unassembled shard (sdfb200_assemble not called yet)");
        SDFB_CUDA(cudaSetDevice(s->device));
        cudaStream_t st = static_cast<cudaStream_t>(cudaStream);
        if (flags & SDFB200_QUERY_DEVICE_POINTERS) queryDevicePointers(*s, xyz, n, dist, grad, flags, st);
        else queryHostPointers(*s, xyz, n, dist, grad, flags, st);
    });
}

int sdfb200_sphere_trace(sdfb200_sdf* s, const float* origins, const float* directions, uint64_t n, float epsilon, float farDistance,
                         uint32_t maxIterations, float* outHit, float* outTravelled, uint32_t* outIterations, int flags, void* cudaStream) {
    return guarded("sdfb200:sphere_trace", [&] {
        if (!s || (n && (!origins || !directions || !outHit || !outTravelled))) throw Error(SDFB200_ERR_INVALID, "null argument");
        if (n == 0) return;
        if (s->format != SDFB200_FORMAT_OCTREE) throw Error(SDFB200_ERR_UNSUPPORTED, "sphere tracing is built for OctreeSdf structures");
        if (!s->dOctree.p) throw Error(SDFB200_ERR_CUDA, "structure is not resident on a CUDA device");
        if (s->isShard) throw Error(SDFB200_ERR_INVALID, "handle is an unassembled shard (sdfb200_assemble not called yet)");
        SDFB_CUDA(cudaSetDevice(s->device));
        cudaStream_t st = static_cast<cudaStream_t>(cudaStream);
        auto launch = [&](const float* dO, const float* dD, float* dH, float* dT, uint32_t* dI) {
            if (flags & SDFB200_QUERY_EXACT_ORDER) launchOctreeTraceExact(*s, dO, dD, n, epsilon, farDistance, maxIterations, dH, dT, dI, st);
            else launchOctreeTraceFast(*s, dO, dD, n, epsilon, farDistance, maxIterations, dH, dT, dI, st);
        };
        if (flags & SDFB200_QUERY_DEVICE_POINTERS) { launch(origins, directions, outHit, outTravelled, outIterations); return; }
        DevBuf<float> dO(3 * n), dD(3 * n), dH(3 * n), dT(n);
        DevBuf<uint32_t> dI(outIterations ? n : 0);
        SDFB_CUDA(cudaStreamSynchronize(cudaStream_t(0)));   // device blocks are ordered on the default stream
        SDFB_CUDA(cudaMemcpyAsync(dO.p, origins, 3 * n * sizeof(float), cudaMemcpyHostToDevice, st));
        SDFB_CUDA(cudaMemcpyAsync(dD.p, directions, 3 * n * sizeof(float), cudaMemcpyHostToDevice, st));
        launch(dO.p, dD.p, dH.p, dT.p, outIterations ? dI.p : nullptr);
        SDFB_CUDA(cudaMemcpyAsync(outHit, dH.p, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, st));
        SDFB_CUDA(cudaMemcpyAsync(outTravelled, dT.p, n * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (outIterations) SDFB_CUDA(cudaMemcpyAsync(outIterations, dI.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        SDFB_CUDA(cudaStreamSynchronize(st));
    });
}

int sdfb200_shard_sizes(const sdfb200_sdf* s, uint32_t* outSizes, uint64_t capacity, uint64_t* outCount) {
    return guarded([&] {
        if (!s || !outCount) throw Error(SDFB200_ERR_INVALID, "null argument");
        *outCount = s->shardSizes.size();
        if (!outSizes) return;   // size query
        if (capacity < s->shardSizes.size()) throw Error(SDFB200_ERR_INVALID, "output buffer too small");
        std::memcpy(outSizes, s->shardSizes.data(), s->shardSizes.size() * sizeof(uint32_t));
    });
}

int sdfb200_shard_finish(sdfb200_sdf* s, const uint32_t* allSizes, uint64_t count) {
    return guarded("sdfb200:shard_finish", [&] {
        if (!s || !allSizes) throw Error(SDFB200_ERR_INVALID, "null argument");
        if (!s->build) throw Error(SDFB200_ERR_INVALID, "handle is not a shard in phase 1 (built with worldSize > 1)");
        if (count != s->shardSizes.size()) throw Error(SDFB200_ERR_INVALID, "size vector length differs from sdfb200_shard_sizes");
        for (uint64_t i = 0; i < count; i++)
            if (s->shardSizes[i] && s->shardSizes[i] != allSizes[i])
                throw Error(SDFB200_ERR_INVALID, "all-reduced sizes disagree with this rank's own roots (ownership overlap?)");
        SDFB_CUDA(cudaSetDevice(s->device));
        s->build->finish(*s, allSizes);
        settleDeviceCache(s->device);
    });
}

int sdfb200_shard_words(const sdfb200_sdf* s, uint64_t* outWords) {
    return guarded([&] {
        if (!s || !outWords) throw Error(SDFB200_ERR_INVALID, "null argument");
        *outWords = shardPayloadWords(*s);
    });
}

int sdfb200_shard_export(const sdfb200_sdf* s, uint32_t* devicePtr, uint64_t capacityWords) {
    return guarded("sdfb200:shard_export", [&] {
        if (!s || !devicePtr) throw Error(SDFB200_ERR_INVALID, "null argument");
        SDFB_CUDA(cudaSetDevice(s->device));
        shardExport(*s, devicePtr, capacityWords);
    });
}

int sdfb200_assemble(sdfb200_sdf* s, const uint32_t* gatheredDevicePtr, const uint64_t* wordsPerRank, uint64_t strideWords,
                     uint32_t worldSize) {
    return guarded("sdfb200:assemble", [&] {
        if (!s || !gatheredDevicePtr || !wordsPerRank) throw Error(SDFB200_ERR_INVALID, "null argument");
        SDFB_CUDA(cudaSetDevice(s->device));
        shardAssemble(*s, gatheredDevicePtr, wordsPerRank, strideWords, worldSize);
    });
}

int sdfb200_triangle_data(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices, float* out37) {
    return guarded([&] {
        if (!out37) throw Error(SDFB200_ERR_INVALID, "null output");
        HostMesh mesh = checkedMesh(vertices, numVertices, indices, numIndices);
        TriVec t = computeTriangleData(mesh);
        std::memcpy(out37, t.data(), t.size() * sizeof(TriData));
    });
}

int sdfb200_bvh_host(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices, void* outNodes, uint64_t capacityNodes) {
    return guarded([&] {
        HostMesh mesh = checkedMesh(vertices, numVertices, indices, numIndices);
        const RawVec<BvhNode> nodes = buildBvh(mesh);
        if (!outNodes || capacityNodes < nodes.size()) throw Error(SDFB200_ERR_INVALID, "the BVH has 2 * triangles - 1 nodes of 80 bytes");
        std::memcpy(outNodes, nodes.data(), nodes.size() * sizeof(BvhNode));
    });
}

int sdfb200_mesh_bvh(const sdfb200_mesh* mesh, void* outNodes, uint64_t capacityNodes) {
    return guarded([&] {
        if (!mesh || !mesh->pm) throw Error(SDFB200_ERR_INVALID, "null mesh handle");
        const PreparedMesh& pm = *mesh->pm;
        if (!pm.hasBvh) throw Error(SDFB200_ERR_INVALID, "the mesh was prepared without SDFB200_MESH_BVH");
        if (!outNodes || capacityNodes < pm.dev.bvh.n) throw Error(SDFB200_ERR_INVALID, "the BVH has 2 * triangles - 1 nodes of 80 bytes");
        int current = 0;
        SDFB_CUDA(cudaGetDevice(&current));
        SDFB_CUDA(cudaSetDevice(pm.device));
        const cudaError_t e = cudaMemcpy(outNodes, pm.dev.bvh.p, pm.dev.bvh.n * sizeof(BvhNode), cudaMemcpyDeviceToHost);
        cudaSetDevice(current);
        SDFB_CUDA(e);
    });
}

int sdfb200_nearest_triangle(const float* vertices, uint32_t numVertices, const uint32_t* indices, uint32_t numIndices,
                             const float* xyz, uint64_t n, uint32_t* outTriangle) {
    return guarded([&] {
        if (n && (!xyz || !outTriangle)) throw Error(SDFB200_ERR_INVALID, "null argument");
        HostMesh mesh = checkedMesh(vertices, numVertices, indices, numIndices);
        requireDevice();
        nearestTriangleOnDevice(mesh, xyz, n, outTriangle);
    });
}

int sdfb200_point_triangle(const float* tri37, const float* v123, const float* xyz, uint64_t n, int mode, float* outDist,
                           float* outGrad) {
    return guarded([&] {
        if (!tri37 || (n && (!xyz || !outDist))) throw Error(SDFB200_ERR_INVALID, "null argument");
        requireDevice();
        pointTriangleOnDevice(tri37, v123, xyz, n, mode, outDist, outGrad);
    });
}

}  // extern "C"
