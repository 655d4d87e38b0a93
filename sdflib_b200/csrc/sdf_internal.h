// Internal definitions shared by the translation units of libsdfb200.so (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sdfb200.h"
#include "mesh_host.h"
#include "tri_math.cuh"

namespace sdfb200 {

// ---- errors ------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
void setLastError(const std::string& msg);

#define SDFB_CUDA(expr)                                                                                   \
    do {                                                                                                  \
        cudaError_t _e = (expr);                                                                          \
        if (_e != cudaSuccess)                                                                            \
            throw ::sdfb200::Error(SDFB200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// ---- NVTX ranges ---------------------------------------------------------------------------------
// Entry points and build phases show up as named ranges ("sdfb200:...") in Nsight Systems / Compute timelines. nvtx3 is
// header-only and resolves the tool's injection library at the first call: without a tool attached a range costs a
// function-pointer test.
}  // namespace sdfb200
#ifndef SDFB_NO_NVTX
#include <nvtx3/nvToolsExt.h>
#define SDFB_NVTX_PUSH(name) nvtxRangePushA(name)
#define SDFB_NVTX_POP() nvtxRangePop()
#else
#define SDFB_NVTX_PUSH(name) ((void)0)
#define SDFB_NVTX_POP() ((void)0)
#endif
namespace sdfb200 {
struct NvtxRange {   // scoped range; next() closes the current phase and opens another one
    explicit NvtxRange(const char* name) { SDFB_NVTX_PUSH(name); }
    void next(const char* name) { SDFB_NVTX_POP(); SDFB_NVTX_PUSH(name); }
    ~NvtxRange() { SDFB_NVTX_POP(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// ---- device buffer -----------------------------------------------------------------------------
// Device blocks: rounded to a size class and recycled through per-class free lists. Every kernel that touches a
// DevBuf runs on the legacy default stream (or is synchronised before the buffer is released), so handing a freed
// block to the next allocation is ordered by the stream itself.
void* deviceBlockAlloc(size_t bytes, size_t* outCapacity);
void deviceBlockFree(void* p, size_t capacity);
void settleDeviceCache(int device);          // end of a top-level build (see host_mem.cpp)
struct DeviceCacheSettle {                  // declare BEFORE the build state it should outlive
    int device;
    explicit DeviceCacheSettle(int d) : device(d) {}
    ~DeviceCacheSettle() { settleDeviceCache(device); }
};
void setDeviceBlockStream(cudaStream_t s);   // thread-local: stream that cache misses / overflows are ordered on (default: legacy stream)

template <class T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0, capBytes = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), capBytes(o.capBytes) { o.p = nullptr; o.n = 0; o.capBytes = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { release(); p = o.p; n = o.n; capBytes = o.capBytes; o.p = nullptr; o.n = 0; o.capBytes = 0; return *this; }
    ~DevBuf() { release(); }
    // device blocks come from a process-wide size-class cache on top of the stream-ordered allocator (host_mem.cpp)
    void alloc(size_t count) {
        release();
        n = count;
        if (count) p = static_cast<T*>(deviceBlockAlloc(count * sizeof(T), &capBytes));
    }
    // grow-only variant for temporaries reused across the levels of a build: contents are not preserved, `n` is the
    // capacity. Re-allocating a slightly larger block every level makes the pool grow (slow) instead of recycling.
    void ensure(size_t count) {
        if (count <= n) return;
        alloc(count + count / 2 + 256);
    }
    void release() { if (p) deviceBlockFree(p, capBytes); p = nullptr; n = 0; capBytes = 0; }
    void upload(const T* src, size_t count, cudaStream_t s = 0) {
        if (count) SDFB_CUDA(cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void download(T* dst, size_t count, cudaStream_t s = 0) const {
        if (count) SDFB_CUDA(cudaMemcpyAsync(dst, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
};

// ---- host mirror of a structure array: uninitialised, pinned when a device is present (host_mem.cpp) ----
void* hostBlockAlloc(size_t bytes, size_t* outCapacity, bool* outPinned);
void hostBlockFree(void* p, size_t capacity, bool pinned);
void configureDevicePool(int device);
void releaseCachedMemory();

template <class T> struct HostArray {
    T* p = nullptr;
    size_t n = 0, capBytes = 0;
    bool pinned = false;
    HostArray() {}
    HostArray(const HostArray&) = delete;
    HostArray& operator=(const HostArray&) = delete;
    ~HostArray() { hostBlockFree(p, capBytes, pinned); }
    // contents are NOT preserved and NOT initialised
    void resize(size_t count) {
        if (count * sizeof(T) > capBytes || !p) {
            hostBlockFree(p, capBytes, pinned);
            p = nullptr;
            p = static_cast<T*>(hostBlockAlloc(count * sizeof(T), &capBytes, &pinned));
        }
        n = count;
    }
    T* data() { return p; }
    const T* data() const { return p; }
    size_t size() const { return n; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
};

// ---- mesh on the device --------------------------------------------------------------------------
struct DeviceMesh {
    const f3* verts;
    const uint32_t* idx;
    const TriData* tris;
    const BvhNode* bvh;   // may be null (exact builder does not need it)
    uint32_t numTriangles;
    const float4* triVerts;   // with bvh: the 3 vertices of every triangle, pre-gathered (one 48-byte record instead of 3 + 3 gathers)
    int rootLink;             // with bvh: link of the root (0, or ~triangleId for a single-triangle mesh)
    int stackDepth;           // with bvh: entries of the per-thread traversal stack (tree height + 1)
};

struct MeshOnDevice {
    DevBuf<f3> verts;
    DevBuf<uint32_t> idx;
    DevBuf<TriData> tris;
    DevBuf<BvhNode> bvh;
    DevBuf<float4> triVerts;
    uint32_t numTriangles = 0;
    int rootLink = 0, stackDepth = 1;
    DeviceMesh view() const { return DeviceMesh{verts.p, idx.p, tris.p, bvh.p, numTriangles, triVerts.p, rootLink, stackDepth}; }
};

// ---- prepared mesh (mesh_device.cu): what every builder reads; computed once per mesh and device ------------------------
struct PreparedMesh {
    int device = 0;
    uint32_t nVerts = 0, nIdx = 0, nTris = 0;
    MeshOnDevice dev;                 // verts, idx, TriangleData; with hasBvh also bvh, triVerts, rootLink, stackDepth
    bool hasBvh = false;              // OctreeSdf builders (nearest-triangle queries)
    bool hasExactParts = false;       // ExactOctreeSdf builder / query: 80-byte frames + ids of the non-degenerate triangles
    DevBuf<float4> frames;
    DevBuf<uint32_t> valid;           // ascending triangle ids passing ExactOctreeSdfDepthFirst.h:106, padded by 8 words
    uint32_t numValid = 0;
    double triangleDataMs = 0, bvhMs = 0, uploadMs = 0;   // host wall clock of the phases (0 for clones / imports)
    std::mutex lazy;
    TriVec hostTris;                  // host copy of TriangleData, fetched from the device on first use
    const TriVec& hostTriangleData();
};
std::shared_ptr<PreparedMesh> prepareMesh(const HostMesh& mesh, bool withBvh, bool withExactParts);
void attachBvh(PreparedMesh& pm, const RawVec<BvhNode>& bvh, double bvhMs);   // a BVH built once on the host, uploaded to pm's device
void buildBvhOnDevice(MeshOnDevice& m, const HostMesh* host = nullptr);   // bvh_device.cu: m.triVerts in place -> m.bvh, rootLink, stackDepth (synchronous); host: the caller's arrays, for the centre sums of the top levels
bool hostBvhRequested();                  // SDFB200_HOST_BVH: the host builder of mesh_host.cpp instead (A/B)
uint64_t meshBlobBytes(const PreparedMesh& pm);
void meshBlobExport(const PreparedMesh& pm, void* dDst, uint64_t capacity, cudaStream_t st);
std::shared_ptr<PreparedMesh> meshBlobImport(const void* dSrc, uint64_t bytes);
std::shared_ptr<PreparedMesh> cloneMeshToCurrentDevice(const PreparedMesh& src);

// ---- sharded construction: state shared by both builders ---------------------------------
// Roots = nodes of the start depth. Their order in the output arrays follows the reference's drivers
// (numThreads < 2: one global stack, virtual levels popped 7-first; numThreads >= 2: start-grid order); the owner of a
// root is chosen by estimated work (longest-processing-time greedy over per-root weights). Identical on every rank.
struct RootPlan {
    uint32_t G3 = 0, world = 1, rank = 0;
    std::vector<uint32_t> rootSlot;   // by root index (position in the start-depth level): start-grid slot
    std::vector<uint32_t> order;      // layout order -> root index
    std::vector<uint8_t> owned;       // by root index: built by this rank
    std::vector<uint32_t> ownerOf;    // by root index
};
RootPlan makeRootPlan(const float4* rootCenterHalf, const uint32_t* rootCoord, uint32_t G, uint32_t startDepth, const float* boxMin,
                      float cellSize, uint32_t numThreads, uint32_t rank, uint32_t world, const uint32_t* weight = nullptr);

// One output array of a structure, as the shard exporter / assembler sees it.
struct ShardStream {
    uint8_t* dBase = nullptr;          // device array
    uint32_t elemBytes = 4;
    std::vector<uint64_t> rootBase;    // by root index: first element of the root's block
    std::vector<uint32_t> rootSize;    // by root index: elements in the root's block (all ranks' roots)
};

// Builder state kept alive between the phases of a sharded build (levels stay on the device).
struct BuildState {
    virtual ~BuildState() {}
    virtual uint32_t numStreams() const = 0;
    // phase 2: global offsets from every root's sizes (numStreams() values per start-grid slot), emit own words
    virtual void finish(sdfb200_sdf& out, const uint32_t* allSizesBySlot) = 0;
};

// host-pointer query staging (query_host.cpp); incomplete here so that the handle can own one
struct QueryStage;
struct QueryStageDeleter { void operator()(QueryStage* p) const; };

// octree node word encoding (reference: OctreeSdf::OctreeNode, include/SdfLib/OctreeSdf.h:39-98)
constexpr uint32_t kLeafBit = 1u << 31;
constexpr uint32_t kOctIndexMask = ~(3u << 30);
constexpr uint32_t kExactIndexMask = ~(1u << 31);

}  // namespace sdfb200

// ---- the opaque handle -------------------------------------------------------------------------
struct sdfb200_sdf {
    int format = SDFB200_FORMAT_OCTREE;
    int device = 0;
    float boxMin[3] = {0, 0, 0}, boxMax[3] = {0, 0, 0};
    int startGridSize = 0;
    uint32_t maxDepth = 0;
    float cellSize = 0.0f;
    // OCTREE
    float valueRange = 0.0f, minBorderValue = 0.0f;
    bool leafBlocksAligned = true;   // every leaf block offset is a multiple of 4 words (checked on load)
    uint32_t relativeDepth = 0;      // loaded files: measured depth below the start grid (validateStructure); built: maxDepth - startDepth
    // EXACT_OCTREE
    uint32_t startDepth = 0, minTrisInLeafs = 0, maxTrisInLeafs = 0, maxTrisEncoded = 0, bitEncodingStartDepth = 0,
             bitsPerIndex = 0;
    // host mirrors (what the getters / .bin writer read)
    sdfb200::HostArray<uint32_t> octree;   // OCTREE: words; EXACT: (childrenIndex, trianglesArrayIndex) pairs
    sdfb200::HostArray<uint32_t> sets;
    sdfb200::HostArray<uint8_t> masks;
    // element counts of the arrays (valid as soon as the structure is complete on the device) and whether the host mirrors
    // above hold them: replicas of a multi-GPU build fetch their mirrors on first use (ensureHostMirror, shard.cpp)
    uint64_t nOctree = 0, nSets = 0, nMasks = 0;   // uint32 words / uint32 words / bytes
    bool hostMirror = false;
    sdfb200::TriVec tris;            // EXACT: loaded from a .bin; built structures fetch it from `mesh` on demand (hostTris())
    uint32_t numTris = 0;
    std::shared_ptr<sdfb200::PreparedMesh> mesh;   // EXACT, built here: owner of the device TriangleData / frames the queries read
    const sdfb200::TriData* qTris = nullptr;       // what the exact query kernels read (mesh->dev.tris or dTris)
    const float4* qFrames = nullptr;               // (mesh->frames or dFrames)
    const sdfb200::TriVec& hostTris() { return mesh ? mesh->hostTriangleData() : tris; }
    // device copies (what the query kernels read)
    sdfb200::DevBuf<uint32_t> dOctree;
    sdfb200::DevBuf<uint32_t> dSets;
    sdfb200::DevBuf<uint8_t> dMasks;
    sdfb200::DevBuf<sdfb200::TriData> dTris;
    // private query-side copies of an EXACT_OCTREE (exact_query.cu): 80-byte triangle frames and the
    // explicit per-leaf triangle lists decoded from sets + mask chains
    sdfb200::DevBuf<float4> dFrames;
    sdfb200::DevBuf<uint64_t> dLeafLo;
    sdfb200::DevBuf<uint32_t> dLeafCnt, dLeafPool;
    // query-side index of an OCTREE (octree_query.cu, prepareOctreeQuery): one word per cell of the grid `topLevels`
    // below the start grid; -1 = the array does not meet the tile kernel's preconditions
    sdfb200::DevBuf<uint32_t> dTopIndex;
    int topLevels = -1, gridShift = 0;
    bool forcePlainQuery = false;    // SDFB200_QUERY_PLAIN=1 at build / load time: one-query-per-thread kernel (A/B measurements)
    bool packedQuery = true;         // SDFB200_QUERY_PACKED=0 at build / load time: tile kernel without the packed float32 instructions (A/B)
    // staging of host-pointer queries (query_host.cpp), created on first use under stageMutex
    std::unique_ptr<sdfb200::QueryStage, sdfb200::QueryStageDeleter> stage;
    std::mutex stageMutex;
    // sharded build: phase state, root plan, per-slot sizes of the own roots, streams for export / assembly
    bool isShard = false;          // true until sdfb200_assemble completed the structure
    std::unique_ptr<sdfb200::BuildState> build;
    sdfb200::RootPlan plan;
    std::vector<uint32_t> shardSizes;          // numStreams values per start-grid slot, 0 for roots of other ranks
    std::vector<sdfb200::ShardStream> streams;
    uint32_t slotWords = 1;                    // words of one start-slot record (OCTREE 1, EXACT 2)
    uint32_t shardScalars[2] = {0, 0};         // OCTREE: valueRange bits / ordered minBorder; EXACT: max leaf / max encoded
    sdfb200_build_stats stats = {};

};

namespace sdfb200 {

// octree_build.cu
// Builders: with world == 1 the structure is complete on return; with world > 1 only phase 1 (levels + sizes of
// the own roots) has run and out.build holds the state for finish().
void buildOctreeOnDevice(sdfb200_sdf& out, const PreparedMesh& mesh, const float* box6, uint32_t depth, uint32_t startDepth,
                         int rule, float param0, float param1, uint32_t numThreads, uint32_t rank, uint32_t world);
void finalizeOctreeScalars(sdfb200_sdf& s);   // shardScalars -> valueRange / minBorderValue
void cubifyBox(sdfb200_sdf& s, const float* box6, uint32_t startDepth);
// octree_cont.cu: InitAlgorithm::CONTINUITY (single device; the structure is complete on return)
// With world > 1 the BVH sampling of every level is sliced over the ranks and all-gathered through `allgather`.
struct SampleExchange {
    uint32_t rank = 0, world = 1;
    sdfb200_allgather_fn allgather = nullptr;
    void* user = nullptr;
};
void buildOctreeContinuityOnDevice(sdfb200_sdf& out, const PreparedMesh& mesh, const float* box6, uint32_t depth, uint32_t startDepth,
                                   int rule, float param0, float param1, const SampleExchange& exchange = SampleExchange());
// multi_device.cpp: one process, several devices
struct MultiBuildRequest {
    int format;   // SDFB200_FORMAT_OCTREE / SDFB200_FORMAT_EXACT_OCTREE
    const float* box6;
    uint32_t depth, startDepth, numThreads;
    int rule; float param0, param1; int algorithm;   // OCTREE
    uint32_t minTris;                                // EXACT_OCTREE
};
void buildMulti(const HostMesh& mesh, const MultiBuildRequest& req, const std::vector<int>& devices, std::vector<std::unique_ptr<sdfb200_sdf>>& out);
bool ncclAvailable();
// shard.cpp
void ensureHostMirror(sdfb200_sdf& s);   // downloads the structure arrays into the host mirrors once (synchronises)
uint64_t shardPayloadWords(const sdfb200_sdf& s);
void shardExport(const sdfb200_sdf& s, uint32_t* dDst, uint64_t capacityWords);
void shardAssemble(sdfb200_sdf& s, const uint32_t* dGathered, const uint64_t* wordsPerRank, uint64_t strideWords, uint32_t world);
void nearestTriangleOnDevice(const HostMesh& mesh, const float* xyz, uint64_t n, uint32_t* outTri);
void pointTriangleOnDevice(const float* tri37, const float* v123, const float* xyz, uint64_t n, int mode, float* outDist,
                           float* outGrad);
// octree_query.cu (two objects: fast = FMA Horner, exact = reference operation order)
// hostMapped: the pointers alias mapped host memory (small-batch slot): no TMA staging
void launchOctreeQueryFast(const sdfb200_sdf& s, const float* dXyz, uint64_t n, float* dDist, float* dGrad, cudaStream_t st,
                           bool hostMapped = false);
void launchOctreeTraceFast(const sdfb200_sdf& s, const float* dOrigin, const float* dDirection, uint64_t n, float epsilon, float farDistance,
                           uint32_t maxIterations, float* dHit, float* dTravelled, uint32_t* dIterations, cudaStream_t st);
void launchOctreeTraceExact(const sdfb200_sdf& s, const float* dOrigin, const float* dDirection, uint64_t n, float epsilon, float farDistance,
                            uint32_t maxIterations, float* dHit, float* dTravelled, uint32_t* dIterations, cudaStream_t st);
void prepareOctreeQuery(sdfb200_sdf& s);   // top index of a complete OCTREE structure (default stream, synchronises)
// query_host.cpp
void queryHostPointers(sdfb200_sdf& s, const float* xyz, uint64_t n, float* dist, float* grad, int flags, cudaStream_t st);
void queryDevicePointers(const sdfb200_sdf& s, const float* xyz, uint64_t n, float* dist, float* grad, int flags, cudaStream_t st);
void launchOctreeQueryExact(const sdfb200_sdf& s, const float* dXyz, uint64_t n, float* dDist, float* dGrad, cudaStream_t st);
// exact_build.cu / exact_query.cu
void buildExactOnDevice(sdfb200_sdf& out, const std::shared_ptr<PreparedMesh>& mesh, const float* box6, uint32_t maxDepth, uint32_t startDepth,
                        uint32_t minTris, uint32_t numThreads, uint32_t rank, uint32_t world);
void prepareExactQuery(sdfb200_sdf& s);
void launchExactQuery(const sdfb200_sdf& s, const float* dXyz, uint64_t n, float* dDist, float* dGrad, cudaStream_t st);
// bin_io.cpp
void saveBin(const sdfb200_sdf& s, const char* path);
void loadBin(sdfb200_sdf& s, const char* path);
void uploadStructure(sdfb200_sdf& s);
void validateStructure(sdfb200_sdf& s);

}  // namespace sdfb200
