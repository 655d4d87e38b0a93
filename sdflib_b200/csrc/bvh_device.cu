// The float64 sphere BVH of the OctreeSdf builders, built ON THE DEVICE (round 1 and most of round 2 built
// it on host threads: 37 ms of the 122 ms C2 build, 2.7 s of the Lucy-class mesh preparation).
// Reference: tmd::TriangleMeshDistance::_build_tree, libs/InteractiveComputerGraphics/.../TriangleMeshDistance.h:421-490.
//
// The tree is the specification — its traversal order decides between equidistant triangles — and two of its ingredients
// look inherently serial: std::sort's unstable permutation of tied keys, and a sequential float64 centre sum. What makes a
// device build possible (kernels and their restated library routines: bvh_build.cuh):
//   * SHAPE IS STATIC. mid = (begin + end) / 2 whatever the data: every node's range and pre-order id follow from n alone
//     (bvhSegOfPosition / bvhSegOfSlot, a few integer steps), so the build is level-synchronous with no allocation, no queues
//     of nodes and no host round trip: one thread per triangle finds its node, one thread per slot finds its range.
//   * std::sort IS A DETERMINISTIC SEQUENCE OF HOARE PARTITIONS. A partition's swaps pair the k-th element from the left that
//     is not below the pivot with the k-th from the right that is not above it — two ordered compactions (ballot + popc +
//     per-warp counts), a monotone predicate for the number of swaps, and disjoint swaps; the recursion is a list of
//     (range, remaining depth) tasks processed in rounds. Ranges of <= 2048 elements finish inside one CTA's shared memory
//     (explicit stack, the same partition step, heap sort by one thread when libstdc++'s depth limit runs out); the final
//     insertion pass is a stable sort of pieces of <= 16 elements, done as ranks. Larger ranges take one global-memory
//     partition step per launch (CTA of 1024 per range); the number of launches per level is the depth limit, most are empty.
//   * ONLY THE ORDER IS ON THE CRITICAL PATH. The split axis needs the bounding box (atomicMin / atomicMax of order-preserving
//     integer images of the coordinates), the keys need the axis; the centre sums (a dependent chain of 3m float64 additions
//     per axis: a warp streams the vertices through shared memory, three lanes add) and the radii (atomicMax of float64 bit
//     patterns) of a level run on side streams from a snapshot of the order while the main stream sorts the next levels. The
//     chains of the top levels (half, a quarter, an eighth of the mesh each) are the critical path of the build — ~30 cycles per
//     dependent DADD — and a host core adds 14 x faster: when the caller's arrays are at hand those levels' sums run on host
//     threads (cudaLaunchHostFunc on the level's side stream: order snapshot down, three doubles per node back).
// Checked without a GPU: tests/cpp/simt_bvh_main.cpp runs this launch sequence from source under a CTA emulation against the
// host builder (which calls libstdc++'s own routines) node for node, and the sort alone, with forced depth limits, against
// std::__introsort_loop / __final_insertion_sort. SDFB200_HOST_BVH=1 selects the host builder (A/B, and the parity test).
#include <cuda_runtime.h>

#include <thread>

#include "device_utils.cuh"
#include "host_sort.h"

#define BVH_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)

namespace sdfb200 {
namespace {
#include "bvh_build.cuh"

constexpr int kSideStreams = 8;

// centre sums of one level on host threads, queued into a side stream with cudaLaunchHostFunc (no CUDA call inside)
struct HostCentreJob {
    const HostMesh* mesh;
    int32_t n;
    int level;
    const int32_t* order;   // pinned copy of the level's order snapshot
    double* centres;        // pinned, 3 per slot
};
void CUDART_CB hostCentreCallback(void* p) {
    const HostCentreJob& j = *static_cast<const HostCentreJob*>(p);
    const HostMesh& m = *j.mesh;
    auto vertex = [&m](int32_t t, int k) { return &m.verts[m.idx[size_t(t) * 3 + size_t(k)]].x; };
    const uint32_t slots = 1u << j.level;
    std::vector<std::thread> team;
    for (uint32_t s = 1; s < slots; s++)
        if (!tryFork(team, [&j, &vertex, s] { bvhHostCentres(j.n, j.level, j.order, s, s + 1, vertex, j.centres); }))
            bvhHostCentres(j.n, j.level, j.order, s, s + 1, vertex, j.centres);
    bvhHostCentres(j.n, j.level, j.order, 0u, 1u, vertex, j.centres);
    for (std::thread& t : team) t.join();
}

struct CudaRt {
    cudaStream_t side[kSideStreams];
    std::vector<cudaEvent_t> events;
    bool used[kSideStreams] = {};
    // host-side centre sums (levels 1 .. hostLevels): staging prepared by buildBvhOnDevice before the first launch
    const HostMesh* hostMesh = nullptr;
    int hostLevels = 0;
    std::vector<HostCentreJob> jobs;            // [level]
    double* dCentres = nullptr;                 // device, 3 doubles per slot, level l at offset 3 * 2^l
    bool hostCentres(int level, int sideIdx, const int32_t* orderDev, int32_t n, BvhNode* nodes) {
        if (!hostMesh || level < 1 || level > hostLevels) return false;
        const cudaStream_t ss = sideStream(sideIdx);
        HostCentreJob& j = jobs[size_t(level)];
        const size_t slots = size_t(1) << level;
        SDFB_CUDA(cudaMemcpyAsync(const_cast<int32_t*>(j.order), orderDev, size_t(n) * 4, cudaMemcpyDeviceToHost, ss));
        SDFB_CUDA(cudaLaunchHostFunc(ss, hostCentreCallback, &j));
        SDFB_CUDA(cudaMemcpyAsync(dCentres + 3 * slots, j.centres, slots * 3 * sizeof(double), cudaMemcpyHostToDevice, ss));
        bvhPutCentresKernel<<<divUp(slots, 64), 64, 0, ss>>>(n, level, dCentres + 3 * slots, nodes);
        return true;
    }
    CudaRt() {
        for (cudaStream_t& s : side) SDFB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    }
    ~CudaRt() {
        for (cudaStream_t s : side) cudaStreamDestroy(s);
        for (cudaEvent_t e : events) cudaEventDestroy(e);
    }
    cudaEvent_t newEvent() {
        cudaEvent_t e;
        SDFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        events.push_back(e);
        return e;
    }
    cudaStream_t mainStream() { return cudaStream_t(0); }   // the legacy stream, like every other kernel of the library (DevBuf blocks are ordered on it)
    cudaStream_t sideStream(int i) { return side[i % kSideStreams]; }
    void sideWaitsForMain(int i) {
        const cudaEvent_t e = newEvent();
        SDFB_CUDA(cudaEventRecord(e, mainStream()));
        SDFB_CUDA(cudaStreamWaitEvent(sideStream(i), e, 0));
        used[i % kSideStreams] = true;
    }
    void mainWaitsForSides() {
        for (int i = 0; i < kSideStreams; i++) {
            if (!used[i]) continue;
            const cudaEvent_t e = newEvent();
            SDFB_CUDA(cudaEventRecord(e, side[i]));
            SDFB_CUDA(cudaStreamWaitEvent(mainStream(), e, 0));
        }
    }
    void fill(void* p, int byte, size_t bytes, cudaStream_t s) { SDFB_CUDA(cudaMemsetAsync(p, byte, bytes, s)); }
    void copy(void* d, const void* s, size_t bytes, cudaStream_t st) { SDFB_CUDA(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, st)); }
};
}  // namespace

// m.triVerts (the pre-gathered vertex records) must be in place; fills m.bvh. Synchronous: returns with the tree complete.
// host (optional): the caller's mesh arrays; with them the centre sums of the top levels run on host threads (bvh_build.cuh).
void buildBvhOnDevice(MeshOnDevice& m, const HostMesh* host) {
    if (m.numTriangles < 1 || m.numTriangles > (1u << 30)) throw Error(SDFB200_ERR_INVALID, "the BVH needs between 1 and 2^30 triangles");
    const int32_t n = int32_t(m.numTriangles);
    const BvhShape shape = bvhShape(n);
    int device = 0;
    SDFB_CUDA(cudaGetDevice(&device));
    m.bvh.alloc(size_t(2) * size_t(n) - 1);
    DevBuf<float> keys(static_cast<size_t>(n));
    DevBuf<int32_t> ids(static_cast<size_t>(n)), orders(size_t(std::max(1, shape.sortLevels)) * size_t(n));
    DevBuf<int32_t> boxMin(size_t(3) << std::max(0, shape.sortLevels - 1)), boxMax(size_t(3) << std::max(0, shape.sortLevels - 1));
    const bool big = n > kBvhSmallMax;
    DevBuf<uint32_t> lpos(big ? size_t(n) : 1), rpos(big ? size_t(n) : 1), counters(kBvhCounters);
    DevBuf<BvhSortTask> bigA(size_t(n) / kBvhSmallMax + 2), bigB(size_t(n) / kBvhSmallMax + 2), small(size_t(n) / 2 + 2);
    const BvhBuffers B{keys.p, ids.p, lpos.p, rpos.p, orders.p, boxMin.p, boxMax.p, {bigA.p, bigB.p}, small.p, counters.p, m.bvh.p};
    // host-side centre sums: levels 1 .. hostLevels hold nodes of at least kBvhHostChain triangles
    static const bool hostChains = [] { const char* e = std::getenv("SDFB200_BVH_HOST_CHAINS"); return !(e && e[0] == '0'); }();
    int hostLevels = 0;
    if (host && hostChains)
        while (hostLevels + 1 < shape.sortLevels && shape.minSize[hostLevels + 1] >= kBvhHostChain && hostLevels < 4) hostLevels++;
    DevBuf<double> dCentres(size_t(6) << hostLevels);
    struct Pinned { void* p = nullptr; size_t cap = 0; bool pinned = false; ~Pinned() { hostBlockFree(p, cap, pinned); } };
    std::vector<Pinned> staging(size_t(2) * (hostLevels + 1));
    {
        CudaRt rt;
        rt.hostMesh = hostLevels ? host : nullptr;
        rt.hostLevels = hostLevels;
        rt.dCentres = dCentres.p;
        rt.jobs.resize(size_t(hostLevels) + 1);
        for (int l = 1; l <= hostLevels; l++) {
            Pinned& o = staging[size_t(2) * l];
            Pinned& c = staging[size_t(2) * l + 1];
            o.p = hostBlockAlloc(size_t(n) * 4, &o.cap, &o.pinned);
            c.p = hostBlockAlloc((size_t(3) << l) * sizeof(double), &c.cap, &c.pinned);
            rt.jobs[size_t(l)] = HostCentreJob{host, n, l, static_cast<const int32_t*>(o.p), static_cast<double*>(c.p)};
        }
        int sms = 0;
        SDFB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
        bvhBuildLevels(rt, n, m.triVerts.p, B, uint32_t(sms > 0 ? sms : 1));
        SDFB_CUDA(cudaGetLastError());
        SDFB_CUDA(cudaDeviceSynchronize());   // side streams included: the temporaries above go back to the block cache after this
    }
    m.rootLink = 0;
    if (n == 1) {   // single-triangle mesh: the root is a leaf, the traversal starts at ~triangleId
        m.rootLink = ~0;
    }
    // height of the median-split tree (halves of floor / ceil size) = deepest possible stack, + 1 spare
    uint32_t t = m.numTriangles, h = 0;
    while (t > 1) { t = t - t / 2; h++; }
    m.stackDepth = int(h) + 1;
}

}  // namespace sdfb200
