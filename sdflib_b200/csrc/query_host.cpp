// Host side of sdfb200_query for HOST pointers (the reference's getDistance signature: caller memory in, caller
// memory out, complete on return). Device-pointer calls never come here: they only enqueue a kernel and share no
// mutable state, so they are re-entrant; host-pointer calls on one handle are serialised by the handle's mutex
// (two threads used to interleave their chunks in the same staging slots).
//
// Three regimes, chosen per call:
//   small  (n <= kSmallBatch)      one slot of MAPPED pinned memory: the points are written into it, the kernel reads
//                                  and writes it over PCIe, one stream synchronisation — no copy calls at all. This is
//                                  what the scalar getDistance(p) of the reference API costs here (launch + sync).
//   pinned (caller memory pinned)  chunks alternate between two streams: H2D of chunk k+1, kernel of chunk k and D2H
//                                  of chunk k-1 overlap (the two copy directions use different DMA engines).
//   pageable                       the same pipeline through a pinned ring owned by the handle: the caller's memory is
//                                  copied into / out of the ring by all host threads while the other slot is on the
//                                  GPU (a cudaMemcpyAsync from pageable memory is staged by the driver on one thread).
#include <algorithm>
#include <cstring>

#include "sdf_internal.h"

namespace sdfb200 {

namespace {
constexpr uint64_t kChunk = uint64_t(1) << 21;     // queries per pipeline step (24 MB of points)
constexpr uint64_t kSmallBatch = 2048;

void parallelCopy(void* dst, const void* src, size_t bytes) {
    constexpr size_t kPiece = size_t(1) << 20;
    const long pieces = long((bytes + kPiece - 1) / kPiece);
    if (pieces <= 2) { std::memcpy(dst, src, bytes); return; }
    // hostThreads(), not the OpenMP default: launchers like torchrun export OMP_NUM_THREADS=1 to every rank
#pragma omp parallel for schedule(static) num_threads(hostThreads())
    for (long i = 0; i < pieces; i++) {
        const size_t at = size_t(i) * kPiece;
        std::memcpy(static_cast<char*>(dst) + at, static_cast<const char*>(src) + at, std::min(kPiece, bytes - at));
    }
}

bool isPinnedOrManaged(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
}  // namespace

struct QueryStage {
    cudaStream_t stream[2] = {nullptr, nullptr};
    cudaEvent_t ordered = nullptr;
    DevBuf<float> dPts, dDist, dGrad;        // 2 slots of kChunk queries each (grown on demand)
    uint64_t slotQueries = 0;
    bool slotHasGrad = false;
    float* ring = nullptr;                   // pinned ring for pageable callers: 2 x (points | distances | gradients)
    uint64_t ringQueries = 0;
    float* small = nullptr;                  // mapped pinned slot: points | distances | gradients of kSmallBatch queries
    float* smallDev = nullptr;

    ~QueryStage() {
        for (cudaStream_t s : stream) if (s) cudaStreamDestroy(s);
        if (ordered) cudaEventDestroy(ordered);
        if (ring) cudaFreeHost(ring);
        if (small) cudaFreeHost(small);
    }
    void init() {
        if (stream[0]) return;
        for (cudaStream_t& s : stream) SDFB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        SDFB_CUDA(cudaEventCreateWithFlags(&ordered, cudaEventDisableTiming));
    }
    void ensureSlots(uint64_t queries, bool grad) {
        bool grown = false;
        if (slotQueries < queries) { dPts.alloc(2 * 3 * queries); dDist.alloc(2 * queries); slotQueries = queries; slotHasGrad = false; grown = true; }
        if (grad && !slotHasGrad) { dGrad.alloc(2 * 3 * slotQueries); slotHasGrad = true; grown = true; }
        if (grown) SDFB_CUDA(cudaStreamSynchronize(cudaStream_t(0)));   // device blocks are ordered on the default stream
    }
    void ensureRing(uint64_t queries) {
        if (ringQueries >= queries) return;
        if (ring) { cudaFreeHost(ring); ring = nullptr; ringQueries = 0; }
        SDFB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ring), 2 * 7 * queries * sizeof(float), cudaHostAllocDefault));
        ringQueries = queries;
    }
    void ensureSmall() {
        if (small) return;
        SDFB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&small), 7 * kSmallBatch * sizeof(float), cudaHostAllocMapped));
        SDFB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&smallDev), small, 0));
    }
};

void QueryStageDeleter::operator()(QueryStage* p) const { delete p; }

static void launchQuery(const sdfb200_sdf& s, const float* dXyz, uint64_t n, float* dDist, float* dGrad, int flags, cudaStream_t on,
                        bool hostMapped) {
    if (s.format == SDFB200_FORMAT_OCTREE) {
        if (flags & SDFB200_QUERY_EXACT_ORDER) launchOctreeQueryExact(s, dXyz, n, dDist, dGrad, on);
        else launchOctreeQueryFast(s, dXyz, n, dDist, dGrad, on, hostMapped);
    } else launchExactQuery(s, dXyz, n, dDist, dGrad, on);
}

void queryDevicePointers(const sdfb200_sdf& s, const float* xyz, uint64_t n, float* dist, float* grad, int flags, cudaStream_t st) {
    launchQuery(s, xyz, n, dist, grad, flags, st, false);
}

void queryHostPointers(sdfb200_sdf& s, const float* xyz, uint64_t n, float* dist, float* grad, int flags, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(s.stageMutex);
    if (!s.stage) s.stage.reset(new QueryStage());
    QueryStage& q = *s.stage;
    q.init();
    if (st) {   // work already queued on the caller's stream comes first (the legacy default stream needs nothing: every
                // call that changes the structure synchronises the device before it returns)
        SDFB_CUDA(cudaEventRecord(q.ordered, st));
        for (cudaStream_t cs : q.stream) SDFB_CUDA(cudaStreamWaitEvent(cs, q.ordered, 0));
    }
    if (n <= kSmallBatch) {
        q.ensureSmall();
        float* hP = q.small; float* hD = hP + 3 * kSmallBatch; float* hG = hD + kSmallBatch;
        std::memcpy(hP, xyz, 3 * n * sizeof(float));
        launchQuery(s, q.smallDev, n, q.smallDev + 3 * kSmallBatch, grad ? q.smallDev + 4 * kSmallBatch : nullptr, flags, q.stream[0], true);
        SDFB_CUDA(cudaStreamSynchronize(q.stream[0]));
        std::memcpy(dist, hD, n * sizeof(float));
        if (grad) std::memcpy(grad, hG, 3 * n * sizeof(float));
        return;
    }
    const uint64_t chunk = std::min<uint64_t>(n, kChunk);
    q.ensureSlots(chunk, grad != nullptr);
    const uint64_t slot = q.slotQueries;
    const bool direct = isPinnedOrManaged(xyz) && isPinnedOrManaged(dist) && (!grad || isPinnedOrManaged(grad));
    if (!direct) q.ensureRing(chunk);
    const uint64_t rq = q.ringQueries;
    struct InFlight { uint64_t first = 0, count = 0; } inFlight[2];
    auto drain = [&](int k) {   // results of the chunk that used slot k are in the ring: hand them to the caller
        if (!inFlight[k].count) return;
        SDFB_CUDA(cudaStreamSynchronize(q.stream[k]));
        const float* hD = q.ring + size_t(k) * 7 * rq + 3 * rq;
        parallelCopy(dist + inFlight[k].first, hD, inFlight[k].count * sizeof(float));
        if (grad) parallelCopy(grad + 3 * inFlight[k].first, hD + rq, 3 * inFlight[k].count * sizeof(float));
        inFlight[k].count = 0;
    };
    uint64_t done = 0;
    for (int k = 0; done < n; k ^= 1) {
        const uint64_t m = std::min(chunk, n - done);
        cudaStream_t cs = q.stream[k];
        float* dP = q.dPts.p + size_t(k) * 3 * slot;
        float* dD = q.dDist.p + size_t(k) * slot;
        float* dG = grad ? q.dGrad.p + size_t(k) * 3 * slot : nullptr;
        if (direct) {
            SDFB_CUDA(cudaMemcpyAsync(dP, xyz + 3 * done, 3 * m * sizeof(float), cudaMemcpyHostToDevice, cs));
            launchQuery(s, dP, m, dD, dG, flags, cs, false);
            SDFB_CUDA(cudaMemcpyAsync(dist + done, dD, m * sizeof(float), cudaMemcpyDeviceToHost, cs));
            if (grad) SDFB_CUDA(cudaMemcpyAsync(grad + 3 * done, dG, 3 * m * sizeof(float), cudaMemcpyDeviceToHost, cs));
        } else {
            drain(k);
            float* hP = q.ring + size_t(k) * 7 * rq;
            parallelCopy(hP, xyz + 3 * done, 3 * m * sizeof(float));
            SDFB_CUDA(cudaMemcpyAsync(dP, hP, 3 * m * sizeof(float), cudaMemcpyHostToDevice, cs));
            launchQuery(s, dP, m, dD, dG, flags, cs, false);
            SDFB_CUDA(cudaMemcpyAsync(hP + 3 * rq, dD, m * sizeof(float), cudaMemcpyDeviceToHost, cs));
            if (grad) SDFB_CUDA(cudaMemcpyAsync(hP + 4 * rq, dG, 3 * m * sizeof(float), cudaMemcpyDeviceToHost, cs));
            inFlight[k].first = done;
            inFlight[k].count = m;
        }
        done += m;
    }
    if (direct) { for (cudaStream_t cs : q.stream) SDFB_CUDA(cudaStreamSynchronize(cs)); }
    else { drain(0); drain(1); }
}

}  // namespace sdfb200
