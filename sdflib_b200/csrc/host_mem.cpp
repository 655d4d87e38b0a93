// Process-wide caches behind the device and host buffers of the handles:
//   * pinned host blocks (cudaHostAlloc) for the host mirrors of the structures: a D2H copy into pageable memory
//     runs at ~2-3 GB/s through the driver's staging buffer, into pinned memory at PCIe speed; allocating pinned
//     memory is slow (page locking), so freed blocks are kept and reused by the next build;
//   * the device side uses the CUDA stream-ordered allocator (cudaMallocAsync) with the pool's release threshold
//     raised, so the hundreds of per-level temporaries of a build are recycled instead of hitting cudaMalloc/cudaFree.
// Without a CUDA device (loading a .bin only to report a file error) host blocks fall back to malloc.
#include <chrono>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "sdf_internal.h"

namespace sdfb200 {

namespace {
std::mutex gMutex;
std::map<size_t, std::vector<void*>> gFreePinned;   // rounded size -> blocks
size_t gCachedBytes = 0;
constexpr size_t kMaxCachedBytes = size_t(8) << 30;

size_t roundBlock(size_t bytes) {   // 1 MiB granularity below 64 MiB, 1/8 steps above: bounded waste, high reuse
    const size_t mib = size_t(1) << 20;
    if (bytes <= 64 * mib) return (bytes + mib - 1) / mib * mib;
    size_t step = mib;
    while (step * 16 < bytes) step <<= 1;
    return (bytes + step - 1) / step * step;
}
}  // namespace

void* hostBlockAlloc(size_t bytes, size_t* outCapacity, bool* outPinned) {
    const size_t cap = roundBlock(bytes ? bytes : 1);
    *outCapacity = cap;
    {
        std::lock_guard<std::mutex> lock(gMutex);
        auto it = gFreePinned.find(cap);
        if (it != gFreePinned.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            gCachedBytes -= cap;
            *outPinned = true;
            return p;
        }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) == cudaSuccess) { *outPinned = true; return p; }
    cudaGetLastError();   // no device / out of pinned memory: plain pageable memory
    p = std::malloc(cap);
    if (!p) throw std::bad_alloc();
    *outPinned = false;
    return p;
}

void hostBlockFree(void* p, size_t capacity, bool pinned) {
    if (!p) return;
    if (!pinned) { std::free(p); return; }
    {
        std::lock_guard<std::mutex> lock(gMutex);
        if (gCachedBytes + capacity <= kMaxCachedBytes) {
            gFreePinned[capacity].push_back(p);
            gCachedBytes += capacity;
            return;
        }
    }
    cudaFreeHost(p);
}

// ---- device blocks ---------------------------------------------------------------------------------------------
// cudaMallocAsync alone recycles memory but not BLOCKS: a build allocates hundreds of temporaries of data-dependent
// sizes, and whenever the pool has no contiguous range for one of them it maps physical pages into a new range —
// tens to hundreds of milliseconds, at random (a 0.21 s CONTINUITY build took 0.9 s one time in five). Requests are
// therefore rounded to size classes (1/8 steps above 1 MiB) and freed blocks go to per-(device, class) lists, so a
// repeated build gets exactly the blocks of the previous one. Blocks above 1 GiB, or beyond 16 GiB of cached bytes,
// go back to the stream-ordered allocator.
namespace {
std::mutex gDevMutex;
std::map<std::pair<int, size_t>, std::vector<void*>> gFreeDevice;
size_t gDevCachedBytes = 0;
constexpr size_t kMaxDevCachedBytes = size_t(16) << 30, kMaxDevCachedBlock = size_t(1) << 30, kTrimOnMissBytes = size_t(2) << 30;

size_t roundDeviceBlock(size_t bytes) {
    if (bytes <= 4096) return 4096;
    size_t step = 512;
    while (step * 16 < bytes) step <<= 1;   // 1/16 .. 1/8 of the size
    return (bytes + step - 1) / step * step;
}

// SDFB200_TIMING: host time spent in the allocator since the last report (misses served by cudaMallocAsync, trims)
struct AllocClock {
    double missMs = 0, trimMs = 0;
    uint64_t misses = 0, hits = 0, trims = 0, missBytes = 0;
} gAllocClock;
inline double nowMs() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void trimDeviceCache(int device) {   // gDevMutex held
    const double t0 = nowMs();
    struct Done { double t0; ~Done() { gAllocClock.trimMs += nowMs() - t0; gAllocClock.trims++; } } done{t0};
    for (auto& kv : gFreeDevice) {
        if (kv.first.first != device) continue;
        for (void* p : kv.second) { cudaFreeAsync(p, cudaStream_t(0)); gDevCachedBytes -= kv.first.second; }
        kv.second.clear();
    }
}
}  // namespace

// Stream that new device memory is ordered on when a request misses the cache (and that overflowing blocks are
// returned on). Thread-local: a builder that runs a pass on a side stream sets it for the duration of that pass.
static thread_local cudaStream_t tBlockStream = nullptr;
void setDeviceBlockStream(cudaStream_t s) { tBlockStream = s; }

static bool blockCacheOff() {
    // One-shot builds of very large structures are better off without the cache (profiles/r1_summary.md, config 4:
    // first call 5.7 s instead of 9.6 s); repeat builds want it.
    static const bool off = std::getenv("SDFB200_NO_BLOCK_CACHE") != nullptr;
    return off;
}

void* deviceBlockAlloc(size_t bytes, size_t* outCapacity) {
    if (blockCacheOff()) {
        *outCapacity = bytes;
        void* q = nullptr;
        cudaError_t err = cudaMallocAsync(&q, bytes, tBlockStream);
        if (err != cudaSuccess) throw Error(SDFB200_ERR_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(err));
        return q;
    }
    const size_t cap = roundDeviceBlock(bytes);
    *outCapacity = cap;
    int device = 0;
    cudaGetDevice(&device);
    {
        std::lock_guard<std::mutex> lock(gDevMutex);
        auto it = gFreeDevice.find(std::make_pair(device, cap));
        if (it != gFreeDevice.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            gDevCachedBytes -= cap;
            gAllocClock.hits++;
            return p;
        }
    }
    {   // A miss while a lot is cached means the sizes have moved on (a first build growing level by level): hand the
        // cached blocks back to the stream-ordered pool so that it can serve this request from their memory instead of
        // mapping fresh pages (first build of the 5.2 M-triangle ExactOctreeSdf: 8.9 s without this, 101 GB touched).
        std::lock_guard<std::mutex> lock(gDevMutex);
        if (gDevCachedBytes > kTrimOnMissBytes) trimDeviceCache(device);
    }
    void* p = nullptr;
    const double tMiss = nowMs();
    struct Miss { double t0; size_t cap; ~Miss() { std::lock_guard<std::mutex> lock(gDevMutex); gAllocClock.missMs += nowMs() - t0; gAllocClock.misses++; gAllocClock.missBytes += cap; } } miss{tMiss, cap};
    cudaError_t e = cudaMallocAsync(&p, cap, tBlockStream);
    if (e == cudaErrorMemoryAllocation) {   // give the cached blocks back and try once more
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> lock(gDevMutex);
            trimDeviceCache(device);
        }
        cudaStreamSynchronize(cudaStream_t(0));
        e = cudaMallocAsync(&p, cap, tBlockStream);
    }
    if (e != cudaSuccess) throw Error(SDFB200_ERR_CUDA, std::string("cudaMallocAsync(") + std::to_string(cap) + " bytes): " + cudaGetErrorString(e));
    return p;
}

// End of a top-level build: a cache that holds more than kTrimOnMissBytes goes back to the stream-ordered pool, so that the
// next build starts from the same state as this one did (an empty cache and a pool whose free ranges have coalesced)
// instead of from whatever the last levels of this build happened to leave: a series of identical large builds then
// repeats itself (before: 0.75 / 0.98 / 1.39 / 1.41 s for the same Dragon-class build, depending on its predecessor).
// Small builds keep their blocks (a C2 / C3 build caches well under the threshold) and stay exact-reuse.
void settleDeviceCache(int device) {
    std::lock_guard<std::mutex> lock(gDevMutex);
    static const bool timing = std::getenv("SDFB200_TIMING") != nullptr;
    struct Report { ~Report() {
        if (!timing) return;
        std::fprintf(stderr, "[sdfb200] allocator: %llu hits, %llu misses (%.1f GB, %.2f ms in cudaMallocAsync), %llu trims (%.2f ms), %.1f GB cached\n",
                     (unsigned long long)gAllocClock.hits, (unsigned long long)gAllocClock.misses, double(gAllocClock.missBytes) / 1e9, gAllocClock.missMs,
                     (unsigned long long)gAllocClock.trims, gAllocClock.trimMs, double(gDevCachedBytes) / 1e9);
        gAllocClock = AllocClock();
    } } report;
    if (gDevCachedBytes <= kTrimOnMissBytes) return;
    int current = 0;
    cudaGetDevice(&current);
    if (current != device) cudaSetDevice(device);
    trimDeviceCache(device);
    if (current != device) cudaSetDevice(current);
    cudaGetLastError();
}

void deviceBlockFree(void* p, size_t capacity) {
    if (!p) return;
    if (blockCacheOff()) { cudaFreeAsync(p, tBlockStream); return; }
    cudaPointerAttributes attr;
    int device = 0;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess) device = attr.device; else cudaGetLastError();
    if (capacity <= kMaxDevCachedBlock) {
        std::lock_guard<std::mutex> lock(gDevMutex);
        if (gDevCachedBytes + capacity <= kMaxDevCachedBytes) {
            gFreeDevice[std::make_pair(device, capacity)].push_back(p);
            gDevCachedBytes += capacity;
            return;
        }
    }
    cudaFreeAsync(p, tBlockStream);
}

void releaseCachedMemory() {
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess) { cudaGetLastError(); nDev = 0; }
    int current = 0;
    if (nDev) cudaGetDevice(&current);
    {
        std::lock_guard<std::mutex> lock(gDevMutex);
        for (auto& kv : gFreeDevice) {
            if (kv.second.empty()) continue;
            cudaSetDevice(kv.first.first);
            cudaDeviceSynchronize();
            for (void* p : kv.second) { cudaFreeAsync(p, cudaStream_t(0)); gDevCachedBytes -= kv.first.second; }
            kv.second.clear();
        }
    }
    for (int d = 0; d < nDev; d++) {
        cudaMemPool_t pool;
        cudaSetDevice(d);
        if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) { cudaDeviceSynchronize(); cudaMemPoolTrimTo(pool, 0); }
    }
    if (nDev) cudaSetDevice(current);
    cudaGetLastError();
    std::lock_guard<std::mutex> lock(gMutex);
    for (auto& kv : gFreePinned) {
        for (void* p : kv.second) { cudaFreeHost(p); gCachedBytes -= kv.first; }
        kv.second.clear();
    }
}

void configureDevicePool(int device) {
    static std::mutex m;
    static bool done[64] = {};
    std::lock_guard<std::mutex> lock(m);
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t threshold = UINT64_MAX;   // keep freed memory in the pool (returned to the driver at process exit / trim)
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    cudaGetLastError();
    done[device] = true;
}

}  // namespace sdfb200
