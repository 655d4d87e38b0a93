// Process-wide caches behind the device and host buffers of the handles:
//   * pinned host blocks (cudaHostAlloc) for the host mirrors of the structures: a D2H copy into pageable memory
//     runs at ~2-3 GB/s through the driver's staging buffer, into pinned memory at PCIe speed; allocating pinned
//     memory is slow (page locking), so freed blocks are kept and reused by the next build;
//   * the device side uses the CUDA stream-ordered allocator (cudaMallocAsync) with the pool's release threshold
//     raised, so the hundreds of per-level temporaries of a build are recycled instead of hitting cudaMalloc/cudaFree.
// Without a CUDA device (loading a .bin only to report a file error) host blocks fall back to malloc.
#include <cstdlib>
#include <map>
#include <mutex>
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
