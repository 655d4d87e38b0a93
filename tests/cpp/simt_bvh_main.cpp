// Runs the device-side BVH build of sdflib_b200/csrc/bvh_build.cuh ON THE CPU, from the very same source and through the very
// same launch sequence (bvhBuildLevels), under a CTA emulation: every thread of a CTA is a host thread, __syncthreads and
// the warp collectives are barriers, __shared__ variables are statics (CTAs run one after the other). The build restates
// libstdc++'s std::sort as data-parallel steps — a mistake would change which of two equidistant triangles the OctreeSdf
// builders pick — so before it ever runs on a GPU it must reproduce the host builder of mesh_host.cpp (which calls the
// library's own sort routines) node for node, bit for bit. Small CTAs and a small shared-memory limit (compile-time
// overrides) push meshes of a few thousand triangles through every path: global partition rounds, shared-memory sorts,
// tiny nodes, warp and thread centre sums.
//
//   simt_bvh_main mesh <isosphere subdivisions> <0|1: displaced> [keep] -> "ok <nodes> nodes identical" (keep: only the first <keep> triangles)
//   simt_bvh_main sort <n> <distinct keys> <depth limit> <seed>      -> "ok sort": partition rounds + shared-memory sort (+ heap
//                                                                       sort once the depth limit is spent) against the library
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <random>
#include <thread>
#include <vector>

#include "sdf_internal.h"

namespace simt {
struct Dim { unsigned x = 0, y = 0, z = 0; };
struct Warp {
    std::barrier<> bar;
    uint64_t slot[32];
    explicit Warp(int lanes) : bar(lanes) {}
};
struct Cta {
    std::barrier<> bar;
    explicit Cta(int threads) : bar(threads) {}
};
thread_local Dim tThread, tBlock, tBlockDim, tGridDim;
thread_local Warp* tWarp = nullptr;
thread_local Cta* tCta = nullptr;
thread_local unsigned tLane = 0;

template <class T> T exchange(T v, unsigned src) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    tWarp->slot[tLane] = raw;
    tWarp->bar.arrive_and_wait();
    raw = tWarp->slot[src & 31u];
    tWarp->bar.arrive_and_wait();
    T out;
    std::memcpy(&out, &raw, sizeof(T));
    return out;
}
inline unsigned ballot(bool pred) {
    tWarp->slot[tLane] = pred ? 1u : 0u;
    tWarp->bar.arrive_and_wait();
    unsigned r = 0;
    for (unsigned l = 0; l < 32; l++) r |= unsigned(tWarp->slot[l] & 1u) << l;
    tWarp->bar.arrive_and_wait();
    return r;
}
template <class K> void launch(unsigned grid, unsigned block, K kernel) {
    if (block % 32) { std::fprintf(stderr, "block size must be a multiple of 32\n"); std::exit(2); }
    for (unsigned b = 0; b < grid; b++) {
        Cta cta{int(block)};
        std::vector<std::unique_ptr<Warp>> warps;
        for (unsigned w = 0; w < block / 32; w++) warps.emplace_back(new Warp(32));
        for (auto& w : warps) std::memset(w->slot, 0, sizeof(w->slot));
        std::vector<std::thread> threads;
        for (unsigned t = 0; t < block; t++)
            threads.emplace_back([&, b, t] {
                tCta = &cta; tWarp = warps[t / 32].get(); tLane = t % 32;
                tThread.x = t; tBlock.x = b; tBlockDim.x = block; tGridDim.x = grid;
                kernel();
                tWarp->bar.arrive_and_drop();
                cta.bar.arrive_and_drop();
            });
        for (std::thread& t : threads) t.join();
    }
}
}  // namespace simt

#define __launch_bounds__(...)
#undef __shared__
#define __shared__ static
#define gridDim simt::tGridDim
#define threadIdx simt::tThread
#define blockIdx simt::tBlock
#define blockDim simt::tBlockDim
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return simt::exchange(v, unsigned(src)); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return simt::exchange(v, simt::tLane ^ unsigned(m)); }
inline unsigned __ballot_sync(unsigned, bool p) { return simt::ballot(p); }
inline void __syncthreads() { simt::tCta->bar.arrive_and_wait(); }
inline void __syncwarp() { simt::tWarp->bar.arrive_and_wait(); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v ? __builtin_clz(unsigned(v)) : 32; }
inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline uint32_t atomicOr(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int32_t atomicMin(int32_t* p, int32_t v) {
    int32_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline int32_t atomicMax(int32_t* p, int32_t v) {
    int32_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}

#define BVH_LAUNCH(kernel, grid, block, stream, ...) simt::launch((grid), (block), [&] { kernel(__VA_ARGS__); })

namespace sdfb200 {
namespace {
#include "bvh_build.cuh"

struct HostRt {   // every "stream" is the calling thread
    const float4* triVerts = nullptr;   // set: the centre sums of the top levels take the host path (bvhHostCentres + bvhPutCentresKernel)
    int32_t hostChainMin = kBvhHostChain;
    bool hostCentres(int level, int, const int32_t* order, int32_t n, BvhNode* nodes) {
        if (!triVerts || (level & 1) == 0) return false;   // odd levels on the host path, even ones through bvhCentreKernel: both are checked
        std::vector<double> centres(size_t(3) << level);
        bvhHostCentres(n, level, order, 0u, 1u << level, [&](int32_t t, int k) { return &triVerts[size_t(t) * 3 + k].x; }, centres.data());
        simt::launch(((1u << level) + 63) / 64, 64, [&] { bvhPutCentresKernel(n, level, centres.data(), nodes); });
        return true;
    }
    int mainStream() { return 0; }
    int sideStream(int) { return 0; }
    void sideWaitsForMain(int) {}
    void mainWaitsForSides() {}
    void fill(void* p, int byte, size_t bytes, int) { std::memset(p, byte, bytes); }
    void copy(void* d, const void* s, size_t bytes, int) { std::memcpy(d, s, bytes); }
};
}  // namespace
}  // namespace sdfb200

using namespace sdfb200;

static int runMesh(uint32_t subdivisions, bool displaced, uint32_t keep) {
    uint32_t nv = 0, ni = 0;
    sdfb200_make_isosphere(subdivisions, nullptr, nullptr, &nv, &ni);
    std::vector<float> verts(size_t(nv) * 3);
    std::vector<uint32_t> idx(ni);
    sdfb200_make_isosphere(subdivisions, verts.data(), idx.data(), &nv, &ni);
    if (keep && keep * 3 < ni) { ni = keep * 3; idx.resize(ni); }
    if (displaced)
        for (uint32_t v = 0; v < nv; v++) {
            float* p = &verts[size_t(v) * 3];
            const float d = 0.2f * std::sin(4.f * p[0]) * std::sin(3.f * p[1] + 2.1f) * std::sin(5.f * p[2] + 0.7f);
            for (int a = 0; a < 3; a++) p[a] = p[a] * (1.f + d) + 0.01f * float(a + 1);
        }
    HostMesh host{reinterpret_cast<const f3*>(verts.data()), nv, idx.data(), ni};
    const RawVec<BvhNode> want = buildBvh(host);
    const int32_t n = int32_t(ni / 3);
    std::vector<float4> triVerts(ni);
    for (uint32_t t = 0; t < ni; t++) triVerts[t] = make_float4(verts[size_t(idx[t]) * 3], verts[size_t(idx[t]) * 3 + 1], verts[size_t(idx[t]) * 3 + 2], 0.f);

    const BvhShape shape = bvhShape(n);
    std::vector<float> keys(n);
    std::vector<int32_t> ids(n), orders(size_t(std::max(1, shape.sortLevels)) * n), boxMin(size_t(3) << shape.sortLevels), boxMax(size_t(3) << shape.sortLevels);
    std::vector<uint32_t> lpos(n), rpos(n), counters(kBvhCounters);
    std::vector<BvhSortTask> bigA(n / kBvhSmallMax + 2), bigB(n / kBvhSmallMax + 2), small(n / 2 + 2);
    std::vector<BvhNode> nodes(size_t(2) * n - 1);
    std::memset(nodes.data(), 0xEE, nodes.size() * sizeof(BvhNode));
    BvhBuffers B{keys.data(), ids.data(), lpos.data(), rpos.data(), orders.data(), boxMin.data(), boxMax.data(), {bigA.data(), bigB.data()}, small.data(),
                 counters.data(), nodes.data()};
    HostRt rt;
    rt.triVerts = triVerts.data();   // (kBvhHostChain is lowered by the build flags of the test, so small meshes reach the host path)
    bvhBuildLevels(rt, n, triVerts.data(), B, 3u);

    long bad = 0;
    for (size_t i = 0; i < nodes.size(); i++) {
        const BvhNode& a = nodes[i];
        const BvhNode& b = want[i];
        bool same = a.left == b.left && a.right == b.right && (a.pad[0] != 0) == (b.pad[0] != 0);
        if (same && !b.pad[0])   // inner node: both child spheres, bit for bit (a leaf node's sphere fields are never written by either builder)
            same = std::memcmp(a.lc, b.lc, 32) == 0 && std::memcmp(a.rc, b.rc, 32) == 0;
        if (!same && bad++ < 5)
            std::fprintf(stderr, "node %zu: links %d %d / %d %d, leaf %d / %d, lr %.17g / %.17g, lc0 %.17g / %.17g\n", i, a.left, a.right, b.left, b.right,
                         a.pad[0], b.pad[0], a.lr, b.lr, a.lc[0], b.lc[0]);
    }
    if (bad) { std::fprintf(stderr, "%ld of %zu nodes differ\n", bad, nodes.size()); return 1; }
    std::printf("ok %zu nodes identical\n", nodes.size());
    return 0;
}

// partition rounds + shared-memory sort from an arbitrary depth limit, against libstdc++'s own introsort loop and final pass
static int runSort(int32_t n, uint32_t distinct, int depth, uint32_t seed) {
    struct Rec { float key; int32_t id; };
    std::mt19937 rng(seed);
    std::vector<Rec> want(n);
    std::vector<float> keys(n);
    std::vector<int32_t> ids(n);
    for (int32_t i = 0; i < n; i++) { keys[i] = float(rng() % distinct) * 0.37f - 3.0f; ids[i] = i; want[i] = Rec{keys[i], i}; }
#if defined(__GLIBCXX__)
    auto comp = __gnu_cxx::__ops::__iter_comp_iter([](const Rec& a, const Rec& b) { return a.key < b.key; });
    std::__introsort_loop(want.begin(), want.end(), long(depth), comp);
    std::__final_insertion_sort(want.begin(), want.end(), comp);
#else
    std::printf("ok sort (skipped: not libstdc++)\n");
    return 0;
#endif
    std::vector<uint32_t> lpos(n), rpos(n), counters(kBvhCounters, 0u);
    std::vector<BvhSortTask> big[2] = {std::vector<BvhSortTask>(n / kBvhSmallMax + 2), std::vector<BvhSortTask>(n / kBvhSmallMax + 2)}, small(n / 2 + 2);
    if (n > kBvhSmallMax) { big[0][0] = BvhSortTask{0, n, depth}; counters[1] = 1; }
    else if (n > 1) { small[0] = BvhSortTask{0, n, depth}; counters[0] = 1; }
    for (int r = 0; r <= depth; r++)
        simt::launch(3, kBvhBigThreads, [&] {
            bvhBigPartitionKernel(keys.data(), ids.data(), lpos.data(), rpos.data(), big[r & 1].data(), &counters[1 + r], big[(r + 1) & 1].data(), &counters[2 + r],
                                  small.data(), &counters[0]);
        });
    simt::launch(5, kBvhSmallThreads, [&] { bvhSmallSortKernel(keys.data(), ids.data(), small.data(), &counters[0]); });
    for (int32_t i = 0; i < n; i++)
        if (ids[i] != want[i].id) { std::fprintf(stderr, "position %d: id %d, library %d (key %g / %g)\n", i, ids[i], want[i].id, keys[i], want[i].key); return 1; }
    std::printf("ok sort\n");
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 4 && std::strcmp(argv[1], "mesh") == 0) return runMesh(uint32_t(std::atoi(argv[2])), std::atoi(argv[3]) != 0, argc > 4 ? uint32_t(std::atoi(argv[4])) : 0u);
    if (argc >= 6 && std::strcmp(argv[1], "sort") == 0) return runSort(std::atoi(argv[2]), uint32_t(std::atoi(argv[3])), std::atoi(argv[4]), uint32_t(std::atoi(argv[5])));
    std::fprintf(stderr, "usage: simt_bvh_main mesh <subdivisions> <displaced> | sort <n> <distinct> <depth> <seed>\n");
    return 2;
}
