// Runs the BVH sampling kernels of sdflib_b200/csrc/bvh_sampler.cuh ON THE CPU, from the same source, under the lock-step
// warp emulation of simt_query_main.cpp (lanes = host threads, warp collectives = barrier exchanges, shared memory = a
// plain array). The experimental lane-refill schedule (sampleOwnersRefillKernel + finishOwnersKernel) was written without
// a GPU at hand: a mistake in its refill / termination logic would hang the GPU box, so it is checked here first — it must
// terminate and give, sample for sample, the bits of the plain one-sample-per-thread kernel (sampleOwnersKernel, the
// measured default path). Mesh set-up (TriangleData, BVH) comes from the product library's host functions.
//
//   simt_sampler_main <isosphere subdivisions> <nodes> [leaf batch] [scheduled: 0|1]      prints "ok <samples> identical"
#include <algorithm>
#include <barrier>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "sdf_internal.h"

namespace simt {
struct Dim { unsigned x = 0, y = 0, z = 0; };
struct Warp {
    std::barrier<> bar{32};
    uint64_t slot[32];
};
thread_local Dim tThread, tBlock, tBlockDim;
thread_local Warp* tWarp = nullptr;
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
    for (unsigned l = 0; l < 32; l++) r |= unsigned(tWarp->slot[l]) << l;
    tWarp->bar.arrive_and_wait();
    return r;
}
template <class K> void launch(unsigned grid, unsigned block, K kernel) {
    for (unsigned b = 0; b < grid; b++)
        for (unsigned w = 0; w < block / 32; w++) {
            Warp warp;
            std::vector<std::thread> lanes;
            for (unsigned l = 0; l < 32; l++)
                lanes.emplace_back([&, b, w, l] {
                    tWarp = &warp; tLane = l;
                    tThread.x = w * 32 + l; tBlock.x = b; tBlockDim.x = block;
                    kernel();
                    warp.bar.arrive_and_drop();
                });
            for (std::thread& t : lanes) t.join();
        }
}
}  // namespace simt

#define __launch_bounds__(...)
#define threadIdx simt::tThread
#define blockIdx simt::tBlock
#define blockDim simt::tBlockDim
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return simt::exchange(v, unsigned(src)); }
inline unsigned __ballot_sync(unsigned, bool p) { return simt::ballot(p); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

namespace sdfb200 {
namespace {
alignas(16) double bvhStackSmem[64 * 128 * 2];   // the CTA's dynamic shared memory (stack columns of 128 threads)
#include "bvh_sampler.cuh"
}
}  // namespace sdfb200

using namespace sdfb200;

int main(int argc, char** argv) {
    const uint32_t subdivisions = argc > 1 ? uint32_t(std::atoi(argv[1])) : 3;
    const uint32_t nNodes = argc > 2 ? uint32_t(std::atoi(argv[2])) : 60;
    const int leafBatch = argc > 3 ? std::atoi(argv[3]) : 16;   // lanes that must hold a leaf before the leaf branch runs
    uint32_t nv = 0, ni = 0;
    sdfb200_make_isosphere(subdivisions, nullptr, nullptr, &nv, &ni);
    std::vector<float> verts(size_t(nv) * 3);
    std::vector<uint32_t> idx(ni);
    sdfb200_make_isosphere(subdivisions, verts.data(), idx.data(), &nv, &ni);
    for (uint32_t v = 0; v < nv; v++) {   // bumpy and off-centre: traversal lengths vary a lot between samples
        float* p = &verts[size_t(v) * 3];
        const float d = 0.2f * std::sin(4.f * p[0]) * std::sin(3.f * p[1] + 2.1f) * std::sin(5.f * p[2] + 0.7f);
        for (int a = 0; a < 3; a++) p[a] = p[a] * (1.f + d) + 0.01f * float(a + 1);
    }
    HostMesh host{reinterpret_cast<const f3*>(verts.data()), nv, idx.data(), ni};
    TriVec tris = computeTriangleData(host);
    RawVec<BvhNode> bvh = buildBvh(host);
    std::vector<float4> triVerts(ni);
    for (uint32_t t = 0; t < ni; t++) triVerts[t] = make_float4(verts[size_t(idx[t]) * 3], verts[size_t(idx[t]) * 3 + 1], verts[size_t(idx[t]) * 3 + 2], 0.f);
    DeviceMesh mesh;
    mesh.verts = host.verts; mesh.idx = host.idx; mesh.tris = tris.data(); mesh.bvh = bvh.data();
    mesh.numTriangles = ni / 3; mesh.triVerts = triVerts.data();
    mesh.rootLink = bvh[0].pad[0] ? ~bvh[0].right : 0;
    uint32_t n = mesh.numTriangles, h = 0;
    while (n > 1) { n = n - n / 2; h++; }
    mesh.stackDepth = int(h) + 1;
    if (size_t(mesh.stackDepth) * 128 * 12 > sizeof(bvhStackSmem)) return 2;

    // nodes scattered through and around the mesh, half sizes over two octaves; every sample of every node is an "owner"
    std::vector<float4> centerHalf(nNodes);
    uint32_t rng = 99991u;
    auto next = [&] { rng = rng * 1664525u + 1013904223u; return float(rng >> 8) / float(1 << 24); };
    for (float4& c : centerHalf) c = make_float4(3.0f * next() - 1.5f, 3.0f * next() - 1.5f, 3.0f * next() - 1.5f, 0.05f + 0.2f * next());
    const uint32_t count = nNodes * 19 - 5, first = 3;   // a slice that starts and ends inside a node, like a rank's share
    std::vector<uint32_t> owners(size_t(nNodes) * 19);
    for (uint32_t t = 0; t < owners.size(); t++) owners[t] = uint32_t((uint64_t(t) * 7919u) % owners.size());   // a permutation (7919 is prime)

    std::vector<float4> plain(count), refill(count);
    std::memset(plain.data(), 0xAB, count * sizeof(float4));
    std::memset(refill.data(), 0xCD, count * sizeof(float4));
    simt::launch((count + 127) / 128, 128, [&] { sampleOwnersKernel(mesh, centerHalf.data(), owners.data(), first, count, plain.data()); });
    uint32_t counter = 0;
    // optional start order (fifth argument != 0): a permutation of the slots, here simply reversed with a stride
    std::vector<uint32_t> schedule;
    if (argc > 4 && std::atoi(argv[4])) { schedule.resize(count); for (uint32_t u = 0; u < count; u++) schedule[u] = uint32_t((uint64_t(count - 1 - u) * 7919u) % count); }
    const unsigned blocks = std::min<unsigned>((count + 127) / 128, 148u * 8u);
    simt::launch(blocks, 128, [&] { sampleOwnersRefillKernel(mesh, centerHalf.data(), owners.data(), first, count, refill.data(), &counter, leafBatch, schedule.empty() ? nullptr : schedule.data()); });
    if (counter < count) { std::fprintf(stderr, "counter %u < count %u\n", counter, count); return 1; }
    for (uint32_t u = 0; u < count; u++) {   // parked nearest triangle must be a valid id before the finishing pass
        int t;
        std::memcpy(&t, &refill[u].x, 4);
        if (t < 0 || uint32_t(t) >= mesh.numTriangles) { std::fprintf(stderr, "sample %u: parked triangle %d\n", u, t); return 1; }
    }
    simt::launch((count + 255) / 256, 256, [&] { finishOwnersKernel(mesh, centerHalf.data(), owners.data(), first, count, refill.data()); });
    if (std::memcmp(plain.data(), refill.data(), count * sizeof(float4)) != 0) {
        for (uint32_t u = 0; u < count; u++)
            if (std::memcmp(&plain[u], &refill[u], sizeof(float4)) != 0) { std::fprintf(stderr, "sample %u differs: %g vs %g\n", u, plain[u].x, refill[u].x); break; }
        return 1;
    }
    std::printf("ok %u samples identical\n", count);
    return 0;
}
