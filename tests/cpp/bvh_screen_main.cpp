// The float-screened BVH traversal of sdflib_b200/csrc/bvh_sampler.cuh (bvhFastInnerStep / bvhFastLeafStep: float32 decisions
// with an error margin, float64 fallback, speculative pops) against the reference-order traversal (bvhNearest) ON THE CPU,
// lane by lane from the same source: millions of sample positions on a displaced icosphere — uniform ones, positions snapped
// to a coarse lattice (many exact ties between triangles), positions on vertices, edge mid-points and sphere centres of the
// tree — must end on the same triangle. Also prints how many visits needed the float64 path.
//
//   bvh_screen_main <isosphere subdivisions> <samples>      prints "ok <samples> identical, <visits> visits, <exact> exact"
#include <atomic>
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
thread_local Dim tThread, tBlock, tBlockDim;
thread_local uint64_t tExact = 0, tVisits = 0;
}  // namespace simt
#define __launch_bounds__(...)
#define threadIdx simt::tThread
#define blockIdx simt::tBlock
#define blockDim simt::tBlockDim
#define SDFB_BVH_SCREEN_STATS simt::tExact++
template <class T> inline T __shfl_sync(unsigned, T v, int) { return v; }
inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

namespace sdfb200 {
namespace {
alignas(16) double bvhStackSmem[64 * 128 * 2];
#include "bvh_sampler.cuh"

uint32_t nearestScreened(const DeviceMesh& m, f3 pf) {
    const BvhStack st = bvhStackOfThread(m);
    BvhFastCursor f;
    f.c.active = false;
    bvhFastStart(m, f, pf);
    while (f.c.active) {
        simt::tVisits++;
        if (f.c.cur >= 0) bvhFastInnerStep(m, f, st);
        else bvhFastLeafStep(m, f, st);
    }
    return uint32_t(f.c.bestTri);
}
}
}  // namespace sdfb200

using namespace sdfb200;

int main(int argc, char** argv) {
    const uint32_t subdivisions = argc > 1 ? uint32_t(std::atoi(argv[1])) : 5;
    const uint32_t nSamples = argc > 2 ? uint32_t(std::atoi(argv[2])) : 200000;
    uint32_t nv = 0, ni = 0;
    sdfb200_make_isosphere(subdivisions, nullptr, nullptr, &nv, &ni);
    std::vector<float> verts(size_t(nv) * 3);
    std::vector<uint32_t> idx(ni);
    sdfb200_make_isosphere(subdivisions, verts.data(), idx.data(), &nv, &ni);
    for (uint32_t v = 0; v < nv; v++) {
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

    const unsigned nThreads = std::min(64u, std::max(1u, std::thread::hardware_concurrency()));
    std::atomic<uint64_t> visits{0}, exact{0}, bad{0};
    std::vector<std::thread> pool;
    for (unsigned w = 0; w < nThreads; w++)
        pool.emplace_back([&, w] {
            simt::tThread.x = w; simt::tBlockDim.x = 128;
            uint32_t rng = 7919u * (w + 1);
            auto next = [&] { rng = rng * 1664525u + 1013904223u; return float(rng >> 8) / float(1 << 24); };
            for (uint32_t s = w; s < nSamples; s += nThreads) {
                f3 p = mk3(3.0f * next() - 1.5f, 3.0f * next() - 1.5f, 3.0f * next() - 1.5f);
                const uint32_t kind = s % 8u;
                if (kind == 1) { p.x = std::round(p.x * 8.f) / 8.f; p.y = std::round(p.y * 8.f) / 8.f; p.z = std::round(p.z * 8.f) / 8.f; }   // lattice: exact ties
                if (kind == 2) { const uint32_t v = uint32_t(next() * float(nv)) % nv; p = mk3(verts[3 * v], verts[3 * v + 1], verts[3 * v + 2]); }   // on a vertex
                if (kind == 3) {   // an edge mid-point
                    const uint32_t t = uint32_t(next() * float(ni / 3)) % (ni / 3);
                    const float* a = &verts[size_t(idx[3 * t]) * 3]; const float* b = &verts[size_t(idx[3 * t + 1]) * 3];
                    p = mk3(0.5f * (a[0] + b[0]), 0.5f * (a[1] + b[1]), 0.5f * (a[2] + b[2]));
                }
                if (kind == 4) {   // (the float image of) a sphere centre of the tree
                    const BvhNode& nd = bvh[uint32_t(next() * float(bvh.size())) % bvh.size()];
                    if (!nd.pad[0]) p = mk3(float(nd.lc[0]), float(nd.lc[1]), float(nd.lc[2]));
                }
                if (kind == 5) p = p * 0.02f;   // deep inside: long traversals, many near-equal candidates
                const uint32_t a = bvhNearest(mesh, p), b = nearestScreened(mesh, p);
                if (a != b) { if (bad.fetch_add(1) < 5) std::fprintf(stderr, "sample %u (kind %u): triangle %u vs %u\n", s, kind, a, b); }
            }
            visits += simt::tVisits; exact += simt::tExact;
        });
    for (std::thread& t : pool) t.join();
    if (bad.load()) { std::fprintf(stderr, "%llu samples differ\n", (unsigned long long)bad.load()); return 1; }
    std::printf("ok %u identical, %llu visits, %llu exact\n", nSamples, (unsigned long long)visits.load(), (unsigned long long)exact.load());
    return 0;
}
