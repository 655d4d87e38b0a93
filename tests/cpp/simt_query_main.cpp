// Runs the OctreeSdf query kernels of sdflib_b200/csrc/octree_query_kernels.cuh ON THE CPU, from the very same source,
// under a lock-step warp emulation: every lane of a warp is a host thread, and the warp collectives (__ballot_sync,
// __shfl_sync, __shfl_xor_sync) are barrier exchanges between the 32 threads. Purpose: the experimental kernels
// (tile kernel: top index, division-free cell selection, quad-cooperative evaluation) this checks the control flow —
// rounds, leader choice, class membership, index packing, partial last warp — against the plain kernel and (in the
// pytest wrapper, tests/test_query_variants_model.py) against the oracle. Host fmaf is a true fused multiply-add and
// the file is compiled with -ffp-contract=off, so the arithmetic is the device's FMA-kernel arithmetic.
//
//   simt_query_main <in.bin> <out.bin> [top index levels]
//   in : 6 f32 box, f32 cell, i32 grid, f32 minBorder, u32 maxDepth, u64 nWords, words, u64 nPoints, points (xyz f32)
//   out: for each of plain, tile: distances (n f32), then distances + gradients of the gradient kernels
// It also compares the tile kernel's division-free cell selection (cellCoordinate / floorSmall) with IEEE division and
// floorf on every coordinate of the batch and on 2^22 random ones; any mismatch fails the run.
// With -DSDFB_QUERY_EXACT the same header yields the reference-order kernels (no cooperative one): their output must
// equal the oracle bit for bit, which also pins the emulation itself (host arithmetic = device arithmetic, -fmad=false).
#include <barrier>
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "sdf_internal.h"   // f3 algebra (tri_math.cuh), node word constants; __device__ & co. are inert attributes for g++

namespace simt {
struct Dim { unsigned x = 0, y = 0, z = 0; };
struct Warp {
    std::barrier<> bar{32};
    uint64_t slot[32];
};
thread_local Dim tThread, tBlock, tBlockDim, tGridDim;
thread_local Warp* tWarp = nullptr;
thread_local unsigned tLane = 0;

template <class T> T exchange(T v, unsigned src) {   // every lane publishes v, reads lane src's
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

// grid x block launch, one warp at a time; blockDim must be a multiple of 32 (the kernels use 256)
template <class K> void launch(unsigned grid, unsigned block, K kernel) {
    for (unsigned b = 0; b < grid; b++)
        for (unsigned w = 0; w < block / 32; w++) {
            Warp warp;
            std::vector<std::thread> lanes;
            for (unsigned l = 0; l < 32; l++)
                lanes.emplace_back([&, b, w, l] {
                    tWarp = &warp; tLane = l;
                    tThread.x = w * 32 + l; tBlock.x = b; tBlockDim.x = block; tGridDim.x = grid;
                    kernel();
                    warp.bar.arrive_and_drop();   // a lane that left (early return / end) no longer takes part
                });
            for (std::thread& t : lanes) t.join();
        }
}
}  // namespace simt

#define __launch_bounds__(...)
#define SDFB_SIMT_HOST 1
#define gridDim simt::tGridDim
#define threadIdx simt::tThread
#define blockIdx simt::tBlock
#define blockDim simt::tBlockDim
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return simt::exchange(v, unsigned(src)); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return simt::exchange(v, simt::tLane ^ unsigned(m)); }
inline unsigned __ballot_sync(unsigned, bool p) { return simt::ballot(p); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline void __syncwarp() { simt::tWarp->bar.arrive_and_wait(); }
// the rounding-mode intrinsics: volatile operands keep g++ from folding or contracting them (-ffp-contract=off anyway)
inline float __fmul_rn(float a, float b) { volatile float x = a, y = b; return x * y; }
inline float __fadd_rn(float a, float b) { volatile float x = a, y = b; return x + y; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fadd_rd(float a, float b) {
    const int old = std::fegetround();
    std::fesetround(FE_DOWNWARD);
    volatile float x = a, y = b;
    volatile float r = x + y;
    std::fesetround(old);
    return r;
}
inline uint32_t atomicOr(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }

namespace sdfb200 {
namespace {
#include "octree_query_kernels.cuh"
}
}  // namespace sdfb200

using namespace sdfb200;

template <class T> static void readVec(FILE* f, std::vector<T>& v) {
    uint64_t n = 0;
    if (std::fread(&n, 8, 1, f) != 1) std::exit(2);
    v.resize(n);
    if (n && std::fread(v.data(), sizeof(T), n, f) != n) std::exit(2);
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    float box[6], cell, minBorder;
    int grid;
    uint32_t maxDepth;
    if (std::fread(box, 4, 6, f) != 6 || std::fread(&cell, 4, 1, f) != 1 || std::fread(&grid, 4, 1, f) != 1 ||
        std::fread(&minBorder, 4, 1, f) != 1 || std::fread(&maxDepth, 4, 1, f) != 1) return 2;
    std::vector<uint32_t> oct;
    std::vector<float> xyz;
    readVec(f, oct);
    readVec(f, xyz);
    std::fclose(f);
    const uint64_t n = xyz.size() / 3;
    QueryParams q;
    q.minx = box[0]; q.miny = box[1]; q.minz = box[2];
    q.maxx = box[3]; q.maxy = box[4]; q.maxz = box[5];
    q.cell = cell; q.grid = grid; q.minBorder = minBorder;
    const unsigned blocks = unsigned((n + 255) / 256);

    // top index exactly as prepareOctreeQuery (octree_query.cu) sizes it
    int startDepth = 0;
    while ((1 << startDepth) < grid) startDepth++;
    int levels = std::min(std::max(0, int(maxDepth) - startDepth), 15);
    while (levels > 0 && 3 * (startDepth + levels) > 21) levels--;
    if (argc > 3) levels = std::min(levels, std::atoi(argv[3]));   // shallower index: the finishing descent gets exercised
    const uint64_t cells = uint64_t(1) << (3 * (startDepth + levels));
    std::vector<uint32_t> index(cells + 1, 0u);
#ifndef SDFB_QUERY_EXACT
    simt::launch(unsigned((cells + 255) / 256), 256, [&] { topIndexKernel(oct.data(), grid, levels, index.data(), index.data() + cells); });
    if (index[cells]) { std::fprintf(stderr, "topIndexKernel flagged the array\n"); return 1; }
    TileQuery tq;
    tq.cellL = std::ldexp(cell, -levels); tq.rcellL = 1.0f / tq.cellL;
    { const float limit = float(uint32_t(grid) << levels); std::memcpy(&tq.limitBits, &limit, 4); }
    tq.shiftN = startDepth + levels; tq.topLevels = levels; tq.G3 = uint32_t(grid) * uint32_t(grid) * uint32_t(grid);
    {   // division-free cell selection against IEEE division / floorf (at the start-grid scale and at the index scale)
        uint64_t bad = 0, state = 0x9E3779B97F4A7C15ull;
        auto check = [&](float x) {
            for (int L = 0; L <= levels; L += std::max(levels, 1)) {
                const float cl = std::ldexp(cell, -L);
                const float want = x / cl, got = cellCoordinate(x, cl, 1.0f / cl);
                if (std::fabs(x) > 1e-30f && __float_as_uint(want) != __float_as_uint(got)) bad++;
                if (__float_as_uint(want) != __float_as_uint(std::ldexp(x / cell, L))) bad++;   // the scaling itself is exact
                if (want >= 0.0f && want < 4194304.0f) {
                    if (fractSmall(want) != want - std::floor(want)) bad++;
                    if ((uint32_t(__float_as_int(__fadd_rd(want, 8388608.0f))) & 0x7FFFFFu) != uint32_t(std::floor(want))) bad++;
                }
            }
        };
        for (uint64_t i = 0; i < 3 * n; i++) check(xyz[i] - box[i % 3]);
        for (int i = 0; i < (1 << 22); i++) {
            state ^= state << 13; state ^= state >> 7; state ^= state << 17;
            const float u = float(state >> 40) * (1.0f / 16777216.0f);
            check((i & 1) ? u * 1.1f * float(grid) * cell : float(state % uint64_t(64 * grid)) / 64.0f * cell);   // uniform / near cell and sub-cell boundaries
        }
        check(0.0f); check(-0.0f);
        if (__float_as_uint(cellCoordinate(-0.0f, cell, 1.0f / cell)) != 0u) bad++;   // the unsigned range test relies on +0
        if (bad) { std::fprintf(stderr, "cell selection differs from IEEE division / floor on %llu values\n", (unsigned long long)bad); return 1; }
    }
#endif

    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 2;
    std::vector<float> dist(n), grad(3 * n);
    auto emit = [&](bool withGrad) {
        std::fwrite(dist.data(), 4, n, o);
        if (withGrad) std::fwrite(grad.data(), 4, 3 * n, o);
    };
    const float* pts = xyz.data();
    // plain
    simt::launch(blocks, 256, [&] { octreeQueryKernel<false, true>(oct.data(), q, pts, n, dist.data(), nullptr); });
    emit(false);
    simt::launch(blocks, 256, [&] { octreeQueryKernel<true, true>(oct.data(), q, pts, n, dist.data(), grad.data()); });
    emit(true);
#ifndef SDFB_QUERY_EXACT   // compiled with -DSDFB_QUERY_EXACT the header holds the reference-order kernel only
    // tile kernel (plain loads instead of the TMA staging): 3 CTAs of 8 warps, so every warp loops over several tiles
    std::fill(dist.begin(), dist.end(), -123.0f);
    simt::launch(3, 256, [&] { octreeQueryTileKernel<false, true>(oct.data(), index.data(), q, tq, pts, n, dist.data(), nullptr); });
    emit(false);
    {   // the scalar form of the same kernel (SDFB200_QUERY_PACKED=0) gives the same bits: packing only pairs up operations
        std::vector<float> scalar(n, -123.0f);
        simt::launch(3, 256, [&] { octreeQueryTileKernel<false, false>(oct.data(), index.data(), q, tq, pts, n, scalar.data(), nullptr); });
        if (std::memcmp(scalar.data(), dist.data(), n * 4) != 0) { std::fprintf(stderr, "packed and scalar tile kernels differ\n"); return 1; }
    }
    std::fill(dist.begin(), dist.end(), -123.0f);
    std::fill(grad.begin(), grad.end(), -123.0f);
    simt::launch(3, 256, [&] { octreeQueryTileKernel<true, false>(oct.data(), index.data(), q, tq, pts, n, dist.data(), grad.data()); });
    emit(true);
#endif
    std::fclose(o);
    std::printf("ok %llu queries, index levels %d\n", (unsigned long long)n, levels);
    return 0;
}
