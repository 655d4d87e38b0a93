// ExactOctreeSdf construction on the GPU (hot path 1, exact variant).
//
// Replaces, behind sdfb200_build_exact, the reference's depth-first builder
//   ExactOctreeSdf::initOctree<PerNodeRegionTrianglesInfluence>   include/SdfLib/ExactOctreeSdfDepthFirst.h:28-651
//   PerNodeRegionTrianglesInfluence::filterTriangles                include/SdfLib/TrianglesInfluence.h:767-860  (HOT LOOP A)
//   GJK::IsNearMinimize / findFurthestPoint                         src/utils/GJK.cpp:830-866, :715-738, :644-652
//   PerNodeRegionTrianglesInfluence::calculateVerticesInfo          include/SdfLib/TrianglesInfluence.h:693-765  (HOT LOOP B)
//   TriangleUtils::getSqDistPointAndTriangle                        include/SdfLib/utils/TriangleUtils.h:76-135
//
// B200 design (not a translation of the CPU stack machine):
//   * level-synchronous and FLAT: the unit of work is a (node, parent-list entry) pair, not a node, so the
//     250 000-triangle lists of interior nodes and the 129-triangle lists of surface nodes load the SMs
//     equally. One launch filters every pair of a depth (Frank-Wolfe per thread), an exclusive scan of the
//     keep flags gives the order-preserving compaction (lists stay ascending, as the reference's do);
//   * HOT LOOP B: a CTA owns a 1024-entry chunk of one subdividing node's list; the chunk's index block is
//     staged into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier) while the 19 sample
//     points are set up, every lane gathers its triangles' 80-byte frames as 5 x 128-bit loads, keeps the 19
//     running minima in registers and the CTA folds them with redux.sync (distance bits, then lowest list
//     position among equals = the serial loop's first strict minimum) into one 64-bit atomicMin per sample;
//   * the reference's post-order merge (children lists -> union list + 8 bit masks, only at depths >=
//     maxDepth-2) is two flat passes over the stored keep flags — no 8-way merge loop;
//   * array ORDER (stack discipline of the reference: block appended at first visit, children popped 7
//     first, masks / union sets appended at the second visit) is rebuilt from subtree sizes: bottom-up size
//     pass, top-down offset pass, emit passes for nodes / packed sets / masks.
// Arithmetic is bit-faithful to the CPU build (this TU is compiled with -fmad=false); the reference's
// history-dependent 32^3 vertex cache is not emulated (DESIGN.md "parity").
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "device_utils.cuh"
#include "sdf_internal.h"

namespace sdfb200 {

namespace {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int kChunk = 512;           // list entries per CTA of the sample kernel
constexpr int kSampleThreads = 128;
constexpr uint64_t kNoKey = ~uint64_t(0);

__constant__ int cMidLattice[19] = {1, 3, 4, 5, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 19, 21, 22, 23, 25};

// ---- per-level node arrays -----------------------------------------------------------------------------
struct LevelView {
    uint32_t count;
    uint32_t depth;
    const float4* centerHalf;     // xyz centre, w half size
    const uint32_t* info;         // 8 per node: nearest triangle of each corner (within the parent's list)
    const uint32_t* parentLo;     // range of the parent's list inside the parent level's list array
    const uint32_t* parentCnt;
    const uint64_t* pairOff;      // count + 1: exclusive scan of parentCnt
    const uint32_t* listLo;       // range of the node's own (filtered) list inside this level's list array
    const uint32_t* listCnt;
};

__device__ __forceinline__ TriFrame loadFrame(const float4* __restrict__ frames, uint32_t t) {
    const float4 a = __ldg(frames + 5 * size_t(t)), b = __ldg(frames + 5 * size_t(t) + 1), c = __ldg(frames + 5 * size_t(t) + 2),
                 d = __ldg(frames + 5 * size_t(t) + 3), e = __ldg(frames + 5 * size_t(t) + 4);
    TriFrame f;
    f.ox = a.x; f.oy = a.y; f.oz = a.z; f.t00 = a.w;
    f.t01 = b.x; f.t02 = b.y; f.t10 = b.z; f.t11 = b.w;
    f.t12 = c.x; f.t20 = c.y; f.t21 = c.z; f.t22 = c.w;
    f.bx = d.x; f.by = d.y; f.cx = d.z; f.cy = d.w;
    f.v2 = e.x; f.v3x = e.y; f.v3y = e.z;
    return f;
}

template <class T> __device__ __forceinline__ uint32_t lastLessEqual(const T* a, uint32_t n, T key) {
    // largest i in [0, n) with a[i] <= key (a non-decreasing, a[0] <= key)
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] <= key) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- HOT LOOP A, part 1: region radii of a node (TrianglesInfluence.h:784-800) -------------------------
// 64 threads per node: thread (i, c) evaluates the distance of corner c to the nearest triangle of corner i.
__global__ void __launch_bounds__(256)
regionKernel(const float4* __restrict__ frames, LevelView lv, float* __restrict__ region /* 72 per node */) {
    const uint32_t node = blockIdx.x * 4 + (threadIdx.x >> 6);
    if (node >= lv.count) return;   // whole 64-thread group leaves together (two full warps)
    const uint32_t e = threadIdx.x & 63u, i = e >> 3, c = e & 7u;
    const float4 ch = lv.centerHalf[node];
    const TriFrame f = loadFrame(frames, lv.info[size_t(node) * 8 + i]);
    const f3 p = mk3(ch.x, ch.y, ch.z) + cornerDir(c) * ch.w;
    const float rho = sqrtf(sqDistPointTriangle(p, f));
    float m = rho;
    for (int o = 4; o > 0; o >>= 1) m = gmin(m, __shfl_xor_sync(0xffffffffu, m, o));   // groups of 8 lanes
    region[size_t(node) * 72 + e] = rho - m;
    if (c == 0) region[size_t(node) * 72 + 64 + i] = m;
}

// ---- HOT LOOP A, part 2: Frank-Wolfe proximity test (a9, src/utils/GJK.cpp:830-866) of every (node, parent-list
// entry) pair, with lane refilling. Trip counts range from 1 to 15, so with one pair per thread a warp runs at the
// pace of its slowest pair (the first version of this kernel: 11.7 of 32 lanes active per instruction on the
// deepest C3 level, 28 ms of the 64 ms build; this one: C3 levels 52.4 -> 39.1 ms, identical flags).
// Here a warp owns a chunk of consecutive pairs and every lane keeps ONE Frank-Wolfe iteration per loop trip in
// flight: a lane whose pair has finished takes the next unassigned pair of the chunk (ballot + prefix), so the
// iteration body always runs with (almost) all lanes. Per-pair arithmetic is unchanged.
constexpr uint32_t kFilterChunk = 1024;   // pairs per warp

__global__ void __launch_bounds__(256)
filterRefillKernel(DeviceMesh mesh, LevelView lv, const uint32_t* __restrict__ parentList, const float* __restrict__ region,
                   uint8_t* __restrict__ flags, uint64_t numPairs) {
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint64_t begin = (uint64_t(blockIdx.x) * 8 + (threadIdx.x >> 5)) * kFilterChunk;
    if (begin >= numPairs) return;
    const uint64_t end = begin + kFilterChunk < numPairs ? begin + kFilterChunk : numPairs;
    uint32_t nodeCur = 0;
    if (lane == 0) nodeCur = lastLessEqual<uint64_t>(lv.pairOff, lv.count + 1, begin);
    nodeCur = __shfl_sync(kFull, nodeCur, 0);
    uint64_t next = begin;
    bool have = false;
    uint64_t myPair = 0;
    float half = 0.0f, thr = 0.0f, sqThr = 0.0f;
    float r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0, r5 = 0, r6 = 0, r7 = 0;
    f3 a = mk3(0, 0, 0), b = a, c = a, x = a;
    uint32_t iter = 0;
    for (;;) {
        const unsigned need = __ballot_sync(kFull, !have);
        if (need && next < end) {
            const uint32_t avail = uint32_t(end - next);
            const uint32_t want = uint32_t(__popc(need));
            const uint32_t take = want < avail ? want : avail;
            const uint32_t rank = uint32_t(__popc(need & ((1u << lane) - 1u)));
            uint32_t node = nodeCur;
            if (!have && rank < take) {
                myPair = next + rank;
                while (myPair >= lv.pairOff[node + 1]) node++;
                const uint32_t j = uint32_t(myPair - lv.pairOff[node]);
                const uint32_t t = parentList[lv.parentLo[node] + j];
                const float4 ch = lv.centerHalf[node];
                const f3 ctr = mk3(ch.x, ch.y, ch.z);
                a = mesh.verts[mesh.idx[3 * size_t(t)]] - ctr;
                b = mesh.verts[mesh.idx[3 * size_t(t) + 1]] - ctr;
                c = mesh.verts[mesh.idx[3 * size_t(t) + 2]] - ctr;
                const f3 g = 0.3333333f * ((a + b) + c);
                const uint32_t v = ((g.z > 0) ? 4u : 0u) + ((g.y > 0) ? 2u : 0u) + ((g.x > 0) ? 1u : 0u);
                if (lv.info[size_t(node) * 8 + v] == t) flags[myPair] = 1;   // the corner's own nearest triangle is always kept
                else {
                    const float4 q0 = *reinterpret_cast<const float4*>(region + size_t(node) * 72 + v * 8);
                    const float4 q1 = *reinterpret_cast<const float4*>(region + size_t(node) * 72 + v * 8 + 4);
                    r0 = q0.x; r1 = q0.y; r2 = q0.z; r3 = q0.w; r4 = q1.x; r5 = q1.y; r6 = q1.z; r7 = q1.w;
                    half = ch.w;
                    thr = region[size_t(node) * 72 + 64 + v];
                    sqThr = thr * thr;
                    x = -a;
                    iter = 0;
                    have = true;
                }
            }
            nodeCur = __reduce_max_sync(kFull, node);
            next += take;
        }
        if (!__any_sync(kFull, have)) {
            if (next >= end) break;
            continue;
        }
        if (have) {   // one iteration of GJK::IsNearMinimize (src/utils/GJK.cpp:830-866)
            const f3 g = normalize3(-x);
            float best = dot3(mk3(-half, -half, -half), g) + r0;
            uint32_t bi = 0;
            float br = r0;
#define SDFB_SUPPORT(i, r)                                            \
            {                                                         \
                const float v = dot3(cornerDir(i) * half, g) + (r);   \
                if (v > best) { best = v; bi = (i); br = (r); }       \
            }
            SDFB_SUPPORT(1u, r1) SDFB_SUPPORT(2u, r2) SDFB_SUPPORT(3u, r3) SDFB_SUPPORT(4u, r4)
            SDFB_SUPPORT(5u, r5) SDFB_SUPPORT(6u, r6) SDFB_SUPPORT(7u, r7)
#undef SDFB_SUPPORT
            const f3 boxPoint = cornerDir(bi) * half + br * g;
            const f3 ng = -g;
            const float d1 = dot3(a, ng), d2 = dot3(b, ng), d3v = dot3(c, ng);
            const f3 triPoint = (d1 > d2) ? ((d1 > d3v) ? a : c) : ((d2 > d3v) ? b : c);
            const f3 p = boxPoint - triPoint;
            const float distToP = dot3(g, p - x);
            const float distToO = dot3(g, -x);
            const f3 dir = p - x;
            const float d = dot3(dir, -x);
            bool done, result;
            if (double(d) < 1.0e-5) { done = true; result = distToO <= distToP + thr; }
            else {
                x = x + dir * gmin(d / dot3(dir, dir), 1.0f);
                const bool isNear = dot3(x, x) < sqThr;
                bool again = false;
                if (!isNear && distToO <= distToP + thr) { iter++; again = iter < 15; }
                done = !again;
                result = isNear || iter >= 15;
            }
            if (done) { flags[myPair] = result ? 1 : 0; have = false; }
        }
    }
}

// Order-preserving compaction of the kept pairs. 16 pairs per thread: the flags arrive as one 128-bit load and
// about 7 of 8 threads leave right there (13 % of the pairs survive the filter on C3); the others look their
// node up once per warp (binary search by lane 0 for the warp's first pair) and walk forward from it.
__global__ void __launch_bounds__(256)
compactKernel(LevelView lv, const uint32_t* __restrict__ parentList, const uint8_t* __restrict__ flags,
              const uint32_t* __restrict__ pos, uint32_t* __restrict__ list, uint64_t numPairs) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = (uint64_t(blockIdx.x) * 256 + (threadIdx.x & ~31u)) * 16;
    if (warp0 >= numPairs) return;
    const uint64_t p0 = warp0 + uint64_t(lane) * 16;
    uint4 f = make_uint4(0, 0, 0, 0);
    if (p0 < numPairs) f = loadFlags16(flags, p0, numPairs);
    const bool any = (f.x | f.y | f.z | f.w) != 0;
    if (!__any_sync(0xffffffffu, any)) return;
    uint32_t node = 0;
    if (lane == 0) node = lastLessEqual<uint64_t>(lv.pairOff, lv.count + 1, warp0);
    node = __shfl_sync(0xffffffffu, node, 0);
    if (!any) return;
    const uint32_t words[4] = {f.x, f.y, f.z, f.w};
    uint32_t at = pos[p0 >> 4];                                            // group positions (FlagScanner::runGroups)
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (!((words[k >> 2] >> (8 * (k & 3))) & 1u)) continue;
        const uint64_t p = p0 + k;
        while (p >= lv.pairOff[node + 1]) node++;
        list[at++] = parentList[lv.parentLo[node] + uint32_t(p - lv.pairOff[node])];
    }
}

__global__ void listRangeKernel(const uint64_t* pairOff, const uint8_t* flags, const uint32_t* pos, uint64_t numPairs, uint32_t total, uint32_t* listLo,
                                uint32_t* listCnt, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t a = pairOff[i], b = pairOff[i + 1];
    const uint32_t lo = a < numPairs ? flagPositionAt(flags, pos, a, numPairs) : total, hi = b < numPairs ? flagPositionAt(flags, pos, b, numPairs) : total;
    listLo[i] = lo;
    listCnt[i] = hi - lo;
}

// terminal rule (ExactOctreeSdfDepthFirst.h:307-314) + chunk counts of the sample pass
__global__ void decideKernel(const uint32_t* listCnt, uint32_t n, uint32_t depth, uint32_t startDepth, uint32_t maxDepth,
                             uint32_t minTris, uint32_t* subdivide, uint32_t* chunks, uint32_t* maxTrisInLeafs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t cnt = listCnt[i];
    const bool terminal = depth >= startDepth && cnt <= minTris;
    const bool sub = !terminal && depth < maxDepth;
    subdivide[i] = sub ? 1u : 0u;
    chunks[i] = sub ? (cnt + kChunk - 1) / kChunk : 0u;
    if (!sub) atomicMax(maxTrisInLeafs, cnt);
}

// ---- HOT LOOP B: nearest list entry for the kPts sample points of a node ---------------------------------
// grid = total chunks; chunkOff (count + 1) maps a chunk to its node. best[node * kPts + s] receives
// min over the list of (sqDist bits << 32 | list position).
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(uint32_t(__cvta_generic_to_shared(bar))), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(uint32_t(__cvta_generic_to_shared(bar))), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(uint32_t(__cvta_generic_to_shared(bar))), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared (16-byte aligned addresses and size)
__device__ __forceinline__ void tmaLoad1d(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     uint32_t(__cvta_generic_to_shared(smemDst))),
                 "l"(gmemSrc), "r"(bytes), "r"(uint32_t(__cvta_generic_to_shared(bar)))
                 : "memory");
}

// One WARP per chunk of 512 list entries (a CTA of four warps works on four chunks, usually four different nodes). With
// a CTA per chunk (round 1) a 129-500 entry list filled its 128 threads for one to four trips, the last one partly: 14 of 32
// threads per instruction (profiles/r1_summary.md). A warp walks its chunk in trips of 32 entries — at most the last trip of a
// chunk is partly filled — and the 19 (distance, position) minima are folded once per chunk instead of once per warp and CTA.
template <int kPts>
__global__ void __launch_bounds__(kSampleThreads)
sampleKernel(const float4* __restrict__ frames, const float4* __restrict__ centerHalf, const uint32_t* __restrict__ list,
             const uint32_t* __restrict__ listLo, const uint32_t* __restrict__ listCnt, const uint32_t* __restrict__ chunkOff,
             uint32_t numNodes, uint32_t numChunks, unsigned long long* __restrict__ best) {
    constexpr int kWarps = kSampleThreads / 32;
    // The chunk's index block is fetched as the 16-byte aligned window that covers it: [winLo, winLo + winBytes)
    __shared__ alignas(16) uint32_t sIdx[kWarps][kChunk + 8];
    __shared__ alignas(8) uint64_t sBar[kWarps];
    __shared__ float sPts[kWarps][kPts][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x * kWarps + warp;
    if (chunk >= numChunks) return;                                       // warp-uniform; no CTA-wide barrier below
    uint32_t node = 0;
    if (lane == 0) {
        node = lastLessEqual<uint32_t>(chunkOff, numNodes + 1, chunk);
        mbarInit(&sBar[warp], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    node = __shfl_sync(0xffffffffu, node, 0);
    const uint32_t first = (chunk - chunkOff[node]) * kChunk;            // position of the chunk inside the node's list
    const uint32_t cnt = min(uint32_t(kChunk), listCnt[node] - first);   // entries of this chunk
    const size_t gFirst = size_t(listLo[node]) + first;                   // index of the first entry in `list`
    const size_t winFirst = gFirst & ~size_t(3);                          // 16-byte aligned window start
    const uint32_t skip = uint32_t(gFirst - winFirst);
    const uint32_t winBytes = ((skip + cnt + 3u) & ~3u) * 4u;             // the list array is padded by 8 words
    if (lane == 0) {
        mbarExpectTx(&sBar[warp], winBytes);
        tmaLoad1d(sIdx[warp], list + winFirst, winBytes, &sBar[warp]);
    }
    if (lane < kPts) {
        const float4 ch = centerHalf[node];
        f3 rel;
        if (kPts == 8) rel = cornerDir(uint32_t(lane));
        else {
            const int L = cMidLattice[lane];
            rel = mk3(float(L % 3 - 1), float((L / 3) % 3 - 1), float(L / 9 - 1));
        }
        const f3 p = mk3(ch.x, ch.y, ch.z) + rel * ch.w;
        sPts[warp][lane][0] = p.x; sPts[warp][lane][1] = p.y; sPts[warp][lane][2] = p.z;
    }
    __syncwarp();
    mbarWait(&sBar[warp], 0);

    float bestD[kPts];
    uint32_t bestJ[kPts];
#pragma unroll
    for (int s = 0; s < kPts; s++) { bestD[s] = INFINITY; bestJ[s] = kNone; }
    for (uint32_t k = uint32_t(lane); k < cnt; k += 32) {                 // ascending positions per thread
        const TriFrame f = loadFrame(frames, sIdx[warp][skip + k]);
#pragma unroll
        for (int s = 0; s < kPts; s++) {
            const float d = sqDistPointTriangle(mk3(sPts[warp][s][0], sPts[warp][s][1], sPts[warp][s][2]), f);
            if (d < bestD[s]) { bestD[s] = d; bestJ[s] = first + k; }
        }
    }
    // fold: distance bits first (non-negative floats order like their bit patterns), then the lowest list
    // position among the lanes that hold the minimum = first strict minimum of the serial ascending loop
    unsigned long long mine = kNoKey;
#pragma unroll
    for (int s = 0; s < kPts; s++) {
        const uint32_t bits = bestJ[s] == kNone ? 0xFFFFFFFFu : __float_as_uint(bestD[s]);
        const uint32_t m = __reduce_min_sync(0xffffffffu, bits);
        const uint32_t j = __reduce_min_sync(0xffffffffu, bits == m ? bestJ[s] : kNone);
        if (lane == s) mine = (static_cast<unsigned long long>(m) << 32) | j;
    }
    if (lane < kPts && (mine >> 32) != 0xFFFFFFFFull) atomicMin(best + size_t(node) * kPts + lane, mine);
}

// best keys -> triangle ids (0 when the list was empty: deterministic stand-in for the reference's untouched slot)
__global__ void resolveKernel(const unsigned long long* best, const uint32_t* list, const uint32_t* listLo, uint32_t perNode,
                              uint32_t* out, uint64_t n) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = best[i];
    out[i] = (k == kNoKey) ? 0u : list[size_t(listLo[i / perNode]) + uint32_t(k & 0xFFFFFFFFull)];
}

// children of the subdividing nodes (ExactOctreeSdfDepthFirst.h:329-440): 64 threads per parent = (child c, corner k)
__global__ void __launch_bounds__(256)
childrenKernel(LevelView lv, const uint32_t* __restrict__ coord, const uint32_t* __restrict__ subdivide,
               const uint32_t* __restrict__ childIdx, const uint32_t* __restrict__ midInfo, uint32_t* __restrict__ childOf,
               float4* __restrict__ nCenterHalf, uint32_t* __restrict__ nInfo, uint32_t* __restrict__ nCoord,
               uint32_t* __restrict__ nParent, uint32_t* __restrict__ nParentLo, uint32_t* __restrict__ nParentCnt) {
    const uint32_t node = blockIdx.x * 4 + (threadIdx.x >> 6);
    if (node >= lv.count) return;
    const uint32_t e = threadIdx.x & 63u, c = e >> 3, k = e & 7u;
    if (!subdivide[node]) { if (e == 0) childOf[node] = kNone; return; }
    const uint32_t base = childIdx[node] * 8u;
    if (e == 0) childOf[node] = base;
    const uint32_t lx = (c & 1u) + (k & 1u), ly = ((c >> 1) & 1u) + ((k >> 1) & 1u), lz = (c >> 2) + (k >> 2);
    uint32_t v;
    if (lx != 1 && ly != 1 && lz != 1) v = lv.info[size_t(node) * 8 + ((lx >> 1) | ((ly >> 1) << 1) | ((lz >> 1) << 2))];
    else {
        const int L = int(lx + 3 * ly + 9 * lz);
        // index of lattice point L among the 19 non-corner points (ascending L)
        int s = 0;
#pragma unroll
        for (int q = 0; q < 19; q++) s += (cMidLattice[q] < L) ? 1 : 0;
        v = midInfo[size_t(node) * 19 + s];
    }
    nInfo[size_t(base + c) * 8 + k] = v;
    if (k == 0) {
        const float4 ch = lv.centerHalf[node];
        const float h = 0.5f * ch.w;
        const f3 ctr = mk3(ch.x, ch.y, ch.z) + cornerDir(c) * h;
        nCenterHalf[base + c] = make_float4(ctr.x, ctr.y, ctr.z, h);
        const uint32_t pc = coord[node];
        const uint32_t ix = ((pc & 1023u) << 1) | (c & 1u), iy = (((pc >> 10) & 1023u) << 1) | ((c >> 1) & 1u),
                       iz = (((pc >> 20) & 1023u) << 1) | (c >> 2);
        nCoord[base + c] = ix | (iy << 10) | (iz << 20);
        nParent[base + c] = node;
        nParentLo[base + c] = lv.listLo[node];
        nParentCnt[base + c] = lv.listCnt[node];
    }
}

// ---- post-order merge as flat passes (ExactOctreeSdfDepthFirst.h:189-286) ------------------------------------
// member[g] (g = global index into this level's list array): bit c set iff entry g of an inner node's list
// survives in child c's FINAL list (filtered, and merged if the child is itself inner).
struct MergeChildView {
    const uint64_t* pairOff;     // child level
    const uint8_t* flags;        // child level keep flags per pair
    const uint32_t* pos;         // child level list position per group of 16 pairs
    uint64_t numPairs;           // child level
    const uint32_t* childOf;     // child level: kNone = leaf
    const uint8_t* member;       // child level member bytes (null when the child level is the deepest one)
};
__global__ void __launch_bounds__(256)
memberKernel(const uint32_t* __restrict__ listLo, const uint32_t* __restrict__ childOf, uint32_t numNodes, uint32_t listTotal,
             MergeChildView ch, uint8_t* __restrict__ member, uint8_t* __restrict__ keep) {
    __shared__ uint32_t sNode;
    const uint32_t g0 = blockIdx.x * 256;
    if (threadIdx.x == 0) sNode = lastLessEqual<uint32_t>(listLo, numNodes, g0);
    __syncthreads();
    const uint32_t g = g0 + threadIdx.x;
    if (g >= listTotal) return;
    uint32_t node = sNode;
    while (node + 1 < numNodes && g >= listLo[node + 1]) node++;
    uint32_t m = 0;
    const uint32_t c0 = childOf[node];
    if (c0 != kNone) {
        const uint32_t j = g - listLo[node];
#pragma unroll
        for (uint32_t c = 0; c < 8; c++) {
            const uint64_t pp = ch.pairOff[c0 + c] + j;
            bool f = ch.flags[pp] != 0;
            if (f && ch.member != nullptr && ch.childOf[c0 + c] != kNone) f = ch.member[flagPositionAt(ch.flags, ch.pos, pp, ch.numPairs)] != 0;
            m |= f ? (1u << c) : 0u;
        }
    }
    member[g] = uint8_t(m);
    keep[g] = m ? 1 : 0;
}

__global__ void __launch_bounds__(256)
mergeCompactKernel(const uint32_t* __restrict__ list, const uint8_t* __restrict__ member, const uint32_t* __restrict__ mpos,
                   uint32_t listTotal, uint32_t* __restrict__ mergedList, uint8_t* __restrict__ mergedMember) {
    const uint32_t g = blockIdx.x * 256 + threadIdx.x;
    if (g >= listTotal || !member[g]) return;
    mergedList[mpos[g]] = list[g];
    mergedMember[mpos[g]] = member[g];
}

__global__ void mergedRangeKernel(const uint32_t* listLo, const uint32_t* listCnt, const uint32_t* mpos, uint32_t listTotal,
                                  uint32_t mergedTotal, uint32_t* mergedLo, uint32_t* mergedCnt, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t a = listLo[i], b = a + listCnt[i];
    const uint32_t lo = a < listTotal ? mpos[a] : mergedTotal, hi = b < listTotal ? mpos[b] : mergedTotal;
    mergedLo[i] = lo;
    mergedCnt[i] = hi - lo;
}

// ---- layout -----------------------------------------------------------------------------------------------
struct LayoutView {   // per level
    const uint32_t* childOf;
    const uint32_t* listCnt;
    const uint32_t* mergedCnt;   // null below/above the merge levels
    uint32_t* nodeCnt;           // subtree sizes: nodes (8 per inner node)
    uint32_t* setWords;          //                words of mTrianglesSets
    uint32_t* maskBytes;         //                bytes of mTrianglesMasks
    uint32_t* slot;              // where the node's own record lives
    uint32_t* nodeBase;          // start of the subtree's regions in the three arrays
    uint32_t* setBase;
    uint32_t* maskBase;
    uint32_t* ownMaskBase;       // start of the node's own 8 masks (after its inner children's)
};

__device__ __forceinline__ uint32_t packedSetWords(uint32_t n, uint32_t bits) {
    return uint32_t((uint64_t(n) * bits + 31) / 32) + 2u;   // count word + packed indices + one zero pad word
}

__global__ void sizesKernel(LayoutView lv, LayoutView next, uint32_t n, uint32_t depth, uint32_t maxDepth, uint32_t bitEnc,
                            uint32_t bits, uint32_t* maxEncoded) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c0 = lv.childOf[i];
    uint32_t nodes = 0, sets = 0, masks = 0;
    if (c0 == kNone) {
        if (depth <= bitEnc) sets = packedSetWords(lv.listCnt[i], bits);
    } else {
        nodes = 8;
        for (int c = 0; c < 8; c++) { nodes += next.nodeCnt[c0 + c]; sets += next.setWords[c0 + c]; masks += next.maskBytes[c0 + c]; }
        if (depth >= bitEnc) {
            const uint32_t m = lv.mergedCnt[i];
            masks += 8u * ((m + 7u) / 8u);
            if (depth == bitEnc) { sets = packedSetWords(m, bits); atomicMax(maxEncoded, m); }
        }
    }
    lv.nodeCnt[i] = nodes;
    lv.setWords[i] = sets;
    lv.maskBytes[i] = masks;
}

// children regions in the order the reference's stack pops them: 7 first
__global__ void offsetsKernel(LayoutView lv, LayoutView next, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c0 = lv.childOf[i];
    if (c0 == kNone) return;
    const uint32_t nb = lv.nodeBase[i];
    if (nb == kNone) {   // subtree not owned by this rank
        for (int c = 0; c < 8; c++) next.nodeBase[c0 + c] = kNone;
        return;
    }
    uint32_t rn = nb + 8u, rs = lv.setBase[i], rm = lv.maskBase[i];
    for (int c = 7; c >= 0; c--) {
        next.slot[c0 + c] = nb + uint32_t(c);
        next.nodeBase[c0 + c] = rn; rn += next.nodeCnt[c0 + c];
        next.setBase[c0 + c] = rs;  rs += next.setWords[c0 + c];
        next.maskBase[c0 + c] = rm; rm += next.maskBytes[c0 + c];
    }
    lv.ownMaskBase[i] = rm;
}

// node records: (childrenIndex | leaf, trianglesArrayIndex)
__global__ void emitNodesKernel(LayoutView lv, uint32_t n, uint32_t depth, uint32_t bitEnc, uint32_t* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t nb = lv.nodeBase[i];
    if (nb == kNone) return;
    const uint32_t c0 = lv.childOf[i], slot = lv.slot[i];
    out[2 * size_t(slot)] = c0 == kNone ? 0xFFFFFFFFu : (nb & kExactIndexMask);
    if (depth <= bitEnc && (c0 == kNone || depth == bitEnc)) out[2 * size_t(slot) + 1] = lv.setBase[i];
    if (c0 != kNone && depth >= bitEnc) {   // the second visit hands every child its mask offset
        const uint32_t bytes = (lv.mergedCnt[i] + 7u) / 8u;
        for (uint32_t c = 0; c < 8; c++) out[2 * size_t(nb + c) + 1] = lv.ownMaskBase[i] + c * bytes;
    }
}

// packed sets: [count][MSB-first indices of `bits` bits][0]   (ExactOctreeSdfDepthFirst.h:263-280, :450-467)
__global__ void __launch_bounds__(128)
emitSetsKernel(LayoutView lv, uint32_t n, uint32_t depth, uint32_t bitEnc, uint32_t bits, const uint32_t* __restrict__ srcList,
               const uint32_t* __restrict__ srcLo, const uint32_t* __restrict__ srcCnt, bool wantInner, uint32_t* __restrict__ sets) {
    const uint32_t i = blockIdx.x;
    if (i >= n || lv.nodeBase[i] == kNone) return;
    const bool inner = lv.childOf[i] != kNone;
    if (inner != wantInner) return;
    if (inner && depth != bitEnc) return;
    const uint32_t cnt = srcCnt[i];
    const uint32_t* src = srcList + srcLo[i];
    uint32_t* dst = sets + lv.setBase[i];
    const uint32_t words = uint32_t((uint64_t(cnt) * bits + 31) / 32);
    if (threadIdx.x == 0) { dst[0] = cnt; dst[1 + words] = 0u; }
    for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) {
        const uint64_t bitLo = uint64_t(w) * 32;
        uint32_t t = uint32_t(bitLo / bits);
        uint32_t acc = 0;
        for (; t < cnt; t++) {
            const uint64_t start = uint64_t(t) * bits;
            if (start >= bitLo + 32) break;
            const uint32_t index = src[t];
            // element occupies bits [start, start + bits) of the stream, MSB first
            const int shift = int(bitLo + 32) - int(start + bits);   // > 0: shift left inside the word
            acc |= shift >= 0 ? (index << shift) : (index >> (-shift));
        }
        dst[1 + w] = acc;
    }
}

// 8 masks of an inner node at depth >= bitEnc: child c, byte b = member bits of merged entries 8b .. 8b+7, MSB first
__global__ void __launch_bounds__(128)
emitMasksKernel(LayoutView lv, uint32_t n, const uint32_t* __restrict__ mergedLo, const uint8_t* __restrict__ mergedMember,
                uint8_t* __restrict__ masks) {
    const uint32_t i = blockIdx.x;
    if (i >= n || lv.nodeBase[i] == kNone || lv.childOf[i] == kNone) return;
    const uint32_t cnt = lv.mergedCnt[i], bytes = (cnt + 7u) / 8u;
    const uint8_t* mem = mergedMember + mergedLo[i];
    uint8_t* dst = masks + lv.ownMaskBase[i];
    for (uint32_t b = threadIdx.x; b < bytes; b += blockDim.x) {
        uint32_t m[8];
#pragma unroll
        for (int q = 0; q < 8; q++) m[q] = (8 * b + q < cnt) ? mem[8 * b + q] : 0u;
#pragma unroll
        for (uint32_t c = 0; c < 8; c++) {
            uint32_t v = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) v |= ((m[q] >> c) & 1u) << (7 - q);
            dst[size_t(c) * bytes + b] = uint8_t(v);
        }
    }
}

__global__ void fillU64(unsigned long long* p, unsigned long long v, uint64_t n) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- host orchestration ------------------------------------------------------------------------------------
struct Level {
    uint32_t count = 0, depth = 0;
    DevBuf<float4> centerHalf;
    DevBuf<uint32_t> info, coord, parent, parentLo, parentCnt, listLo, listCnt, childOf, list, pos;
    DevBuf<uint64_t> pairOff;
    DevBuf<uint8_t> flags;
    uint64_t numPairs = 0;
    uint32_t listTotal = 0;
    // merge levels
    DevBuf<uint8_t> member, mergedMember;
    DevBuf<uint32_t> mergedLo, mergedCnt, mergedList;
    uint32_t mergedTotal = 0;
    // layout
    DevBuf<uint32_t> nodeCnt, setWords, maskBytes, slot, nodeBase, setBase, maskBase, ownMaskBase;

    void allocNodes(uint32_t n, uint32_t d) {
        count = n; depth = d;
        centerHalf.alloc(n); info.alloc(size_t(n) * 8); coord.alloc(n); parent.alloc(n); parentLo.alloc(n); parentCnt.alloc(n);
        listLo.alloc(n); listCnt.alloc(n); childOf.alloc(n); pairOff.alloc(size_t(n) + 1);
    }
    LevelView view() const {
        return LevelView{count, depth, centerHalf.p, info.p, parentLo.p, parentCnt.p, pairOff.p, listLo.p, listCnt.p};
    }
    void allocLayout() {
        nodeCnt.alloc(count); setWords.alloc(count); maskBytes.alloc(count); slot.alloc(count); nodeBase.alloc(count);
        setBase.alloc(count); maskBase.alloc(count); ownMaskBase.alloc(count);
    }
    LayoutView layout() const {
        return LayoutView{childOf.p, listCnt.p, mergedCnt.p, nodeCnt.p, setWords.p, maskBytes.p, slot.p,
                          nodeBase.p, setBase.p, maskBase.p, ownMaskBase.p};
    }
};

__global__ void zeroUnownedKernel(const uint8_t* owned, uint32_t* parentCnt, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !owned[i]) parentCnt[i] = 0u;
}

struct ExactBuildState : BuildState {
    std::vector<std::unique_ptr<Level>> levels;
    uint32_t maxDepth = 0, startDepth = 0, minTris = 0, bitEnc = 0, bits = 0;
    DevBuf<uint32_t> scalars;   // [0] maxTrianglesInLeafs, [1] maxTrianglesEncodedInLeafs
    uint32_t numStreams() const override { return 3; }

    void buildLevels(sdfb200_sdf& out, const std::shared_ptr<PreparedMesh>& meshPtr, uint32_t numThreads, uint32_t rank, uint32_t world) {
        sdfb200_build_stats& st = out.stats;
        const PreparedMesh& mesh = *meshPtr;
        if (!mesh.hasExactParts) throw Error(SDFB200_ERR_INVALID, "ExactOctreeSdf needs a mesh prepared with its frames (SDFB200_MESH_EXACT)");
        // ---- set-up of the reference (TriangleData, degenerate filter): done by mesh_device.cu ---------------------------
        st.triangle_data_ms = mesh.triangleDataMs;
        st.upload_ms = mesh.uploadMs;
        const uint32_t nT = mesh.nTris;
        out.mesh = meshPtr;          // the structure's queries read the mesh's TriangleData and frames
        out.numTris = nT;
        out.qTris = mesh.dev.tris.p;
        out.qFrames = mesh.frames.p;
        out.bitsPerIndex = uint32_t(int32_t(std::ceil(std::log2(float(nT)))));   // ExactOctreeSdfDepthFirst.h:61
        bits = out.bitsPerIndex;
        if (bits == 0 || bits > 31) throw Error(SDFB200_ERR_INVALID, "ExactOctreeSdf needs between 2 and 2^31 triangles");
        const float4* dFramesP = mesh.frames.p;
        const uint32_t* dAllP = mesh.valid.p;
        const uint32_t numAll = mesh.numValid;
        const DeviceMesh dmesh{mesh.dev.verts.p, mesh.dev.idx.p, nullptr, nullptr, nT};

        NvtxRange nvtx("sdfb200:exact:levels");
        auto t0 = std::chrono::steady_clock::now();
        const uint32_t d0 = std::min(startDepth, 1u);
        const f3 boxMin = mk3(out.boxMin[0], out.boxMin[1], out.boxMin[2]);
        const float boxSize = out.boxMax[0] - out.boxMin[0];
        levels.resize(maxDepth + 1);
        ScannerT<uint32_t, uint32_t> scan32;
        ScannerT<uint32_t, uint64_t> scan64;
        FlagScanner scanFlags;
        scalars.alloc(2);
        SDFB_CUDA(cudaMemset(scalars.p, 0, 8));
        DevBuf<unsigned long long> best;
        DevBuf<uint32_t> subdivide, chunks, childIdx, chunkOff, midInfo;
        DevBuf<float> region;
        DevBuf<uint8_t> dOwned;

        auto runSample = [&](auto ptsTag, const Level& L, const uint32_t* list, const uint32_t* lo, const uint32_t* cnt,
                             const uint32_t* chOff, uint32_t nChunks, uint32_t* outInfo) {
            constexpr int kPts = decltype(ptsTag)::value;
            const uint64_t nKeys = uint64_t(L.count) * kPts;
            if (best.n < nKeys) best.alloc(nKeys);
            fillU64<<<divUp(nKeys, 256), 256>>>(best.p, kNoKey, nKeys);
            if (nChunks) sampleKernel<kPts><<<divUp(nChunks, kSampleThreads / 32), kSampleThreads>>>(dFramesP, L.centerHalf.p, list, lo, cnt, chOff, L.count, nChunks, best.p);
            resolveKernel<<<divUp(nKeys, 256), 256>>>(best.p, list, lo, kPts, outInfo, nKeys);
            st.kernel_launches += 3;
            SDFB_CUDA(cudaGetLastError());
        };

        {   // seeds at depth d0; corner info = nearest of ALL valid triangles (ExactOctreeSdfDepthFirst.h:113-150)
            const float h0 = float(0.5f * boxSize * std::pow(0.5f, d0));
            const f3 c0 = boxMin + mk3(h0, h0, h0);
            const uint32_t per = 1u << d0;
            std::vector<float4> ch;
            std::vector<uint32_t> coord;
            for (uint32_t k = 0; k < per; k++)
                for (uint32_t j = 0; j < per; j++)
                    for (uint32_t i = 0; i < per; i++) {
                        const f3 c = c0 + (mk3(float(i), float(j), float(k)) * 2.0f) * h0;
                        ch.push_back(make_float4(c.x, c.y, c.z, h0));
                        coord.push_back(i | (j << 10) | (k << 20));
                    }
            levels[d0].reset(new Level());
            Level& L = *levels[d0];
            const uint32_t n = uint32_t(ch.size());
            L.allocNodes(n, d0);
            L.centerHalf.upload(ch.data(), n);
            L.coord.upload(coord.data(), n);
            std::vector<uint32_t> zero(n, 0u), cntAll(n, numAll), chOff(n + 1);
            const uint32_t perNode = divUp(numAll, kChunk);
            for (uint32_t i = 0; i <= n; i++) chOff[i] = i * perNode;
            L.parent.upload(zero.data(), n);
            L.parentLo.upload(zero.data(), n);
            L.parentCnt.upload(cntAll.data(), n);
            chunkOff.alloc(n + 1);
            chunkOff.upload(chOff.data(), n + 1);
            runSample(std::integral_constant<int, 8>(), L, dAllP, L.parentLo.p, L.parentCnt.p, chunkOff.p, n * perNode, L.info.p);
        }

        static const bool timing = std::getenv("SDFB200_TIMING") != nullptr;
        auto tLevel = std::chrono::steady_clock::now();
        auto levelDone = [&](uint32_t d, const Level& L) {   // SDFB200_TIMING: synchronised per-depth times (diagnostic runs only)
            if (!timing) return;
            SDFB_CUDA(cudaDeviceSynchronize());
            std::fprintf(stderr, "[sdfb200] exact depth %u: %9u nodes %12llu pairs %10u kept %8.2f ms\n", d, L.count,
                         (unsigned long long)L.numPairs, L.listTotal, msSince(tLevel));
            tLevel = std::chrono::steady_clock::now();
        };
        for (uint32_t d = d0; d <= maxDepth; d++) {
            Level& L = *levels[d];
            const uint32_t* parentList = d == d0 ? dAllP : levels[d - 1]->list.p;
            if (d == startDepth) {
                makePlan(out, L, numThreads, rank, world);
                if (world > 1) {   // roots of other ranks: empty list -> terminal, nothing below them is built here
                    dOwned.alloc(L.count);
                    dOwned.upload(out.plan.owned.data(), L.count);
                    zeroUnownedKernel<<<divUp(L.count, 256), 256>>>(dOwned.p, L.parentCnt.p, L.count);
                }
            }
            if (L.count == 0) {
                if (d < maxDepth) { levels[d + 1].reset(new Level()); levels[d + 1]->depth = d + 1; }
                L.list.alloc(8);
                continue;
            }
            // HOT LOOP A
            L.numPairs = scan64.run(L.parentCnt.p, L.pairOff.p, L.count, true);
            // pair indices are 64-bit everywhere; list positions (kept pairs) are 32-bit and are checked after the filter
            if (L.numPairs >= (uint64_t(1) << 38)) throw Error(SDFB200_ERR_INVALID, "more than 2^38 (node, triangle) pairs on one octree level");
            if (region.n < size_t(L.count) * 72) region.alloc(size_t(L.count) * 72);
            regionKernel<<<divUp(L.count, 4), 256>>>(dFramesP, L.view(), region.p);
            L.flags.alloc(L.numPairs + 1);
            L.pos.alloc(L.numPairs / 16 + 2);   // one position per group of 16 pairs
            if (L.numPairs) filterRefillKernel<<<divUp(L.numPairs, 8 * kFilterChunk), 256>>>(dmesh, L.view(), parentList, region.p, L.flags.p, L.numPairs);
            SDFB_CUDA(cudaGetLastError());
            if (L.numPairs >= (uint64_t(1) << 32)) {   // the 32-bit scan below would wrap silently: count the kept pairs in 64 bits first
                const uint64_t kept = countFlags64(L.flags.p, L.numPairs);
                if (kept >= (1ull << 32)) throw Error(SDFB200_ERR_INVALID, "more than 2^32 triangle-list entries on one octree level");
            }
            L.listTotal = L.numPairs ? scanFlags.runGroups(L.flags.p, L.pos.p, L.numPairs) : 0u;
            L.list.alloc(size_t(L.listTotal) + 8);   // + 8: the TMA window of the sample kernel may read past the end
            SDFB_CUDA(cudaMemsetAsync(L.list.p + L.listTotal, 0, 8 * sizeof(uint32_t)));
            if (L.numPairs) compactKernel<<<divUp(L.numPairs, 256 * 16), 256>>>(L.view(), parentList, L.flags.p, L.pos.p, L.list.p, L.numPairs);
            listRangeKernel<<<divUp(L.count, 256), 256>>>(L.pairOff.p, L.flags.p, L.pos.p, L.numPairs, L.listTotal, L.listLo.p, L.listCnt.p, L.count);
            st.kernel_launches += 10;
            st.nodes_processed += L.count;
            st.samples_evaluated += L.numPairs;   // Frank-Wolfe runs
            // terminal rule
            subdivide.alloc(L.count); chunks.alloc(L.count); childIdx.alloc(L.count); chunkOff.alloc(size_t(L.count) + 1);
            decideKernel<<<divUp(L.count, 256), 256>>>(L.listCnt.p, L.count, d, startDepth, maxDepth, minTris, subdivide.p, chunks.p, scalars.p);
            if (d == maxDepth) { SDFB_CUDA(cudaMemsetAsync(L.childOf.p, 0xFF, size_t(L.count) * 4)); levelDone(d, L); break; }
            const uint32_t nSub = scan32.run(subdivide.p, childIdx.p, L.count);
            const uint32_t nChunks = scan32.run(chunks.p, chunkOff.p, L.count, true);
            // HOT LOOP B
            midInfo.alloc(size_t(L.count) * 19);
            runSample(std::integral_constant<int, 19>(), L, L.list.p, L.listLo.p, L.listCnt.p, chunkOff.p, nChunks, midInfo.p);
            st.leaves += uint64_t(L.listTotal);   // list entries kept at this depth (19 distance evaluations each where subdividing)
            levels[d + 1].reset(new Level());
            Level& N = *levels[d + 1];
            N.allocNodes(nSub * 8, d + 1);
            childrenKernel<<<divUp(L.count, 4), 256>>>(L.view(), L.coord.p, subdivide.p, childIdx.p, midInfo.p, L.childOf.p, N.centerHalf.p,
                                                      N.info.p, N.coord.p, N.parent.p, N.parentLo.p, N.parentCnt.p);
            st.kernel_launches += 8;
            SDFB_CUDA(cudaGetLastError());
            if (d + 1 < maxDepth) { L.flags.release(); L.pos.release(); }   // only the flags of levels maxDepth-1 and maxDepth feed the merge
            levelDone(d, L);
        }
        SDFB_CUDA(cudaDeviceSynchronize());
        st.levels_ms = msSince(t0);

        // ---- post-order merge: levels maxDepth-1, then maxDepth-2 ------------------------------------------------
        nvtx.next("sdfb200:exact:merge_and_layout");
        t0 = std::chrono::steady_clock::now();
        DevBuf<uint8_t> keep;
        DevBuf<uint32_t> mpos;
        for (uint32_t d = maxDepth - 1; d + 1 > bitEnc; d--) {
            Level& L = *levels[d];
            Level& C = *levels[d + 1];
            L.mergedLo.alloc(L.count); L.mergedCnt.alloc(L.count);
            if (L.count == 0) { if (d == 0) break; continue; }
            L.member.alloc(size_t(L.listTotal) + 1); keep.alloc(size_t(L.listTotal) + 1); mpos.alloc(size_t(L.listTotal) + 1);
            if (L.listTotal) {
                MergeChildView cv{C.pairOff.p, C.flags.p, C.pos.p, C.numPairs, C.childOf.p, d + 1 == maxDepth ? nullptr : C.member.p};
                memberKernel<<<divUp(L.listTotal, 256), 256>>>(L.listLo.p, L.childOf.p, L.count, L.listTotal, cv, L.member.p, keep.p);
                L.mergedTotal = scanFlags.run(keep.p, mpos.p, L.listTotal);
            }
            L.mergedList.alloc(size_t(L.mergedTotal) + 1); L.mergedMember.alloc(size_t(L.mergedTotal) + 1);
            if (L.listTotal) mergeCompactKernel<<<divUp(L.listTotal, 256), 256>>>(L.list.p, L.member.p, mpos.p, L.listTotal, L.mergedList.p, L.mergedMember.p);
            mergedRangeKernel<<<divUp(L.count, 256), 256>>>(L.listLo.p, L.listCnt.p, mpos.p, L.listTotal, L.mergedTotal, L.mergedLo.p, L.mergedCnt.p, L.count);
            st.kernel_launches += 6;
            SDFB_CUDA(cudaGetLastError());
            if (d == 0) break;
        }

        // ---- layout: subtree sizes bottom-up ------------------------------------------------------------------------
        for (int d = int(maxDepth); d >= int(startDepth); d--) {
            Level& L = *levels[size_t(d)];
            L.allocLayout();
            if (!L.count) continue;
            LayoutView next = d < int(maxDepth) ? levels[size_t(d) + 1]->layout() : LayoutView{};
            sizesKernel<<<divUp(L.count, 256), 256>>>(L.layout(), next, L.count, uint32_t(d), maxDepth, bitEnc, bits, scalars.p + 1);
            st.kernel_launches++;
        }
        Level& R = *levels[startDepth];
        const uint32_t G3 = out.plan.G3;
        std::vector<uint32_t> rootNodes(G3), rootSets(G3), rootMasks(G3);
        R.nodeCnt.download(rootNodes.data(), G3);
        R.setWords.download(rootSets.data(), G3);
        R.maskBytes.download(rootMasks.data(), G3);
        SDFB_CUDA(cudaDeviceSynchronize());
        out.shardSizes.assign(size_t(G3) * 3, 0u);
        for (uint32_t r = 0; r < G3; r++) {
            if (!out.plan.owned[r]) continue;
            const size_t s = size_t(out.plan.rootSlot[r]) * 3;
            out.shardSizes[s] = rootNodes[r];
            out.shardSizes[s + 1] = rootSets[r];
            out.shardSizes[s + 2] = rootMasks[r];
        }
        st.layout_ms = msSince(t0);
    }

    void makePlan(sdfb200_sdf& out, Level& R, uint32_t numThreads, uint32_t rank, uint32_t world) {
        const uint32_t G = uint32_t(out.startGridSize), G3 = G * G * G;
        if (R.count != G3) throw Error(SDFB200_ERR_INVALID, "internal: start level is not a full grid");
        std::vector<float4> rootCH(G3);
        std::vector<uint32_t> rootCoord(G3);
        std::vector<uint32_t> weight(G3);   // work estimate of a start voxel: the triangles its parent kept (its own candidates)
        R.centerHalf.download(rootCH.data(), G3);
        R.coord.download(rootCoord.data(), G3);
        R.parentCnt.download(weight.data(), G3);
        SDFB_CUDA(cudaDeviceSynchronize());
        out.plan = makeRootPlan(rootCH.data(), rootCoord.data(), G, startDepth, out.boxMin, out.cellSize, numThreads, rank, world,
                                world > 1 ? weight.data() : nullptr);
    }

    void finish(sdfb200_sdf& out, const uint32_t* allSizesBySlot) override {
        sdfb200_build_stats& st = out.stats;
        NvtxRange nvtx("sdfb200:exact:emit");
        auto t0 = std::chrono::steady_clock::now();
        const RootPlan& plan = out.plan;
        const uint32_t G3 = plan.G3;
        Level& R = *levels[startDepth];
        out.streams.assign(3, ShardStream());
        const uint32_t elem[3] = {8, 4, 1};
        for (int k = 0; k < 3; k++) { out.streams[k].elemBytes = elem[k]; out.streams[k].rootBase.resize(G3); out.streams[k].rootSize.resize(G3); }
        std::vector<uint32_t> nodeBase(G3), setBase(G3), maskBase(G3);
        uint64_t run[3] = {G3, 0, 0};
        for (uint32_t i = 0; i < G3; i++) {
            const uint32_t r = plan.order[i];
            for (int k = 0; k < 3; k++) {
                out.streams[k].rootBase[r] = run[k];
                out.streams[k].rootSize[r] = allSizesBySlot[size_t(plan.rootSlot[r]) * 3 + k];
            }
            nodeBase[r] = plan.owned[r] ? uint32_t(run[0]) : kNone;
            setBase[r] = uint32_t(run[1]);
            maskBase[r] = uint32_t(run[2]);
            for (int k = 0; k < 3; k++) run[k] += out.streams[k].rootSize[r];
        }
        const uint64_t rn = run[0], rs = run[1], rm = run[2];
        if (rn > uint64_t(kExactIndexMask) || rs > 0xFFFFFFFFull || rm > 0xFFFFFFFFull)
            throw Error(SDFB200_ERR_INVALID, "ExactOctreeSdf exceeds the 32-bit index space of its arrays");
        R.slot.upload(plan.rootSlot.data(), G3);
        R.nodeBase.upload(nodeBase.data(), G3);
        R.setBase.upload(setBase.data(), G3);
        R.maskBase.upload(maskBase.data(), G3);
        for (uint32_t d = startDepth; d < maxDepth; d++) {
            Level& L = *levels[d];
            Level& N = *levels[d + 1];
            if (!L.count) continue;
            offsetsKernel<<<divUp(L.count, 256), 256>>>(L.layout(), N.layout(), L.count);
            st.kernel_launches++;
        }
        // ---- emit -----------------------------------------------------------------------------------------------------
        out.dOctree.alloc(size_t(rn) * 2);
        out.dSets.alloc(size_t(rs) + 1);
        out.dMasks.alloc(size_t(rm) + 8);
        out.streams[0].dBase = reinterpret_cast<uint8_t*>(out.dOctree.p);
        out.streams[1].dBase = reinterpret_cast<uint8_t*>(out.dSets.p);
        out.streams[2].dBase = reinterpret_cast<uint8_t*>(out.dMasks.p);
        SDFB_CUDA(cudaMemsetAsync(out.dOctree.p, 0, size_t(rn) * 8));
        SDFB_CUDA(cudaMemsetAsync(out.dSets.p, 0, (size_t(rs) + 1) * 4));
        SDFB_CUDA(cudaMemsetAsync(out.dMasks.p, 0, size_t(rm) + 8));
        for (uint32_t d = startDepth; d <= maxDepth; d++) {
            Level& L = *levels[d];
            if (!L.count) continue;
            emitNodesKernel<<<divUp(L.count, 256), 256>>>(L.layout(), L.count, d, bitEnc, out.dOctree.p);
            if (d <= bitEnc) {
                emitSetsKernel<<<L.count, 128>>>(L.layout(), L.count, d, bitEnc, bits, L.list.p, L.listLo.p, L.listCnt.p, false, out.dSets.p);
                if (d == bitEnc)
                    emitSetsKernel<<<L.count, 128>>>(L.layout(), L.count, d, bitEnc, bits, L.mergedList.p, L.mergedLo.p, L.mergedCnt.p, true, out.dSets.p);
            }
            if (d >= bitEnc && d < maxDepth) emitMasksKernel<<<L.count, 128>>>(L.layout(), L.count, L.mergedLo.p, L.mergedMember.p, out.dMasks.p);
            st.kernel_launches += 4;
            SDFB_CUDA(cudaGetLastError());
        }
        scalars.download(out.shardScalars, 2);
        SDFB_CUDA(cudaDeviceSynchronize());
        st.layout_ms += msSince(t0);
        levels.clear();   // release the builder's working set before the query-side pool is allocated
        out.nOctree = rn * 2; out.nSets = rs; out.nMasks = rm;
        out.hostMirror = false;
        if (plan.world == 1) {
            out.maxTrisInLeafs = out.shardScalars[0];
            out.maxTrisEncoded = out.shardScalars[1];
            nvtx.next("sdfb200:exact:download");
            t0 = std::chrono::steady_clock::now();
            ensureHostMirror(out);
            st.download_ms = msSince(t0);
            prepareExactQuery(out);
            out.isShard = false;
        }
        SDFB_CUDA(cudaDeviceSynchronize());
    }
};

}  // namespace

void cubifyBox(sdfb200_sdf& s, const float* box6, uint32_t startDepth);   // octree_build.cu

void buildExactOnDevice(sdfb200_sdf& out, const std::shared_ptr<PreparedMesh>& mesh, const float* box6, uint32_t maxDepth, uint32_t startDepth,
                        uint32_t minTris, uint32_t numThreads, uint32_t rank, uint32_t world) {
    const auto tStart = std::chrono::steady_clock::now();
    out.stats = sdfb200_build_stats{};
    if (maxDepth > 10) throw Error(SDFB200_ERR_INVALID, "octree depth > 10 is not supported (node coordinates are packed in 3x10 bits)");
    if (maxDepth < startDepth + 2)
        throw Error(SDFB200_ERR_INVALID, "ExactOctreeSdf needs maxDepth >= startDepth + 2 (the reference dereferences a null node otherwise)");
    out.format = SDFB200_FORMAT_EXACT_OCTREE;
    out.maxDepth = maxDepth;
    out.startDepth = startDepth;
    out.minTrisInLeafs = minTris;
    out.bitEncodingStartDepth = maxDepth - 2;
    out.slotWords = 2;
    cubifyBox(out, box6, startDepth);
    SDFB_CUDA(cudaGetDevice(&out.device));
    DeviceCacheSettle settle(out.device);
    std::unique_ptr<ExactBuildState> state(new ExactBuildState());
    state->maxDepth = maxDepth; state->startDepth = startDepth; state->minTris = minTris; state->bitEnc = maxDepth - 2;
    out.isShard = true;
    state->buildLevels(out, mesh, numThreads, rank, world);
    if (world == 1) state->finish(out, out.shardSizes.data());
    else out.build = std::move(state);
    out.stats.total_ms = msSince(tStart);
}

}  // namespace sdfb200
