// Device code shared by the OctreeSdf builders (octree_build.cu: NO_CONTINUITY, octree_cont.cu: CONTINUITY):
// float64 sphere-BVH nearest-triangle descent, point sampling, the 64x64 Hermite map and the exact-order
// polynomial evaluation. Everything has internal linkage: each translation unit gets its own copy of the
// __constant__ tables (no relocatable device code).
#pragma once
#include <cfloat>
#include <cmath>
#include <algorithm>
#include <cstdlib>

#include "device_utils.cuh"
#include "sdf_internal.h"

namespace sdfb200 {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr uint32_t kNoChild = 0xFFFFFFFFu;

// lattice index L = x + 3y + 9z of the 19 mid-points, in the reference's sample order
__constant__ int cSampleLattice[19] = {1, 3, 4, 5, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 19, 21, 22, 23, 25};

// Non-zero entries of the 64x64 Hermite map W = H(x)H(x)H, row-major, columns ascending.
struct alignas(16) HermiteTable {   // sizeof is a multiple of 16, so the word-wise copy below is exact
    uint16_t rowStart[65];
    uint8_t col[1000];
    int8_t weight[1000];
    uint8_t order[64];   // rows sorted by descending number of terms (load balance across lanes)
};
__constant__ HermiteTable cHermite;

HermiteTable makeHermiteTable() {
    static const int H[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {-3, -2, 3, -1}, {2, 1, -2, 1}};
    static const int slot[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
    HermiteTable t;
    int n = 0;
    for (int row = 0; row < 64; row++) {
        t.rowStart[row] = uint16_t(n);
        const int i = row & 3, j = (row >> 2) & 3, k = row >> 4;
        for (int c = 0; c < 8; c++)
            for (int s = 0; s < 8; s++) {
                const int w = H[i][2 * (c & 1) + slot[s][0]] * H[j][2 * ((c >> 1) & 1) + slot[s][1]] *
                              H[k][2 * ((c >> 2) & 1) + slot[s][2]];
                if (w != 0) { t.col[n] = uint8_t(c * 8 + s); t.weight[n] = int8_t(w); n++; }
            }
    }
    t.rowStart[64] = uint16_t(n);
    int o = 0;
    for (int want : {64, 16, 4, 1})
        for (int row = 0; row < 64; row++)
            if (t.rowStart[row + 1] - t.rowStart[row] == want) t.order[o++] = uint8_t(row);
    return t;
}

// ---- device helpers ------------------------------------------------------------------------------

// Squared point-triangle distance in float64 (the vendored BVH's point_triangle_sq_unsigned, Eberly's
// region method; TriangleMeshDistance.h:542-798). Same operations per region as eberlySqDist (tri_math.cuh),
// restructured so that a warp whose lanes fall into different regions still shares the expensive part: the
// region only selects a mode and one (numerator, denominator) pair, then ONE division and ONE quadratic form
// are evaluated by all lanes together.
__device__ __forceinline__ double eberlySqDistConverged(d3 p, d3 v0, d3 v1, d3 v2) {
    const d3 diff = v0 - p, e0 = v1 - v0, e1 = v2 - v0;
    const double a00 = ddot(e0, e0), a01 = ddot(e0, e1), a11 = ddot(e1, e1);
    const double b0 = ddot(diff, e0), b1 = ddot(diff, e1), c = ddot(diff, diff);
    const double det = fabs(a00 * a11 - a01 * a01);
    const double s = a01 * b1 - a11 * b0;
    const double t = a01 * b0 - a00 * b1;
    enum { kC, kV1, kV2, kE0, kE1, kQ0, kQs, kQt };
    int mode;
    double num = 1.0, den = 1.0;
    if (s + t <= det) {
        if (s < 0) {
            if (t < 0 && b0 < 0) mode = (-b0 >= a00) ? kV1 : kE0;
            else mode = (b1 >= 0) ? kC : ((-b1 >= a11) ? kV2 : kE1);
        } else if (t < 0) {
            mode = (b0 >= 0) ? kC : ((-b0 >= a00) ? kV1 : kE0);
        } else {
            mode = kQ0;
            den = det;
        }
    } else if (s < 0) {   // region 2
        const double tmp0 = a01 + b0, tmp1 = a11 + b1;
        if (tmp1 > tmp0) {
            num = tmp1 - tmp0; den = a00 - 2 * a01 + a11;
            if (num >= den) { mode = kV1; num = 1.0; den = 1.0; } else mode = kQs;
        } else mode = (tmp1 <= 0) ? kV2 : ((b1 >= 0) ? kC : kE1);
    } else if (t < 0) {   // region 6
        const double tmp0 = a01 + b1, tmp1 = a00 + b0;
        if (tmp1 > tmp0) {
            num = tmp1 - tmp0; den = a00 - 2 * a01 + a11;
            if (num >= den) { mode = kV2; num = 1.0; den = 1.0; } else mode = kQt;
        } else mode = (tmp1 <= 0) ? kV1 : ((b0 >= 0) ? kC : kE0);
    } else {              // region 1
        num = a11 + b1 - a01 - b0;
        if (num <= 0) { mode = kV2; num = 1.0; }
        else {
            den = a00 - 2 * a01 + a11;
            if (num >= den) { mode = kV1; num = 1.0; den = 1.0; } else mode = kQs;
        }
    }
    if (mode == kE0) { num = -b0; den = a00; }
    if (mode == kE1) { num = -b1; den = a11; }
    const double q = num / den;
    double ss, tt;
    if (mode == kQ0) { ss = s * q; tt = t * q; }
    else if (mode == kQt) { tt = q; ss = 1 - tt; }
    else { ss = q; tt = 1 - ss; }
    const double quad = ss * (a00 * ss + a01 * tt + 2 * b0) + tt * (a01 * ss + a11 * tt + 2 * b1) + c;
    double d2;
    switch (mode) {
        case kC: d2 = c; break;
        case kV1: d2 = a00 + 2 * b0 + c; break;
        case kV2: d2 = a11 + 2 * b1 + c; break;
        case kE0: d2 = b0 * q + c; break;
        case kE1: d2 = b1 * q + c; break;
        default: d2 = quad; break;
    }
    return d2 < 0 ? 0 : d2;
}

// Nearest triangle id with the reference's traversal order (TriangleMeshDistance.h:492-540): near child
// first, the far child is re-tested against the running best when the near subtree is done, leaves replace
// the best only on strict '<' against the re-squared running distance. Links < 0 are ~triangleId, so no
// leaf node is ever loaded. Per-lane semantics are identical in all variants below; they differ only in how
// the lanes of a warp are kept together (measured on B200, see profiles/).
// The traversal stack lives in SHARED memory, one column per thread: entry i of thread t is at [i * blockDim + t],
// so a warp access is conflict-free whatever the per-lane stack pointers are. (In local memory the same accesses
// were uncoalesced — 1.9 useful bytes per 32-byte sector — and made up 60 % of the L1 wavefronts of the sampling
// kernel: profiles/r1_sample_lattice_*.) Depth = height of the median-split BVH, known on the host. The full height
// costs occupancy (20 entries x 12 B x 128 threads on the C2 mesh = 28 resident warps per SM), but keeping only 8
// entries in shared memory and spilling the rest to a local array was measured slower (C2 levels 112.5 against 83 ms):
// the dynamically indexed spill arrays put the whole cursor back into local memory.
struct BvhStack {
    double* dist;   // [depth][blockDim]
    int* node;      // [depth][blockDim]
    int stride;
};
constexpr int kBvhThreads = 128;   // CTA size of every kernel that traverses the BVH
inline size_t bvhStackBytes(const DeviceMesh& m, int threads = kBvhThreads) { return size_t(m.stackDepth) * threads * 12; }
__device__ __forceinline__ BvhStack bvhStackOfThread(const DeviceMesh& m) {
    extern __shared__ double bvhStackSmem[];
    BvhStack st;
    st.stride = int(blockDim.x);
    st.dist = bvhStackSmem + threadIdx.x;
    st.node = reinterpret_cast<int*>(bvhStackSmem + size_t(m.stackDepth) * blockDim.x) + threadIdx.x;
    return st;
}

struct BvhCursor {
    d3 p;
    double best;
    int bestTri, sp, cur;
    bool active;
};

__device__ __forceinline__ void bvhPop(BvhCursor& c, const BvhStack& st) {
    c.active = false;
    while (c.sp > 0) {
        c.sp--;
        if (st.dist[c.sp * st.stride] < c.best) { c.cur = st.node[c.sp * st.stride]; c.active = true; break; }
    }
}

__device__ __forceinline__ void bvhInnerStep(const DeviceMesh& m, BvhCursor& c, const BvhStack& st) {
    const BvhNode nd = m.bvh[c.cur];
    const d3 dl3 = c.p - mkd(nd.lc[0], nd.lc[1], nd.lc[2]);
    const d3 dr3 = c.p - mkd(nd.rc[0], nd.rc[1], nd.rc[2]);
    const double dl = sqrt(ddot(dl3, dl3)) - nd.lr;
    const double dr = sqrt(ddot(dr3, dr3)) - nd.rr;
    const bool leftFirst = dl < dr;
    const int first = leftFirst ? nd.left : nd.right, second = leftFirst ? nd.right : nd.left;
    const double dFirst = leftFirst ? dl : dr, dSecond = leftFirst ? dr : dl;
    // the far child is re-tested against the running best when it is popped; the best only shrinks, so a far
    // child that already fails now can never pass later and is not pushed at all
    if (dSecond < c.best) {
        st.node[c.sp * st.stride] = second;
        st.dist[c.sp * st.stride] = dSecond;
        c.sp++;
    }
    if (dFirst < c.best) c.cur = first;
    else bvhPop(c, st);
}

__device__ __forceinline__ void bvhLeafStep(const DeviceMesh& m, BvhCursor& c, const BvhStack& st) {
    const int t = ~c.cur;
    const float4 a = m.triVerts[3 * size_t(t)], b = m.triVerts[3 * size_t(t) + 1], v = m.triVerts[3 * size_t(t) + 2];
    const double d2 = eberlySqDistConverged(c.p, mkd(double(a.x), double(a.y), double(a.z)), mkd(double(b.x), double(b.y), double(b.z)),
                                            mkd(double(v.x), double(v.y), double(v.z)));
    if (d2 < c.best * c.best) { c.best = sqrt(d2); c.bestTri = t; }
    bvhPop(c, st);
}

// One node per iteration and lane. Requesting both children (prefetch.global.L1) as soon as their links are known
// was measured too: 86.1 against 83.7 ms. Two warp-synchronous schedules were measured on the C2 build and dropped:
// "while-while" (walk inner nodes until a leaf is held, then evaluate; 354 ms against 158 ms) and a ballot-driven
// schedule where the whole warp does either an inner or a leaf step per iteration (213 ms): lanes are bound by
// their own dependent-load chains, and waiting for the slowest lane costs more than the divergence.
__device__ uint32_t bvhNearest(const DeviceMesh& m, f3 pf) {
    const BvhStack st = bvhStackOfThread(m);
    BvhCursor c;
    c.p = mkd(double(pf.x), double(pf.y), double(pf.z));
    c.best = DBL_MAX;
    c.bestTri = -1;
    c.sp = 0;
    c.cur = m.rootLink;
    c.active = true;
    while (c.active) {
        if (c.cur >= 0) bvhInnerStep(m, c, st);
        else bvhLeafStep(m, c, st);
    }
    return uint32_t(c.bestTri);
}

// TriCubicInterpolation::calculatePointValues: (signed distance, unit gradient) of the nearest triangle
__device__ __forceinline__ float4 samplePoint(const DeviceMesh& m, f3 p) {
    const uint32_t t = bvhNearest(m, p);
    f3 g;
    const float4 a = m.triVerts[3 * size_t(t)], b = m.triVerts[3 * size_t(t) + 1], c = m.triVerts[3 * size_t(t) + 2];
    const float d = signedDistGradMesh(p, m.tris[t], mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), g);
    return make_float4(d, g.x, g.y, g.z);
}

// The 19 mid-point samples of every node of a level, one thread per (node, sample): all 32 lanes of a warp
// traverse the BVH (a warp-per-node layout leaves 13 of 32 lanes idle during the dominant phase).
// out[(node * 19 + s) * stride] = (d, gx, gy, gz); with stride 2 the mixed-derivative half is zeroed.
__global__ void __launch_bounds__(kBvhThreads) sampleLatticeKernel(DeviceMesh mesh, const float4* centerHalf, uint32_t count, float4* out, int stride) {
    const uint64_t t = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= uint64_t(count) * 19) return;
    const uint32_t node = uint32_t(t / 19), s = uint32_t(t % 19);
    const float4 ch = centerHalf[node];
    const int L = cSampleLattice[s];
    const f3 rel = mk3(float(L % 3 - 1), float((L / 3) % 3 - 1), float(L / 9 - 1));
    out[t * stride] = samplePoint(mesh, mk3(ch.x, ch.y, ch.z) + rel * ch.w);
    if (stride == 2) out[t * 2 + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- sampling with exact de-duplication ---------------------------------------------------------------------
// Neighbouring nodes share lattice points, and the nearest-triangle query is a pure function of the sample
// position, so positions that are BIT-IDENTICAL need one traversal only. (The positions of one lattice point
// computed from different nodes differ by an ulp about 40 % of the time, because each node's centre went through its
// own chain of float additions; those stay separate queries, which keeps every value exactly what the node would
// have computed on its own.) On a uniformly refined level 59 % of the 19 N samples are distinct.
// Open-addressing table of sample indices: a slot holds the index t of its owner, whose position is recomputed from
// the node arrays for the comparison, so no key has to be published next to the claim.
__device__ __forceinline__ f3 latticeSamplePosition(const float4* __restrict__ centerHalf, uint32_t t) {
    const float4 ch = centerHalf[t / 19u];
    const int L = cSampleLattice[t % 19u];
    const f3 rel = mk3(float(L % 3 - 1), float((L / 3) % 3 - 1), float(L / 9 - 1));
    return mk3(ch.x, ch.y, ch.z) + rel * ch.w;
}

__global__ void dedupeInsertKernel(const float4* __restrict__ centerHalf, uint32_t nSamples, uint32_t* table, uint32_t mask, uint32_t* slotOf) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nSamples) return;
    const f3 p = latticeSamplePosition(centerHalf, t);
    const uint32_t bx = __float_as_uint(p.x), by = __float_as_uint(p.y), bz = __float_as_uint(p.z);
    uint32_t h = bx * 0x9E3779B1u ^ by * 0x85EBCA77u ^ bz * 0xC2B2AE3Du;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    uint32_t slot = h & mask;
    for (;;) {
        const uint32_t old = atomicCAS(&table[slot], 0xFFFFFFFFu, t);
        if (old == 0xFFFFFFFFu) break;
        const f3 q = latticeSamplePosition(centerHalf, old);   // any member of the slot's group has the group's position
        if (__float_as_uint(q.x) == bx && __float_as_uint(q.y) == by && __float_as_uint(q.z) == bz) {
            atomicMin(&table[slot], t);   // the owner is the LOWEST sample index: identical on every rank of a collective build
            break;
        }
        slot = (slot + 1) & mask;
    }
    slotOf[t] = slot;
}

__global__ void dedupeResolveKernel(const uint32_t* __restrict__ table, uint32_t nSamples, uint32_t* rep /* in: slot, out: owner */, uint32_t* isOwner) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nSamples) return;
    const uint32_t owner = table[rep[t]];
    rep[t] = owner;
    isOwner[t] = owner == t ? 1u : 0u;
}

__global__ void dedupeOwnersKernel(const uint32_t* __restrict__ isOwner, const uint32_t* __restrict__ pos, uint32_t nSamples, uint32_t* owners) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nSamples && isOwner[t]) owners[pos[t]] = t;
}

__global__ void __launch_bounds__(kBvhThreads)
sampleOwnersKernel(DeviceMesh mesh, const float4* __restrict__ centerHalf, const uint32_t* __restrict__ owners, uint32_t first, uint32_t count,
                   float4* results) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= count) return;
    results[u] = samplePoint(mesh, latticeSamplePosition(centerHalf, owners[first + u]));
}

// ---- EXPERIMENTAL (off unless SDFB200_SAMPLE_REFILL=1; not yet measured on a GPU) ----------------
// Lane-refill schedule for the same traversals. With one sample per thread a lane whose traversal ends waits for the
// longest one of its warp: on the C2 levels the lanes are busy 62-68 % of the warp's trips (tests/model_bvh_traversal.py;
// ~650 node visits per sample, spread widely). Here a lane that runs out of work takes the next unassigned sample of
// the launch (one atomicAdd per refill event, ballot + prefix inside the warp) and the warp leaves when the counter is
// exhausted and every lane has finished. The per-sample arithmetic is untouched — each traversal is still one lane's
// private loop — so the results are bit-identical; the nearest triangle is parked in results[u].x and turned into
// (distance, gradient) by a dense second kernel instead of inside the divergent loop.
__global__ void __launch_bounds__(kBvhThreads)
sampleOwnersRefillKernel(DeviceMesh mesh, const float4* __restrict__ centerHalf, const uint32_t* __restrict__ owners, uint32_t first,
                         uint32_t count, float4* results, uint32_t* counter) {
    constexpr unsigned kFull = 0xffffffffu;
    const BvhStack st = bvhStackOfThread(mesh);
    const unsigned lane = threadIdx.x & 31u;
    BvhCursor c;
    c.active = false;
    c.bestTri = -1; c.best = DBL_MAX; c.sp = 0; c.cur = mesh.rootLink; c.p = mkd(0.0, 0.0, 0.0);
    uint32_t item = 0xFFFFFFFFu;   // sample this lane is traversing for
    bool drained = false;          // warp-uniform: the counter has passed `count`
    for (;;) {
        const unsigned idle = __ballot_sync(kFull, !c.active);
        if (idle) {
            if (!c.active && item != 0xFFFFFFFFu) {
                results[item] = make_float4(__int_as_float(c.bestTri), 0.f, 0.f, 0.f);
                item = 0xFFFFFFFFu;
            }
            if (!drained) {
                const int leader = __ffs(int(idle)) - 1;
                uint32_t base = 0;
                if (int(lane) == leader) base = atomicAdd(counter, uint32_t(__popc(idle)));
                base = __shfl_sync(kFull, base, leader);
                if (!c.active) {
                    const uint32_t mine = base + uint32_t(__popc(idle & ((1u << lane) - 1u)));
                    if (mine < count) {
                        item = mine;
                        const f3 pf = latticeSamplePosition(centerHalf, owners[first + mine]);
                        c.p = mkd(double(pf.x), double(pf.y), double(pf.z));
                        c.best = DBL_MAX; c.bestTri = -1; c.sp = 0; c.cur = mesh.rootLink;
                        c.active = true;
                    }
                }
                drained = base + uint32_t(__popc(idle)) >= count;
            }
            if (drained && __ballot_sync(kFull, c.active) == 0) break;
        }
        if (c.active) {
            if (c.cur >= 0) bvhInnerStep(mesh, c, st);
            else bvhLeafStep(mesh, c, st);
        }
    }
}

__global__ void finishOwnersKernel(DeviceMesh mesh, const float4* __restrict__ centerHalf, const uint32_t* __restrict__ owners, uint32_t first,
                                   uint32_t count, float4* results) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= count) return;
    const uint32_t t = uint32_t(__float_as_int(results[u].x));
    const f3 p = latticeSamplePosition(centerHalf, owners[first + u]);
    f3 g;
    const float4 a = mesh.triVerts[3 * size_t(t)], b = mesh.triVerts[3 * size_t(t) + 1], c = mesh.triVerts[3 * size_t(t) + 2];
    const float d = signedDistGradMesh(p, mesh.tris[t], mk3(a.x, a.y, a.z), mk3(b.x, b.y, b.z), mk3(c.x, c.y, c.z), g);
    results[u] = make_float4(d, g.x, g.y, g.z);
}

inline bool sampleRefillEnabled() {
    static const bool on = [] { const char* e = std::getenv("SDFB200_SAMPLE_REFILL"); return e && e[0] == '1'; }();
    return on;
}

__global__ void dedupeScatterKernel(const uint32_t* __restrict__ rep, const uint32_t* __restrict__ pos, const float4* __restrict__ results,
                                    uint32_t nSamples, float4* out, int stride) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nSamples) return;
    out[size_t(t) * stride] = results[pos[rep[t]]];
    if (stride == 2) out[size_t(t) * 2 + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// The 19 mid-point samples of every node of a level into out[(node * 19 + s) * stride]; returns the number of BVH
// traversals actually run. Workspace buffers are grow-only and live as long as the builder.
struct LevelSampler {
    DevBuf<uint32_t> table, rep, isOwner, pos, owners;
    DevBuf<float4> results, slice;
    DevBuf<uint32_t> refillCounter;   // EXPERIMENTAL lane-refill schedule: next unassigned sample of the launch
    Scanner scanner;
    SampleExchange exchange;   // world > 1: every rank traverses the BVH for its slice of the distinct positions only
    cudaStream_t stream = nullptr;   // where this sampler's launches go (legacy default stream unless a pass runs on a side stream)

    // results[0 .. n) for n work items, item u computed by launch(first, count, dst) -> dst[0 .. count) = items first ..
    template <class Launch> const float4* shared(uint32_t n, Launch launch) {
        if (exchange.world <= 1) {
            results.ensure(std::max<uint32_t>(n, 1));
            if (n) launch(0u, n, results.p);
            return results.p;
        }
        const uint32_t per = (n + exchange.world - 1) / exchange.world;   // equal blocks, the tail of the last ones is padding
        results.ensure(size_t(std::max<uint32_t>(per, 1)) * exchange.world);
        slice.ensure(std::max<uint32_t>(per, 1));
        if (per == 0) return results.p;
        const uint32_t first = std::min<uint64_t>(uint64_t(exchange.rank) * per, n), count = std::min<uint32_t>(per, n - first);
        SDFB_CUDA(cudaMemsetAsync(slice.p, 0, size_t(per) * sizeof(float4), stream));
        if (count) launch(first, count, slice.p);
        if (exchange.allgather(exchange.user, slice.p, results.p, uint64_t(per) * sizeof(float4)) != 0)
            throw Error(SDFB200_ERR_CUDA, "the all-gather hook of the collective build failed");
        return results.p;   // item u sits at block u / per, offset u % per = index u
    }

    uint32_t run(const DeviceMesh& mesh, const float4* centerHalf, uint32_t count, float4* out, int stride) {
        const uint64_t n64 = uint64_t(count) * 19;
        if (n64 >= (uint64_t(1) << 30)) {   // the table (2 n slots, 32-bit sample indices) would pass 2^31 slots: plain path
            if (exchange.world > 1) throw Error(SDFB200_ERR_INVALID, "more than 2^30 samples on one level of a collective build");
            sampleLatticeKernel<<<divUp(n64, kBvhThreads), kBvhThreads, bvhStackBytes(mesh), stream>>>(mesh, centerHalf, count, out, stride);
            return 0xFFFFFFFFu;
        }
        const uint32_t n = uint32_t(n64);
        uint32_t size = 1024;
        while (size < 2 * n) size <<= 1;
        table.ensure(size); rep.ensure(n); isOwner.ensure(n); pos.ensure(n);
        SDFB_CUDA(cudaMemsetAsync(table.p, 0xFF, size_t(size) * 4, stream));
        dedupeInsertKernel<<<divUp(n, 256), 256, 0, stream>>>(centerHalf, n, table.p, size - 1, rep.p);
        dedupeResolveKernel<<<divUp(n, 256), 256, 0, stream>>>(table.p, n, rep.p, isOwner.p);
        const uint32_t nUnique = scanner.run(isOwner.p, pos.p, n, false, stream);
        owners.ensure(nUnique);
        dedupeOwnersKernel<<<divUp(n, 256), 256, 0, stream>>>(isOwner.p, pos.p, n, owners.p);
        const uint32_t* ownersPtr = owners.p;
        const float4* res = shared(nUnique, [&](uint32_t first, uint32_t cnt, float4* dst) {
            if (sampleRefillEnabled()) {   // EXPERIMENTAL, see sampleOwnersRefillKernel
                refillCounter.ensure(1);
                SDFB_CUDA(cudaMemsetAsync(refillCounter.p, 0, sizeof(uint32_t), stream));
                const uint32_t blocks = std::min<uint32_t>(divUp(cnt, kBvhThreads), 148u * 8u);
                sampleOwnersRefillKernel<<<blocks, kBvhThreads, bvhStackBytes(mesh), stream>>>(mesh, centerHalf, ownersPtr, first, cnt, dst, refillCounter.p);
                finishOwnersKernel<<<divUp(cnt, 256), 256, 0, stream>>>(mesh, centerHalf, ownersPtr, first, cnt, dst);
                return;
            }
            sampleOwnersKernel<<<divUp(cnt, kBvhThreads), kBvhThreads, bvhStackBytes(mesh), stream>>>(mesh, centerHalf, ownersPtr, first, cnt, dst);
        });
        dedupeScatterKernel<<<divUp(n, 256), 256, 0, stream>>>(rep.p, pos.p, res, n, out, stride);
        return nUnique;
    }

    // explicit point list (fix-up pass of the CONTINUITY builder)
    const float4* runPoints(const DeviceMesh& mesh, const float4* points, uint32_t n);
};

// explicit point list (fix-up pass of the CONTINUITY builder)
__global__ void __launch_bounds__(kBvhThreads) samplePointsKernel(DeviceMesh mesh, const float4* points, uint32_t first, uint32_t n, float4* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float4 p = points[first + t];
    out[t] = samplePoint(mesh, mk3(p.x, p.y, p.z));
}

inline const float4* LevelSampler::runPoints(const DeviceMesh& mesh, const float4* points, uint32_t n) {
    return shared(n, [&](uint32_t first, uint32_t cnt, float4* dst) {
        samplePointsKernel<<<divUp(cnt, kBvhThreads), kBvhThreads, bvhStackBytes(mesh), stream>>>(mesh, points, first, cnt, dst);
    });
}

__global__ void gatherTriVertsKernel(const f3* verts, const uint32_t* idx, uint32_t nTris, float4* triVerts) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTris * 3) return;
    const f3 v = verts[idx[t]];
    triVerts[t] = make_float4(v.x, v.y, v.z, 0.f);
}

// interpolateValue, scalar branch: acc = 0 + sum_n ((c_n * x^i) * y^j) * z^k, n ascending, left to right.
__device__ __forceinline__ float polyValueExact(const float* c, float x, float y, float z) {
    float acc = 0.0f;
#pragma unroll
    for (int n = 0; n < 64; n++) {
        float t = c[n];
#pragma unroll
        for (int a = 0; a < (n & 3); a++) t *= x;
#pragma unroll
        for (int a = 0; a < ((n >> 2) & 3); a++) t *= y;
#pragma unroll
        for (int a = 0; a < (n >> 4); a++) t *= z;
        acc += t;
    }
    return acc;
}

// One Hermite row: left-to-right float sum of w * in[col] over the row's non-zero columns.
// in[col] = corner value scaled by nodeSize^order (slots 4..7 are the always-zero mixed derivatives).
__device__ __forceinline__ float hermiteRow(const HermiteTable& tab, int row, const float4* lattice, float nodeSize) {
    const int b = tab.rowStart[row], e = tab.rowStart[row + 1];
    float acc = 0.0f;
    for (int i = b; i < e; i++) {
        const int col = tab.col[i], c = col >> 3, s = col & 7;
        // corner c = (x,y,z) bits -> lattice point (2x, 2y, 2z)
        const float4 v = lattice[2 * (c & 1) + 6 * ((c >> 1) & 1) + 18 * (c >> 2)];
        float in;
        if (s == 0) in = v.x;
        else if (s == 1) in = v.y * nodeSize;
        else if (s == 2) in = v.z * nodeSize;
        else if (s == 3) in = v.w * nodeSize;
        else in = 0.0f;   // 0 * nodeSize^2 (or ^3) = +0
        const float term = float(int(tab.weight[i])) * in;
        acc = (i == b) ? term : acc + term;
    }
    return acc;
}

void uploadHermite() {
    static bool done[64] = {};
    int dev = 0;
    SDFB_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && done[dev]) return;
    const HermiteTable t = makeHermiteTable();
    SDFB_CUDA(cudaMemcpyToSymbol(cHermite, &t, sizeof(t)));
    if (dev < 64) done[dev] = true;
}

inline void uploadMesh(MeshOnDevice& m, const HostMesh& mesh, const TriVec& tris, const RawVec<BvhNode>* bvh) {
    m.numTriangles = mesh.numTriangles();
    m.verts.alloc(mesh.nVerts); m.verts.upload(mesh.verts, mesh.nVerts);
    m.idx.alloc(mesh.nIdx); m.idx.upload(mesh.idx, mesh.nIdx);
    m.tris.alloc(tris.size()); m.tris.upload(tris.data(), tris.size());
    if (bvh) {
        m.bvh.alloc(bvh->size()); m.bvh.upload(bvh->data(), bvh->size());
        m.rootLink = (*bvh)[0].pad[0] ? ~(*bvh)[0].right : 0;   // single-triangle mesh: the root is a leaf
        // height of the median-split tree (mesh_host.cpp: halves of floor / ceil size) = deepest possible stack, + 1 spare
        uint32_t n = m.numTriangles, h = 0;
        while (n > 1) { n = n - n / 2; h++; }
        m.stackDepth = int(h) + 1;
        m.triVerts.alloc(size_t(m.numTriangles) * 3);
        gatherTriVertsKernel<<<divUp(uint64_t(m.numTriangles) * 3, 256), 256>>>(m.verts.p, m.idx.p, m.numTriangles, m.triVerts.p);
    }
}

// The two serial set-up steps of the reference constructors. Running them side by side was measured twice (OpenMP
// task team and plain threads in the BVH builder): TriangleData saturates every host core, so the BVH thread only
// gets going once it is done — no gain; they run one after the other.
inline void buildHostStructures(const HostMesh& mesh, TriVec& tris, RawVec<BvhNode>& bvh, sdfb200_build_stats& st) {
    auto t0 = std::chrono::steady_clock::now();
    tris = computeTriangleData(mesh);
    st.triangle_data_ms = msSince(t0);
    t0 = std::chrono::steady_clock::now();
    bvh = buildBvh(mesh);
    st.bvh_ms = msSince(t0);
}

// Weight of a mid-point in the error integral: trapezoid / by-distance 2^k/64 (OctreeSdfUtils.h:60-138),
// Simpson 4^k/216 (:213-238), k = number of centred axes; the float constant is formed first, as in the
// reference expression `w / 64.0f * pow2(..)`.
__device__ __forceinline__ float errorWeight(int rule, int centred) {
    if (rule == SDFB200_RULE_SIMPSONS) return (centred == 1 ? 4.0f : (centred == 2 ? 16.0f : 64.0f)) / 216.0f;
    return (centred == 1 ? 2.0f : (centred == 2 ? 4.0f : 8.0f)) / 64.0f;
}

}  // namespace
}  // namespace sdfb200
