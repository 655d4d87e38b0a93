// BVH sampling device code of the OctreeSdf builders: float64 sphere-BVH nearest-triangle traversal in the reference's
// order, point sampling, and the kernels that run one traversal per distinct sample position (plain and lane-refill
// schedules). Part of octree_device.cuh (included from there, inside namespace sdfb200 { namespace { ); kept in its own
// header so that tests/cpp/simt_sampler_main.cpp can run the same source on the CPU under a lock-step warp emulation.
#pragma once

// lattice index L = x + 3y + 9z of the 19 mid-points, in the reference's sample order
__constant__ int cSampleLattice[19] = {1, 3, 4, 5, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 19, 21, 22, 23, 25};

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

// position of sample t = node * 19 + s of a level (centre + lattice offset * half size)
__device__ __forceinline__ f3 latticeSamplePosition(const float4* __restrict__ centerHalf, uint32_t t) {
    const float4 ch = centerHalf[t / 19u];
    const int L = cSampleLattice[t % 19u];
    const f3 rel = mk3(float(L % 3 - 1), float((L / 3) % 3 - 1), float(L / 9 - 1));
    return mk3(ch.x, ch.y, ch.z) + rel * ch.w;
}

__global__ void __launch_bounds__(kBvhThreads)
sampleOwnersKernel(DeviceMesh mesh, const float4* __restrict__ centerHalf, const uint32_t* __restrict__ owners, uint32_t first, uint32_t count,
                   float4* results) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= count) return;
    results[u] = samplePoint(mesh, latticeSamplePosition(centerHalf, owners[first + u]));
}

// ---- default schedule (SDFB200_SAMPLE_REFILL=0 selects sampleOwnersKernel above) ------------------------------------------
// Lane-refill schedule for the same traversals. With one sample per thread a lane whose traversal ends waits for the
// longest one of its warp: on the C2 levels the lanes are busy 62-68 % of the warp's trips (tests/model_bvh_traversal.py;
// ~650 node visits per sample, spread widely). Here a lane that runs out of work takes the next unassigned sample of
// the launch (one atomicAdd per refill event, ballot + prefix inside the warp) and the warp leaves when the counter is
// exhausted and every lane has finished. The per-sample arithmetic is untouched — each traversal is still one lane's
// private loop — so the results are bit-identical; the nearest triangle is parked in results[u].x and turned into
// (distance, gradient) by a dense second kernel instead of inside the divergent loop.
__global__ void __launch_bounds__(kBvhThreads)
sampleOwnersRefillKernel(DeviceMesh mesh, const float4* __restrict__ centerHalf, const uint32_t* __restrict__ owners, uint32_t first,
                         uint32_t count, float4* results, uint32_t* counter, int leafBatch, const uint32_t* __restrict__ schedule) {
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
                    const uint32_t slot = base + uint32_t(__popc(idle & ((1u << lane) - 1u)));
                    if (slot < count) {
                        // `schedule` (optional) is the order in which the samples are STARTED: the expensive ones first, so that
                        // their long traversals run under everybody else's work instead of after it (LevelSampler::run)
                        const uint32_t mine = schedule ? schedule[slot] : slot;
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
        // Leaf batching. A fifth of the steps are point-triangle tests, and they are the expensive ones (float64 Eberly: a
        // division, a square root): taken as they come, nearly every trip of the warp runs the leaf branch for ~6 of its 32
        // lanes — two thirds of the issued instructions at a fifth of the lanes. A lane that reaches a leaf therefore WAITS
        // (its traversal is its own: nothing else depends on it) until `leafBatch` lanes of the warp hold one, or no lane has
        // an inner node left; the inner lanes never wait. Per-lane arithmetic and order are untouched: same bits.
        if (c.active && c.cur >= 0) bvhInnerStep(mesh, c, st);
        const unsigned leaves = __ballot_sync(kFull, c.active && c.cur < 0);
        if (leaves) {
            const unsigned inner = __ballot_sync(kFull, c.active && c.cur >= 0);
            if (c.active && c.cur < 0 && (__popc(leaves) >= leafBatch || inner == 0)) bvhLeafStep(mesh, c, st);
        }
    }
}

// Measured and dropped (profiles/r2_summary.md, r2_ab_screened_sampler_packed_query.log; the code is in the history at commit
// 12a8917): a float-screened visit — the three decisions of an inner step (nearer child, near child against the best, far
// child against the best) taken from float32 images inside a rigorous error margin with the float64 path as fallback (one
// visit in 2000), the node a lane works on held in registers and fetched as soon as the lane knows it, the far child's exact
// square root and the leaf's point-triangle test run in the shadow of that load (speculative pop). Bit-identical (3 M samples
// against bvhNearest on the CPU, full-size hashes on the GPU) and no faster: C2 levels 57.1 -> 56.7 ms at 72 registers / 7
// CTAs per SM, 57.4 ms at 78 registers / 6 CTAs. Shortening a visit's dependent chain does not move the kernel: it trades
// ~16 float64 instructions for ~25 float32 ones, and the warp's issue slots (61 % busy at 10 of 32 lanes) are what is spent.

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
