// The kernels of the bulk OctreeSdf query (octree_query.cu is the translation unit: see its header for the design).
// Kept in a header so that tests/cpp/simt_query_main.cpp can run the very same source on the CPU under a lock-step
// warp emulation (32 host threads per warp, collectives = barrier exchanges; it defines SDFB_SIMT_HOST, which replaces
// the TMA staging of the tile kernel by plain loads): control flow — rounds, leaders, class membership, index packing,
// the division-free cell selection — is checked against the plain kernel and the oracle without a GPU.
// Include inside namespace sdfb200, inside an anonymous namespace.
#pragma once


struct QueryParams {
    float minx, miny, minz;
    float maxx, maxy, maxz;
    float cell;
    int grid;
    float minBorder;
};

__device__ __forceinline__ float boxDistance(const QueryParams& q, f3 p) {   // Mesh.h:42-46
    const f3 size = mk3(q.maxx - q.minx, q.maxy - q.miny, q.maxz - q.minz);
    const f3 center = mk3(q.minx, q.miny, q.minz) + 0.5f * size;
    const f3 d = p - center;
    const f3 h = 0.5f * size;
    const f3 a = mk3(gabs(d.x) - h.x, gabs(d.y) - h.y, gabs(d.z) - h.z);
    const f3 ap = mk3(gmax(a.x, 0.0f), gmax(a.y, 0.0f), gmax(a.z, 0.0f));
    return sqrtf(dot3(ap, ap)) + gmin(gmax(a.x, gmax(a.y, a.z)), 0.0f);
}

// Mesh.h:48-63. Kept bug-compatible: it measures |p| - size (not centred on the box) and leaves the
// components it does not write untouched (the caller's gradient is zero-initialised here).
__device__ __forceinline__ float boxDistanceGrad(const QueryParams& q, f3 p, f3& g) {
    const float s[3] = {q.maxx - q.minx, q.maxy - q.miny, q.maxz - q.minz};
    const float pp[3] = {p.x, p.y, p.z};
    float a[3], gg[3] = {g.x, g.y, g.z};
    for (int i = 0; i < 3; i++) a[i] = gabs(pp[i]) - s[i];
    const int k = a[0] > a[1] ? 0 : 1;
    const int l = a[2] > a[k] ? 2 : k;
    if (a[l] < 0) gg[l] = pp[l] / gabs(pp[l]);
    else {
        float b[3];
        for (int i = 0; i < 3; i++) b[i] = gmax(a[i], 0.0f);
        const float tx = b[0] * b[0], ty = b[1] * b[1], tz = b[2] * b[2];
        const float c = sqrtf(tx + ty + tz);
        for (int i = 0; i < 3; i++) gg[i] = a[i] > 0 ? b[i] / c * pp[i] / gabs(pp[i]) : 0.0f;
    }
    g = mk3(gg[0], gg[1], gg[2]);
    return boxDistance(q, p);
}

#ifdef SDFB_QUERY_EXACT
__device__ __forceinline__ float monomialExact(float c, int i, int j, int k, float x, float y, float z) {
    float t = c;
    for (int a = 0; a < i; a++) t *= x;
    for (int a = 0; a < j; a++) t *= y;
    for (int a = 0; a < k; a++) t *= z;
    return t;
}
__device__ __forceinline__ float polyValue(const float* c, float x, float y, float z) {
    float acc = 0.0f;
#pragma unroll
    for (int n = 0; n < 64; n++) acc += monomialExact(c[n], n & 3, (n >> 2) & 3, n >> 4, x, y, z);
    return acc;
}
// interpolateGradient (InterpolationMethods.h:442-455): per component, terms in ascending n, the
// integer factor multiplies the coefficient first, no leading zero.
template <int AX> __device__ __forceinline__ float polyDerivative(const float* c, float x, float y, float z) {
    float acc = 0.0f;
    bool first = true;
#pragma unroll
    for (int n = 0; n < 64; n++) {
        const int i = n & 3, j = (n >> 2) & 3, k = n >> 4;
        const int p = AX == 0 ? i : (AX == 1 ? j : k);
        if (p == 0) continue;
        const float t = monomialExact(float(p) * c[n], i - (AX == 0), j - (AX == 1), k - (AX == 2), x, y, z);
        acc = first ? t : acc + t;
        first = false;
    }
    return acc;
}
template <bool kGrad, bool kVec> __device__ __forceinline__ float evalLeaf(const float* c, float x, float y, float z, f3& g) {
    if (kGrad) g = normalize3(mk3(polyDerivative<0>(c, x, y, z), polyDerivative<1>(c, x, y, z), polyDerivative<2>(c, x, y, z)));
    return polyValue(c, x, y, z);
}
#else
// Horner in x, then y, then z with derivative recurrences; all FMA. kVec: the 64 coefficients are
// fetched as 16 x 128-bit read-only loads (legal whenever leaf blocks are 16-byte aligned, which holds
// for every start grid with G^3 % 4 == 0 because all blocks are 8 or 64 words long).
template <bool kGrad, bool kVec> __device__ __forceinline__ float evalLeaf(const float* c, float x, float y, float z, f3& g) {
    float v = 0.0f, vx = 0.0f, vy = 0.0f, vz = 0.0f;
#pragma unroll
    for (int k = 3; k >= 0; k--) {
        float a = 0.0f, ax = 0.0f, ay = 0.0f;
#pragma unroll
        for (int j = 3; j >= 0; j--) {
            float c0, c1, c2, c3;
            if (kVec) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(c) + 4 * k + j);
                c0 = q.x; c1 = q.y; c2 = q.z; c3 = q.w;
            } else {
                c0 = __ldg(c + 16 * k + 4 * j); c1 = __ldg(c + 16 * k + 4 * j + 1);
                c2 = __ldg(c + 16 * k + 4 * j + 2); c3 = __ldg(c + 16 * k + 4 * j + 3);
            }
            const float r = fmaf(fmaf(fmaf(c3, x, c2), x, c1), x, c0);
            if (kGrad) {
                const float rx = fmaf(fmaf(3.0f * c3, x, 2.0f * c2), x, c1);
                ay = fmaf(ay, y, a);
                ax = fmaf(ax, y, rx);
            }
            a = fmaf(a, y, r);
        }
        if (kGrad) {
            vz = fmaf(vz, z, v);
            vx = fmaf(vx, z, ax);
            vy = fmaf(vy, z, ay);
        }
        v = fmaf(v, z, a);
    }
    if (kGrad) g = normalize3(mk3(vx, vy, vz));
    return v;
}
#endif

template <bool kGrad, bool kVec>
__global__ void __launch_bounds__(256)
octreeQueryKernel(const uint32_t* __restrict__ oct, const QueryParams q, const float* __restrict__ xyz, uint64_t n,
                  float* __restrict__ dist, float* __restrict__ grad) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f3 p = mk3(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2));
    float fx = (p.x - q.minx) / q.cell, fy = (p.y - q.miny) / q.cell, fz = (p.z - q.minz) / q.cell;
    const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    const int ix = int(flx), iy = int(fly), iz = int(flz);
    fx -= flx; fy -= fly; fz -= flz;
    f3 g = mk3(0.0f, 0.0f, 0.0f);
    float d;
    if (ix < 0 || ix >= q.grid || iy < 0 || iy >= q.grid || iz < 0 || iz >= q.grid) {
        d = (kGrad ? boxDistanceGrad(q, p, g) : boxDistance(q, p)) + q.minBorder;
    } else {
        // The reference descends with child = (frac >= 0.5) per axis and frac <- fract(2 frac). Both
        // steps are exact in binary floating point (doubling, floor and the subtraction introduce no
        // rounding), so the same path and the same final frac are obtained from the leading bits of the
        // start-cell fraction: level j uses bit j of floor(frac * 2^kPathBits), and a leaf reached after k
        // steps evaluates at frac * 2^k - floor(frac * 2^k).
        constexpr int kPathBits = 16;
        const uint32_t bx = uint32_t(fx * float(1 << kPathBits)), by = uint32_t(fy * float(1 << kPathBits)),
                       bz = uint32_t(fz * float(1 << kPathBits));
        uint32_t node = __ldg(oct + (iz * q.grid + iy) * q.grid + ix);
        int k = 0;
        while (!(node & kLeafBit)) {
            const int sh = kPathBits - 1 - k;
            const uint32_t child = ((bx >> sh) & 1u) | (((by >> sh) & 1u) << 1) | (((bz >> sh) & 1u) << 2);
            node = __ldg(oct + (node & kOctIndexMask) + child);
            k++;
        }
        const float scale = float(1u << k);
        fx *= scale; fy *= scale; fz *= scale;
        fx -= floorf(fx); fy -= floorf(fy); fz -= floorf(fz);
        const float* c = reinterpret_cast<const float*>(oct + (node & kOctIndexMask));
        d = evalLeaf<kGrad, kVec>(c, fx, fy, fz, g);
    }
    dist[i] = d;
    if (kGrad) { grad[3 * i] = g.x; grad[3 * i + 1] = g.y; grad[3 * i + 2] = g.z; }
}

// ---- the default kernel of the FMA path: warp-per-query-batch, TMA-staged point tiles, quad-cooperative evaluation ----------
// Measured on the 256^3 workload (profiles/r2_summary.md): the one-query-per-thread kernel above is bound by the LSU
// data pipe (every lane fills 64 coefficient registers: 16 x LDG.128, one wavefront per DISTINCT leaf in the warp) and,
// once that is cut, by the latency of its dependent chain (point -> 5-6 node gathers -> coefficients) at 58 % achieved
// occupancy and by XU-pipe conversions (IEEE division, floor, float <-> int). This kernel removes each of them:
//   * every WARP is a persistent worker over 32-query tiles (tile = 384 contiguous bytes of xyz): lane 0 fetches the
//     next tile with one 1-D TMA bulk copy into the warp's double-buffered shared-memory slot while the current tile
//     is evaluated, so the point loads cost no LSU wavefronts and their DRAM latency is off the chain; unit gradients
//     leave the same way (shared memory -> one bulk store of 384 bytes) instead of stride-3 scalar stores;
//   * a dense TOP INDEX (one word per cell of the grid `topLevels` below the start grid, <= 2^21 cells = 8 MB,
//     L2-resident) replaces the first `topLevels` dependent gathers by one load that is coalesced for coherent
//     batches; only leaves deeper than the index (0.2 % of the 256^3 queries) continue the descent;
//   * the cell selection uses no division and no conversion instruction but gives the reference's bits:
//     (p - min) / cell is formed as two Newton steps on x * RN(1 / cell) with exact FMA residuals (Markstein: the
//     result is the correctly rounded quotient for normal operands; tests/cpp/simt_query_main.cpp and
//     tests/test_query_variants_model.py check it against IEEE division), floor() is an add of 2^23 rounded towards
//     minus infinity, the integer cell and the 16 path bits are read from the mantissa of that sum;
//   * the polynomial is evaluated by the four lanes of an aligned quad working on one (leaf, y, z) class at a time:
//     lane r loads the vectors c[0..3][j = r][k = 0..3] (64 contiguous bytes per quad and step), forms
//     y^r * Horner_z, two butterfly shuffles give every lane the same sums A_i(y, z), each member evaluates the cubic
//     in its own x. Grid rows put 4-8 neighbours of a row into one class; unrelated points take one round per lane
//     and still use every loaded byte.
// Same leaf and same leaf-local fractions as the reference; only the summation order of the polynomial differs
// (independent of where in the batch a query sits), within the tolerance of the FMA path.
#ifndef SDFB_QUERY_EXACT
struct TileQuery {
    float rcell;             // RN(1 / cell), computed on the host
    float gridf;             // float(grid)
    int gridShift;           // log2(grid)
    int topLevels;           // levels below the start grid the top index resolves
    uint32_t G3;             // grid^3
    uint32_t tmaPoints;      // xyz is 16-byte aligned: tiles are fetched by TMA
    uint32_t tmaGrad;        // grad is 16-byte aligned: gradients leave by bulk stores
};

// top index word: bit 31 = leaf, bits 27-30 = steps below the start grid at which the leaf / node was reached,
// bits 0-26 = (block - G^3) / 8 (every block of the array is 8 or 64 words long and starts after the G^3 start words)
constexpr uint32_t kTopLeaf = 1u << 31;
constexpr uint32_t kTopBlockMask = (1u << 27) - 1u;
constexpr int kPathBits = 16;

__global__ void __launch_bounds__(256)
topIndexKernel(const uint32_t* __restrict__ oct, int grid, int levels, uint32_t* __restrict__ index, uint32_t* __restrict__ bad) {
    const uint32_t N = uint32_t(grid) << levels;
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= uint64_t(N) * N * N) return;
    const uint32_t cx = uint32_t(i % N), cy = uint32_t((i / N) % N), cz = uint32_t(i / (uint64_t(N) * N));
    const uint32_t G3 = uint32_t(grid) * uint32_t(grid) * uint32_t(grid);
    uint32_t node = __ldg(oct + ((cz >> levels) * uint32_t(grid) + (cy >> levels)) * uint32_t(grid) + (cx >> levels));
    int k = 0;
    while (!(node & kLeafBit) && k < levels) {
        const int sh = levels - 1 - k;
        const uint32_t child = ((cx >> sh) & 1u) | (((cy >> sh) & 1u) << 1) | (((cz >> sh) & 1u) << 2);
        node = __ldg(oct + (node & kOctIndexMask) + child);
        k++;
    }
    const uint32_t block = node & kOctIndexMask;
    if (block < G3 || ((block - G3) & 7u) || ((block - G3) >> 3) > kTopBlockMask) { atomicOr(bad, 1u); return; }
    index[i] = ((block - G3) >> 3) | (uint32_t(k) << 27) | ((node & kLeafBit) ? kTopLeaf : 0u);
}

// (x - lo) / cell, correctly rounded, without a division: q0 = x * rc, two steps q <- q + (x - q * cell) * rc with exact
// residuals. Operands far below the normal range (|x| < 2^-100: the quotient is a denormal fraction of cell 0 either way)
// may differ from IEEE division in their last bits; no cell or leaf decision depends on them.
__device__ __forceinline__ float cellCoordinate(float x, float cell, float rc) {
    float qv = __fmul_rn(x, rc);
    qv = __fmaf_rn(__fmaf_rn(-qv, cell, x), rc, qv);
    return __fmaf_rn(__fmaf_rn(-qv, cell, x), rc, qv);
}
// floor(v) for 0 <= v < 2^22 and the integer itself, from the mantissa of v + 2^23 rounded towards minus infinity
__device__ __forceinline__ float floorSmall(float v, uint32_t& asInt) {
    const float t = __fadd_rd(v, 8388608.0f);
    asInt = uint32_t(__float_as_int(t)) & 0x7FFFFFu;
    return __fadd_rn(t, -8388608.0f);
}

#ifndef SDFB_SIMT_HOST
__device__ __forceinline__ void qMbarInit(uint64_t* bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(uint32_t(__cvta_generic_to_shared(bar))));
}
__device__ __forceinline__ void qTileLoad(void* smemDst, const void* gmemSrc, uint64_t* bar) {   // 384 bytes, 16-byte aligned both sides
    const uint32_t b = uint32_t(__cvta_generic_to_shared(bar));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 384;" ::"r"(b) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 384, [%2];" ::"r"(
                     uint32_t(__cvta_generic_to_shared(smemDst))), "l"(gmemSrc), "r"(b) : "memory");
}
__device__ __forceinline__ void qMbarWait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "QWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra QDONE_%=;\n"
        "bra QWAIT_%=;\n"
        "QDONE_%=:\n"
        "}\n" ::"r"(uint32_t(__cvta_generic_to_shared(bar))), "r"(parity) : "memory");
}
__device__ __forceinline__ void qTileStore(void* gmemDst, const void* smemSrc) {   // 384 bytes shared -> global, one bulk group
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 384;" ::"l"(gmemDst), "r"(uint32_t(__cvta_generic_to_shared(smemSrc))) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
#endif

constexpr int kTileWarps = 8;   // warps per CTA of the tile kernel

template <bool kGrad>
__global__ void __launch_bounds__(kTileWarps * 32, kGrad ? 6 : 8)
octreeQueryTileKernel(const uint32_t* __restrict__ oct, const uint32_t* __restrict__ top, const QueryParams q, const TileQuery tq,
                      const float* __restrict__ xyz, uint64_t n, float* __restrict__ dist, float* __restrict__ grad) {
    constexpr unsigned kFull = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tiles = uint32_t((n + 31) >> 5), fullTiles = uint32_t(n >> 5);   // the launcher keeps n below 2^36
    const uint32_t stride = gridDim.x * kTileWarps;
    uint32_t t = blockIdx.x * kTileWarps + warp;
#ifndef SDFB_SIMT_HOST
    __shared__ alignas(16) float sPts[kTileWarps][2][96];
    __shared__ alignas(16) float sGrad[kGrad ? kTileWarps : 1][96];
    __shared__ alignas(8) uint64_t sBar[kTileWarps][2];
    if (lane == 0) {
        qMbarInit(&sBar[warp][0]);
        qMbarInit(&sBar[warp][1]);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (tq.tmaPoints && lane == 0 && t < fullTiles) qTileLoad(sPts[warp][0], xyz + uint64_t(t) * 96, &sBar[warp][0]);
#endif
    for (uint32_t it = 0; t < tiles; it++) {
        const uint64_t i = (uint64_t(t) << 5) + lane;
        const bool valid = i < n;
        f3 p;
#ifndef SDFB_SIMT_HOST
        if (tq.tmaPoints && t < fullTiles) {
            const int s = int(it & 1u);
            __syncwarp();   // every lane has taken its point of the previous tile out of slot s ^ 1
            if (lane == 0 && uint64_t(t) + stride < fullTiles) qTileLoad(sPts[warp][s ^ 1], xyz + (uint64_t(t) + stride) * 96, &sBar[warp][s ^ 1]);
            qMbarWait(&sBar[warp][s], (it >> 1) & 1u);
            const float* sp = sPts[warp][s] + 3 * lane;
            p = mk3(sp[0], sp[1], sp[2]);
        } else
#endif
            p = valid ? mk3(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2)) : mk3(0.0f, 0.0f, 0.0f);
        float fx = cellCoordinate(__fadd_rn(p.x, -q.minx), q.cell, tq.rcell), fy = cellCoordinate(__fadd_rn(p.y, -q.miny), q.cell, tq.rcell),
              fz = cellCoordinate(__fadd_rn(p.z, -q.minz), q.cell, tq.rcell);
        // floor(f) in [0, grid) <=> 0 <= f < grid (NaN compares false: outside, like the reference's int conversion)
        const bool inside = valid && fx >= 0.0f && fy >= 0.0f && fz >= 0.0f && fx < tq.gridf && fy < tq.gridf && fz < tq.gridf;
        f3 g = mk3(0.0f, 0.0f, 0.0f);
        float d = 0.0f;
        uint32_t block = 0;
        if (valid && !inside) d = (kGrad ? boxDistanceGrad(q, p, g) : boxDistance(q, p)) + q.minBorder;
        if (inside) {
            uint32_t ix, iy, iz, bx, by, bz;
            fx = __fadd_rn(fx, -floorSmall(fx, ix)); fy = __fadd_rn(fy, -floorSmall(fy, iy)); fz = __fadd_rn(fz, -floorSmall(fz, iz));
            // path bits: bit j (from the top) of floor(frac * 2^16) is the child choice at level j (doubling, floor and
            // subtraction are exact in binary floating point, so this is the reference's fract(2 f) chain)
            floorSmall(__fmul_rn(fx, 65536.0f), bx); floorSmall(__fmul_rn(fy, 65536.0f), by); floorSmall(__fmul_rn(fz, 65536.0f), bz);
            const int L = tq.topLevels, ns = tq.gridShift + L;
            const uint32_t cx = (ix << L) | (bx >> (kPathBits - L)), cy = (iy << L) | (by >> (kPathBits - L)), cz = (iz << L) | (bz >> (kPathBits - L));
            const uint32_t e = __ldg(top + ((((uint64_t(cz) << ns) | cy) << ns) | cx));
            block = ((e & kTopBlockMask) << 3) + tq.G3;
            int k = int((e >> 27) & 15u);
            if (!(e & kTopLeaf)) {
                for (;;) {   // leaves below the index: finish the descent from the cell's node
                    const int sh = kPathBits - 1 - k;
                    const uint32_t child = ((bx >> sh) & 1u) | (((by >> sh) & 1u) << 1) | (((bz >> sh) & 1u) << 2);
                    const uint32_t node = __ldg(oct + block + child);
                    k++;
                    block = node & kOctIndexMask;
                    if (node & kLeafBit) break;
                }
            }
            const float scale = __int_as_float((127 + k) << 23);   // 2^k
            uint32_t unused;
            fx = __fmul_rn(fx, scale); fy = __fmul_rn(fy, scale); fz = __fmul_rn(fz, scale);
            fx = __fadd_rn(fx, -floorSmall(fx, unused)); fy = __fadd_rn(fy, -floorSmall(fy, unused)); fz = __fadd_rn(fz, -floorSmall(fz, unused));
        }
        bool pending = inside;
        const unsigned quadBase = lane & ~3u, r = lane & 3u;
        const unsigned quadMask = 0xFu << quadBase;
        for (;;) {
            const unsigned pend = __ballot_sync(kFull, pending);
            if (pend == 0) break;                                   // warp-uniform
            const unsigned mine = pend & quadMask;
            const int leader = mine ? __ffs(int(mine)) - 1 : int(quadBase);   // a finished quad idles through the shuffles
            const uint32_t lb = __shfl_sync(kFull, block, leader);
            const float ly = __shfl_sync(kFull, fy, leader), lz = __shfl_sync(kFull, fz, leader);
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;       // this lane's part of A_i
            float y0 = 0.0f, y1 = 0.0f, y2 = 0.0f, y3 = 0.0f;       // ... of dA_i/dy
            float z0 = 0.0f, z1 = 0.0f, z2 = 0.0f, z3 = 0.0f;       // ... of dA_i/dz
            if (mine) {
                const float4* c = reinterpret_cast<const float4*>(oct + lb) + r;   // vector m = r + 4 k holds c[0..3][j = r][k]
                const float4 c0 = __ldg(c), c1 = __ldg(c + 4), c2 = __ldg(c + 8), c3 = __ldg(c + 12);
                const float yy = ly * ly;
                const float yr = r == 0 ? 1.0f : (r == 1 ? ly : (r == 2 ? yy : yy * ly));           // y^r
                const float h0 = fmaf(fmaf(fmaf(c3.x, lz, c2.x), lz, c1.x), lz, c0.x), h1 = fmaf(fmaf(fmaf(c3.y, lz, c2.y), lz, c1.y), lz, c0.y);
                const float h2 = fmaf(fmaf(fmaf(c3.z, lz, c2.z), lz, c1.z), lz, c0.z), h3 = fmaf(fmaf(fmaf(c3.w, lz, c2.w), lz, c1.w), lz, c0.w);
                a0 = yr * h0; a1 = yr * h1; a2 = yr * h2; a3 = yr * h3;
                if (kGrad) {
                    const float dyr = r == 0 ? 0.0f : (r == 1 ? 1.0f : (r == 2 ? 2.0f * ly : 3.0f * yy));   // d y^r / dy
                    y0 = dyr * h0; y1 = dyr * h1; y2 = dyr * h2; y3 = dyr * h3;
                    z0 = yr * fmaf(fmaf(3.0f * c3.x, lz, 2.0f * c2.x), lz, c1.x);
                    z1 = yr * fmaf(fmaf(3.0f * c3.y, lz, 2.0f * c2.y), lz, c1.y);
                    z2 = yr * fmaf(fmaf(3.0f * c3.z, lz, 2.0f * c2.z), lz, c1.z);
                    z3 = yr * fmaf(fmaf(3.0f * c3.w, lz, 2.0f * c2.w), lz, c1.w);
                }
            }
#pragma unroll
            for (int m = 1; m <= 2; m <<= 1) {                     // butterfly over the quad: all four lanes end with the same sums
                a0 += __shfl_xor_sync(kFull, a0, m); a1 += __shfl_xor_sync(kFull, a1, m);
                a2 += __shfl_xor_sync(kFull, a2, m); a3 += __shfl_xor_sync(kFull, a3, m);
                if (kGrad) {
                    y0 += __shfl_xor_sync(kFull, y0, m); y1 += __shfl_xor_sync(kFull, y1, m);
                    y2 += __shfl_xor_sync(kFull, y2, m); y3 += __shfl_xor_sync(kFull, y3, m);
                    z0 += __shfl_xor_sync(kFull, z0, m); z1 += __shfl_xor_sync(kFull, z1, m);
                    z2 += __shfl_xor_sync(kFull, z2, m); z3 += __shfl_xor_sync(kFull, z3, m);
                }
            }
            if (pending && block == lb && __float_as_uint(fy) == __float_as_uint(ly) && __float_as_uint(fz) == __float_as_uint(lz)) {
                d = fmaf(fmaf(fmaf(a3, fx, a2), fx, a1), fx, a0);
                if (kGrad) {
                    const float gx = fmaf(fmaf(3.0f * a3, fx, 2.0f * a2), fx, a1);
                    const float gy = fmaf(fmaf(fmaf(y3, fx, y2), fx, y1), fx, y0);
                    const float gz = fmaf(fmaf(fmaf(z3, fx, z2), fx, z1), fx, z0);
                    g = normalize3(mk3(gx, gy, gz));
                }
                pending = false;   // the leader always matches itself bit for bit, so every round retires at least one lane per quad
            }
        }
        if (valid) dist[i] = d;
        if (kGrad) {
#ifndef SDFB_SIMT_HOST
            if (tq.tmaGrad && t < fullTiles) {
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous tile's store has read the slot
                __syncwarp();
                float* sg = sGrad[warp] + 3 * lane;
                sg[0] = g.x; sg[1] = g.y; sg[2] = g.z;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) qTileStore(grad + uint64_t(t) * 96, sGrad[warp]);
            } else
#endif
            if (valid) { grad[3 * i] = g.x; grad[3 * i + 1] = g.y; grad[3 * i + 2] = g.z; }
        }
        if (tiles - t <= stride) break;   // t + stride could wrap
        t += stride;
    }
#ifndef SDFB_SIMT_HOST
    if (kGrad && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
}
#endif
