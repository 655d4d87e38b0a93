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

// Packed float32 pairs (sm_100: FFMA2 / FMUL2 / FADD2 — one issue slot for two IEEE operations, each half rounded exactly
// like its scalar instruction, so a packed evaluation has the bits of the scalar one). The tile kernel is issue bound.
struct P2 { float x, y; };
__device__ __forceinline__ P2 mkp(float x, float y) { P2 r; r.x = x; r.y = y; return r; }
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned long long p2Bits(P2 v) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y)); return r; }
__device__ __forceinline__ P2 p2From(unsigned long long b) { P2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(b)); return r; }
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(p2Bits(a)), "l"(p2Bits(b)), "l"(p2Bits(c))); return p2From(r); }
__device__ __forceinline__ P2 mul2(P2 a, P2 b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(p2Bits(a)), "l"(p2Bits(b))); return p2From(r); }
__device__ __forceinline__ P2 add2(P2 a, P2 b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(p2Bits(a)), "l"(p2Bits(b))); return p2From(r); }
__device__ __forceinline__ P2 add2Down(P2 a, P2 b) { unsigned long long r; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(p2Bits(a)), "l"(p2Bits(b))); return p2From(r); }
#else   // the warp emulation of tests/cpp/simt_query_main.cpp runs this source on the CPU
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c) { return mkp(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y)); }
__device__ __forceinline__ P2 mul2(P2 a, P2 b) { return mkp(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
__device__ __forceinline__ P2 add2(P2 a, P2 b) { return mkp(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ P2 add2Down(P2 a, P2 b) { return mkp(__fadd_rd(a.x, b.x), __fadd_rd(a.y, b.y)); }
#endif
#ifdef SDFB_QUERY_EXACT
__device__ __forceinline__ float monomialExact(float c, int i, int j, int k, float x, float y, float z) {
    float t = c;
    for (int a = 0; a < i; a++) t *= x;
    for (int a = 0; a < j; a++) t *= y;
    for (int a = 0; a < k; a++) t *= z;
    return t;
}
// interpolateValue (InterpolationMethods.h:432-439, scalar branch): acc = 0 + sum over n ascending of ((c_n * x^i) * y^j) * z^k,
// every product left to right. interpolateGradient (:442-455): per component the terms with a non-zero exponent in ascending
// n, the integer factor multiplies the coefficient first, no leading zero. The four sums are independent chains, so one pass
// over the coefficients feeds all of them without changing any sum's order; kVec: the coefficients arrive as 16 x 128-bit
// read-only loads instead of 64 (+ 3 x 48 with gradients) scalar ones — the one-load-per-term form sat on the L1 data pipe
// (a wavefront per distinct leaf and load), same bits.
// Value only: the four product chains of a coefficient vector (i = 0..3, same j and k) are multiplied in PAIRS by the packed
// float32 multiply of sm_100 (FMUL2: two IEEE products per issue slot) — (c0, c1 x) and (c2, c3 x) after one scalar product each,
// then x, x for the second pair and y^j, z^k for both — 160 multiply instructions instead of 288; the 64 additions stay one
// serial chain in the reference's order. Every product is the same IEEE operation on the same operands as in monomialExact.
template <bool kVec> __device__ __forceinline__ float evalLeafValuePacked(const float* c, float x, float y, float z) {
    const P2 xx = mkp(x, x), yy = mkp(y, y), zz = mkp(z, z);
    float acc = 0.0f;
#pragma unroll
    for (int m = 0; m < 16; m++) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(c) + m);
        const int j = m & 3, k = m >> 2;
        P2 lo = mkp(q.x, __fmul_rn(q.y, x));                           // i = 0, 1
        P2 hi = mul2(mul2(mkp(q.z, __fmul_rn(q.w, x)), xx), xx);      // i = 2, 3
#pragma unroll
        for (int a = 0; a < j; a++) { lo = mul2(lo, yy); hi = mul2(hi, yy); }
#pragma unroll
        for (int a = 0; a < k; a++) { lo = mul2(lo, zz); hi = mul2(hi, zz); }
        acc = __fadd_rn(acc, lo.x); acc = __fadd_rn(acc, lo.y); acc = __fadd_rn(acc, hi.x); acc = __fadd_rn(acc, hi.y);
    }
    return acc;
}

template <bool kGrad, bool kVec> __device__ __forceinline__ float evalLeaf(const float* c, float x, float y, float z, f3& g) {
    if (!kGrad && kVec) return evalLeafValuePacked<kVec>(c, x, y, z);
    float acc = 0.0f, gx = 0.0f, gy = 0.0f, gz = 0.0f;
#pragma unroll
    for (int m = 0; m < 16; m++) {
        float cc[4];
        if (kVec) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(c) + m);
            cc[0] = q.x; cc[1] = q.y; cc[2] = q.z; cc[3] = q.w;
        } else {
            cc[0] = __ldg(c + 4 * m); cc[1] = __ldg(c + 4 * m + 1); cc[2] = __ldg(c + 4 * m + 2); cc[3] = __ldg(c + 4 * m + 3);
        }
        const int j = m & 3, k = m >> 2;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int n = 4 * m + i;
            acc += monomialExact(cc[i], i, j, k, x, y, z);
            if (kGrad) {
                if (i > 0) { const float t = monomialExact(float(i) * cc[i], i - 1, j, k, x, y, z); gx = n == 1 ? t : gx + t; }
                if (j > 0) { const float t = monomialExact(float(j) * cc[i], i, j - 1, k, x, y, z); gy = n == 4 ? t : gy + t; }
                if (k > 0) { const float t = monomialExact(float(k) * cc[i], i, j, k - 1, x, y, z); gz = n == 16 ? t : gz + t; }
            }
        }
    }
    if (kGrad) g = normalize3(mk3(gx, gy, gz));
    return acc;
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

// OctreeSdf::getDistance at one point (src/sdf/OctreeSdf.cpp:93-152): start cell by the reference's float operations, the
// descent, the leaf polynomial (evalLeaf of this translation unit: FMA Horner, or the reference's literal order).
template <bool kGrad, bool kVec>
__device__ __forceinline__ float octreeDistanceAt(const uint32_t* __restrict__ oct, const QueryParams& q, f3 p, f3& g) {
    float fx = (p.x - q.minx) / q.cell, fy = (p.y - q.miny) / q.cell, fz = (p.z - q.minz) / q.cell;
    const float flx = floorf(fx), fly = floorf(fy), flz = floorf(fz);
    const int ix = int(flx), iy = int(fly), iz = int(flz);
    fx -= flx; fy -= fly; fz -= flz;
    g = mk3(0.0f, 0.0f, 0.0f);
    if (ix < 0 || ix >= q.grid || iy < 0 || iy >= q.grid || iz < 0 || iz >= q.grid)
        return (kGrad ? boxDistanceGrad(q, p, g) : boxDistance(q, p)) + q.minBorder;
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
    return evalLeaf<kGrad, kVec>(reinterpret_cast<const float*>(oct + (node & kOctIndexMask)), fx, fy, fz, g);
}

template <bool kGrad, bool kVec>
__global__ void __launch_bounds__(256)
octreeQueryKernel(const uint32_t* __restrict__ oct, const QueryParams q, const float* __restrict__ xyz, uint64_t n,
                  float* __restrict__ dist, float* __restrict__ grad) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f3 p = mk3(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2));
    f3 g;
    const float d = octreeDistanceAt<kGrad, kVec>(oct, q, p, g);
    dist[i] = d;
    if (kGrad) { grad[3 * i] = g.x; grad[3 * i + 1] = g.y; grad[3 * i + 2] = g.z; }
}

// ---- sphere tracing: the consumer of getDistance in the reference's viewer -----------------------------
// Reference: raycast() of src/render_engine/shaders/sdfOctreeRender.comp:392-410 —
//     while (last > eps && acc < far && it < maxIterations) { hit = pos; last = map(pos); step = max(last, 0); acc += step; pos += dir * step; it++ }
//     return last < eps;
// one ray per thread, map() = octreeDistanceAt above; the march is three multiplies and three adds per step in the
// shader's order, so the reference-order object reproduces a CPU loop over getDistance bit for bit.
struct TraceParams { float epsilon, farDistance; uint32_t maxIterations; };

template <bool kVec>
__global__ void __launch_bounds__(128)
octreeTraceKernel(const uint32_t* __restrict__ oct, const QueryParams q, const TraceParams tp, const float* __restrict__ origin,
                  const float* __restrict__ direction, uint64_t n, float* __restrict__ hitPos, float* __restrict__ travelled,
                  uint32_t* __restrict__ iterations) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f3 pos = mk3(__ldg(origin + 3 * i), __ldg(origin + 3 * i + 1), __ldg(origin + 3 * i + 2));
    const f3 dir = mk3(__ldg(direction + 3 * i), __ldg(direction + 3 * i + 1), __ldg(direction + 3 * i + 2));
    f3 hit = pos, g;
    float acc = 0.0f, last = 1e8f;
    uint32_t it = 0;
    while (last > tp.epsilon && acc < tp.farDistance && it < tp.maxIterations) {
        hit = pos;
        last = octreeDistanceAt<false, kVec>(oct, q, pos, g);
        const float step = gmax(last, 0.0f);
        acc = acc + step;
        pos = pos + dir * step;
        it++;
    }
    hitPos[3 * i] = hit.x; hitPos[3 * i + 1] = hit.y; hitPos[3 * i + 2] = hit.z;
    travelled[i] = last < tp.epsilon ? acc : -1.0f;   // -1: no surface within the limits
    if (iterations) iterations[i] = it;
}

// ---- the default kernel of the FMA path: warp-per-query-batch, top index, quad-cooperative evaluation -------------------------
// What bounds the one-query-per-thread kernel above (profiles/r2_summary.md, 256^3 workload): the LSU data pipe — every
// lane fills 64 coefficient registers, 16 x LDG.128 costing one wavefront per DISTINCT leaf in the warp — then the latency of
// its dependent chain (point -> 5-6 node gathers -> coefficients) and the XU pipe (IEEE division, floor, float <-> int). Once
// those are cut the kernel is ISSUE bound, so everything below is also chosen for its instruction count:
//   * every WARP is a persistent worker over 32-query tiles and loads the NEXT tile's points before it evaluates the
//     current one (their DRAM latency is off the chain). Staging the tiles through shared memory with one 1-D TMA bulk
//     copy per tile was built and measured: the mbarrier / elect / address bookkeeping costs ~75 instructions per tile
//     against 9 for three strided loads, and the kernel was no faster (0.235 ms vs 0.232 ms); it was dropped;
//   * a dense TOP INDEX (one word per cell of the grid `topLevels` below the start grid, <= 2^21 cells = 8 MB,
//     L2-resident) replaces the first `topLevels` dependent gathers by one load that is coalesced for coherent batches;
//     only leaves deeper than the index (0.2 % of the 256^3 queries) continue the descent;
//   * the cell selection uses no division and no conversion instruction but gives the reference's bits:
//     u = (p - min) / (cell / 2^L) — the reference's quotient scaled by the exact factor 2^L — is formed as two Newton
//     steps on x * RN(2^L / cell) with exact FMA residuals (Markstein: the correctly rounded quotient for normal
//     operands; tests/cpp/simt_query_main.cpp checks it against IEEE division), floor() is an add of 2^23 rounded
//     towards minus infinity whose mantissa is the integer cell, and "0 <= u < N" is one unsigned compare of the float
//     bits (negative values and NaN have larger patterns);
//   * the polynomial is evaluated by the four lanes of an aligned quad working on one (leaf, y, z) class at a time:
//     lane r loads the vectors c[0..3][j = r][k = 0..3] (64 contiguous bytes per quad and step), forms
//     y^r * Horner_z, two butterfly shuffles give every lane the same sums A_i(y, z), each member evaluates the cubic
//     in its own x. Grid rows put 4-8 neighbours of a row into one class; unrelated points take one round per lane
//     and still use every loaded byte.
// Same leaf and same leaf-local fractions as the reference (doubling, floor and subtraction are exact in binary floating
// point, so frac(2^k f) taken from u equals the reference's fract(2 f) chain); only the summation order of the polynomial
// differs (independent of where in the batch a query sits), within the tolerance of the FMA path.
#ifndef SDFB_QUERY_EXACT
struct TileQuery {
    float cellL;             // cell / 2^L (exact: a power-of-two scaling)
    float rcellL;            // RN(1 / cellL), computed on the host
    uint32_t limitBits;      // bit pattern of float(grid << L)
    int shiftN;              // log2(grid << L)
    int topLevels;           // L: levels below the start grid the top index resolves
    uint32_t G3;             // grid^3
};

// top index word: bit 31 = leaf, bits 27-30 = steps below the start grid at which the leaf / node was reached,
// bits 0-26 = (block - G^3) / 8 (every block of the array is 8 or 64 words long and starts after the G^3 start words)
constexpr uint32_t kTopLeaf = 1u << 31;
constexpr uint32_t kTopBlockMask = (1u << 27) - 1u;

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

// x / cell, correctly rounded, without a division: q0 = x * rc, two steps q <- q + (x - q * cell) * rc with exact
// residuals. Operands far below the normal range (|x| < 2^-100: the quotient is a denormal fraction of cell 0 either way)
// may differ from IEEE division in their last bits; no cell or leaf decision depends on them. A zero of either sign comes
// out as +0 (the residual of -0 is +0), which the unsigned range test relies on.
__device__ __forceinline__ float cellCoordinate(float x, float cell, float rc) {
    float qv = __fmul_rn(x, rc);
    qv = __fmaf_rn(__fmaf_rn(-qv, cell, x), rc, qv);
    return __fmaf_rn(__fmaf_rn(-qv, cell, x), rc, qv);
}
// v - floor(v) for 0 <= v < 2^22 (floor = v + 2^23 rounded towards minus infinity, minus 2^23; every step exact)
__device__ __forceinline__ float fractSmall(float v) {
    return __fadd_rn(v, -__fadd_rn(__fadd_rd(v, 8388608.0f), -8388608.0f));
}

// cellCoordinate / fractSmall for two coordinates at once (same operations per half)
__device__ __forceinline__ P2 cellCoordinate2(P2 x, P2 cell, P2 rc) {
    const P2 ncell = mkp(-cell.x, -cell.y);
    P2 qv = mul2(x, rc);
    qv = fma2(fma2(qv, ncell, x), rc, qv);      // fma(-q, cell, x) == fma(q, -cell, x): the product is exact either way
    return fma2(fma2(qv, ncell, x), rc, qv);
}
__device__ __forceinline__ P2 fractSmall2(P2 v) {
    const P2 big = mkp(8388608.0f, 8388608.0f), nbig = mkp(-8388608.0f, -8388608.0f);
    const P2 fl = add2(add2Down(v, big), nbig);
    return add2(v, mkp(-fl.x, -fl.y));
}

constexpr int kTileWarps = 8;   // warps per CTA of the tile kernel

// 6 resident CTAs per SM = 40 registers: measured faster than 8 x 32 registers (which spills 36 bytes): 0.162 vs 0.174 ms.
// Also measured and dropped: issuing the top-index load one tile ahead (a two-stage software pipeline, 48 registers,
// 5 CTAs per SM): 0.160 vs 0.163 ms — the kernel sits on a balance of issue slots (71 %), LSU wavefronts (61 %) and
// exposed L2 latency, and moving one of the three does not move the total.
template <bool kGrad, bool kPacked>   // kPacked: the value part and the x / y front end use the packed float32 instructions
__global__ void __launch_bounds__(kTileWarps * 32, 6)
octreeQueryTileKernel(const uint32_t* __restrict__ oct, const uint32_t* __restrict__ top, const QueryParams q, const TileQuery tq,
                      const float* __restrict__ xyz, uint64_t n, float* __restrict__ dist, float* __restrict__ grad) {
    constexpr unsigned kFull = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned r = lane & 3u, quadBase = lane & ~3u;
    const uint32_t tiles = uint32_t((n + 31) >> 5);   // the launcher keeps n below 2^36
    const uint32_t stride = gridDim.x * kTileWarps;
    uint32_t t = blockIdx.x * kTileWarps + (threadIdx.x >> 5);
    if (t >= tiles) return;
    uint64_t i = (uint64_t(t) << 5) + lane;
    f3 p = i < n ? mk3(__ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2)) : mk3(0.0f, 0.0f, 0.0f);
    for (;;) {
        const bool valid = i < n;
        const uint32_t tNext = t + stride;           // < 2^31 + 2^20: no wrap
        const uint64_t iNext = (uint64_t(tNext) << 5) + lane;
        f3 pNext = mk3(0.0f, 0.0f, 0.0f);
        if (tNext < tiles && iNext < n) pNext = mk3(__ldg(xyz + 3 * iNext), __ldg(xyz + 3 * iNext + 1), __ldg(xyz + 3 * iNext + 2));

        // u = (p - min) / cell * 2^L: integer part = cell of the top index, fraction = position inside it
        float ux, uy;
        if (kPacked) {
            const P2 u = cellCoordinate2(add2(mkp(p.x, p.y), mkp(-q.minx, -q.miny)), mkp(tq.cellL, tq.cellL), mkp(tq.rcellL, tq.rcellL));
            ux = u.x; uy = u.y;
        } else {
            ux = cellCoordinate(__fadd_rn(p.x, -q.minx), tq.cellL, tq.rcellL); uy = cellCoordinate(__fadd_rn(p.y, -q.miny), tq.cellL, tq.rcellL);
        }
        const float uz = cellCoordinate(__fadd_rn(p.z, -q.minz), tq.cellL, tq.rcellL);
        // 0 <= u < N as one unsigned compare of the bit pattern (negative values and NaN compare larger; -0 never occurs)
        const bool inside = valid && __float_as_uint(ux) < tq.limitBits && __float_as_uint(uy) < tq.limitBits && __float_as_uint(uz) < tq.limitBits;
        f3 g = mk3(0.0f, 0.0f, 0.0f);
        float d = 0.0f, fx = 0.0f, fy = 0.0f, fz = 0.0f;
        uint32_t block = 0;
        if (valid && !inside) d = (kGrad ? boxDistanceGrad(q, p, g) : boxDistance(q, p)) + q.minBorder;
        if (inside) {
            const uint32_t cx = uint32_t(__float_as_int(__fadd_rd(ux, 8388608.0f))) & 0x7FFFFFu, cy = uint32_t(__float_as_int(__fadd_rd(uy, 8388608.0f))) & 0x7FFFFFu,
                           cz = uint32_t(__float_as_int(__fadd_rd(uz, 8388608.0f))) & 0x7FFFFFu;
            const uint32_t e = __ldg(top + ((((cz << tq.shiftN) | cy) << tq.shiftN) | cx));
            block = ((e & kTopBlockMask) << 3) + tq.G3;
            int k = int((e >> 27) & 15u);
            if (!(e & kTopLeaf)) {
                // leaves below the index: the next child choices are the leading bits of the fraction inside the index cell
                const uint32_t bx = uint32_t(__float_as_int(__fadd_rd(__fmul_rn(fractSmall(ux), 65536.0f), 8388608.0f))) & 0xFFFFu,
                               by = uint32_t(__float_as_int(__fadd_rd(__fmul_rn(fractSmall(uy), 65536.0f), 8388608.0f))) & 0xFFFFu,
                               bz = uint32_t(__float_as_int(__fadd_rd(__fmul_rn(fractSmall(uz), 65536.0f), 8388608.0f))) & 0xFFFFu;
                for (int sh = 15;; sh--) {
                    const uint32_t child = ((bx >> sh) & 1u) | (((by >> sh) & 1u) << 1) | (((bz >> sh) & 1u) << 2);
                    const uint32_t node = __ldg(oct + block + child);
                    k++;
                    block = node & kOctIndexMask;
                    if (node & kLeafBit) break;
                }
            }
            const float scale = __int_as_float((127 + k - tq.topLevels) << 23);   // 2^(k - L): leaf-local fraction = frac(f * 2^k)
            if (kPacked) {
                const P2 fr = fractSmall2(mul2(mkp(ux, uy), mkp(scale, scale)));
                fx = fr.x; fy = fr.y;
            } else {
                fx = fractSmall(__fmul_rn(ux, scale)); fy = fractSmall(__fmul_rn(uy, scale));
            }
            fz = fractSmall(__fmul_rn(uz, scale));
        }
        bool pending = inside;
        for (;;) {
            const unsigned pend = __ballot_sync(kFull, pending);
            if (pend == 0) break;                                   // warp-uniform
            const unsigned mine = (pend >> quadBase) & 0xFu;
            const int leader = int(quadBase) + (mine ? __ffs(int(mine)) - 1 : 0);   // a finished quad idles through the shuffles
            const uint32_t lb = __shfl_sync(kFull, block, leader);
            const float ly = __shfl_sync(kFull, fy, leader), lz = __shfl_sync(kFull, fz, leader);
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;       // this lane's part of A_i
            float y0 = 0.0f, y1 = 0.0f, y2 = 0.0f, y3 = 0.0f;       // ... of dA_i/dy
            float z0 = 0.0f, z1 = 0.0f, z2 = 0.0f, z3 = 0.0f;       // ... of dA_i/dz
            if (mine) {
                const float4* c = reinterpret_cast<const float4*>(oct + lb) + r;   // vector m = r + 4 k holds c[0..3][j = r][k]
                const float4 c0 = __ldg(c), c1 = __ldg(c + 4), c2 = __ldg(c + 8), c3 = __ldg(c + 12);
                const float yy = ly * ly;
                const float yr = ((r & 1u) ? ly : 1.0f) * ((r & 2u) ? yy : 1.0f);                     // y^r
                float h0, h1, h2, h3;
                if (kPacked) {
                    const P2 z2 = mkp(lz, lz), y2 = mkp(yr, yr);
                    const P2 h01 = fma2(fma2(fma2(mkp(c3.x, c3.y), z2, mkp(c2.x, c2.y)), z2, mkp(c1.x, c1.y)), z2, mkp(c0.x, c0.y));
                    const P2 h23 = fma2(fma2(fma2(mkp(c3.z, c3.w), z2, mkp(c2.z, c2.w)), z2, mkp(c1.z, c1.w)), z2, mkp(c0.z, c0.w));
                    const P2 a01 = mul2(y2, h01), a23 = mul2(y2, h23);
                    h0 = h01.x; h1 = h01.y; h2 = h23.x; h3 = h23.y;
                    a0 = a01.x; a1 = a01.y; a2 = a23.x; a3 = a23.y;
                } else {
                    h0 = fmaf(fmaf(fmaf(c3.x, lz, c2.x), lz, c1.x), lz, c0.x); h1 = fmaf(fmaf(fmaf(c3.y, lz, c2.y), lz, c1.y), lz, c0.y);
                    h2 = fmaf(fmaf(fmaf(c3.z, lz, c2.z), lz, c1.z), lz, c0.z); h3 = fmaf(fmaf(fmaf(c3.w, lz, c2.w), lz, c1.w), lz, c0.w);
                    a0 = yr * h0; a1 = yr * h1; a2 = yr * h2; a3 = yr * h3;
                }
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
                if (kPacked) {
                    const P2 s01 = add2(mkp(a0, a1), mkp(__shfl_xor_sync(kFull, a0, m), __shfl_xor_sync(kFull, a1, m)));
                    const P2 s23 = add2(mkp(a2, a3), mkp(__shfl_xor_sync(kFull, a2, m), __shfl_xor_sync(kFull, a3, m)));
                    a0 = s01.x; a1 = s01.y; a2 = s23.x; a3 = s23.y;
                } else {
                    a0 += __shfl_xor_sync(kFull, a0, m); a1 += __shfl_xor_sync(kFull, a1, m);
                    a2 += __shfl_xor_sync(kFull, a2, m); a3 += __shfl_xor_sync(kFull, a3, m);
                }
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
        if (valid) {
            dist[i] = d;
            if (kGrad) { grad[3 * i] = g.x; grad[3 * i + 1] = g.y; grad[3 * i + 2] = g.z; }
        }
        if (tNext >= tiles) break;
        t = tNext; i = iNext; p = pNext;
    }
}

#endif
