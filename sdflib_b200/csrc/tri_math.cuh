// Device math shared by the builders and the exact-octree query: float3 algebra in glm's operation
// order, the point-triangle kernels and the float64 closest-point routine of the BVH.
//
// EVERY translation unit that includes this header for topology-deciding work is compiled with
// -fmad=false (see build.py): the reference is built for x86-64 without FMA, so a fused a*b+c here
// would change the last bit of distances and, through `error < threshold^2` and the Frank-Wolfe
// exits, the octree topology. sqrtf / division stay IEEE (-prec-sqrt/-prec-div default to true).
//
// Reference semantics restated (paths relative to the reference tree):
//   TriangleUtils::getSqDistPointAndTriangle            include/SdfLib/utils/TriangleUtils.h:76-135
//   TriangleUtils::getSignedDistPointAndTriangle (x3)   include/SdfLib/utils/TriangleUtils.h:137-376
//   tmd::point_triangle_sq_unsigned (Eberly)            libs/InteractiveComputerGraphics/.../TriangleMeshDistance.h:542-798
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace sdfb200 {

struct f3 { float x, y, z; };
__host__ __device__ __forceinline__ f3 mk3(float a, float b, float c) { f3 r; r.x = a; r.y = b; r.z = c; return r; }
__host__ __device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__host__ __device__ __forceinline__ f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
__host__ __device__ __forceinline__ float dot3(f3 a, f3 b) {
    const float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
    return tx + ty + tz;
}
__host__ __device__ __forceinline__ f3 normalize3(f3 v) { return v * (1.0f / sqrtf(dot3(v, v))); }
__host__ __device__ __forceinline__ float sign1(float x) { return float((0.0f < x) - (x < 0.0f)); }
__host__ __device__ __forceinline__ float gmin(float a, float b) { return (b < a) ? b : a; }  // glm::min
__host__ __device__ __forceinline__ float gmax(float a, float b) { return (a < b) ? b : a; }  // glm::max
__host__ __device__ __forceinline__ float gabs(float a) { return a >= 0.0f ? a : -a; }       // glm::abs
// corner / child direction: index c = x | y<<1 | z<<2, components are -1 / +1
__host__ __device__ __forceinline__ f3 cornerDir(uint32_t c) {
    return mk3((c & 1u) ? 1.0f : -1.0f, (c & 2u) ? 1.0f : -1.0f, (c & 4u) ? 1.0f : -1.0f);
}

// TriangleData, 37 floats, same field order as the reference serialises (TriangleUtils.h:53):
// origin(3) transform(9, column-major) b(2) c(2) v2(1) v3(2) edgesNormal(9) verticesNormal(9)
struct TriData {
    float origin[3];
    float T[3][3];
    float b[2], c[2];
    float v2;
    float v3[2];
    float edgesNormal[3][3];
    float verticesNormal[3][3];
};
static_assert(sizeof(TriData) == 148, "TriangleData is 37 floats");

__host__ __device__ __forceinline__ f3 matMul(const float T[3][3], f3 v) {
    return mk3(T[0][0] * v.x + T[1][0] * v.y + T[2][0] * v.z,
               T[0][1] * v.x + T[1][1] * v.y + T[2][1] * v.z,
               T[0][2] * v.x + T[1][2] * v.y + T[2][2] * v.z);
}
__host__ __device__ __forceinline__ f3 matTMul(const float T[3][3], f3 v) {
    return mk3(T[0][0] * v.x + T[0][1] * v.y + T[0][2] * v.z,
               T[1][0] * v.x + T[1][1] * v.y + T[1][2] * v.z,
               T[2][0] * v.x + T[2][1] * v.y + T[2][2] * v.z);
}
__host__ __device__ __forceinline__ f3 triNormal(const TriData& d) { return mk3(d.T[0][2], d.T[1][2], d.T[2][2]); }
__host__ __device__ __forceinline__ f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }

// ---- region classification in the triangle frame -----------------------------------------------
enum : int { REG_V1 = 0, REG_V2 = 1, REG_V3 = 2, REG_E1 = 3, REG_E2 = 4, REG_E3 = 5, REG_FACE = 6 };

struct TriFrame {   // the 19-float "distance part" of TriangleData, in registers
    float ox, oy, oz;
    float t00, t01, t02, t10, t11, t12, t20, t21, t22;   // tCR = T[C][R]
    float bx, by, cx, cy, v2, v3x, v3y;
};

__host__ __device__ __forceinline__ TriFrame frameOf(const TriData& d) {
    TriFrame f;
    f.ox = d.origin[0]; f.oy = d.origin[1]; f.oz = d.origin[2];
    f.t00 = d.T[0][0]; f.t01 = d.T[0][1]; f.t02 = d.T[0][2];
    f.t10 = d.T[1][0]; f.t11 = d.T[1][1]; f.t12 = d.T[1][2];
    f.t20 = d.T[2][0]; f.t21 = d.T[2][1]; f.t22 = d.T[2][2];
    f.bx = d.b[0]; f.by = d.b[1]; f.cx = d.c[0]; f.cy = d.c[1];
    f.v2 = d.v2; f.v3x = d.v3[0]; f.v3y = d.v3[1];
    return f;
}

struct Classified { f3 q; int region; float de; };

__host__ __device__ __forceinline__ Classified classify(f3 p, const TriFrame& f) {
    Classified o;
    const float dx = p.x - f.ox, dy = p.y - f.oy, dz = p.z - f.oz;
    o.q = mk3(f.t00 * dx + f.t10 * dy + f.t20 * dz, f.t01 * dx + f.t11 * dy + f.t21 * dz,
              f.t02 * dx + f.t12 * dy + f.t22 * dz);
    const f3 q = o.q;
    const float de1 = -q.y;
    const float de2 = (q.x - f.v2) * f.by - q.y * f.bx;
    const float de3 = q.x * f.cy - q.y * f.cx;
    o.de = 0.0f;
    if (de1 >= 0) {
        if (q.x <= 0) o.region = REG_V1;
        else if (q.x >= f.v2) o.region = REG_V2;
        else { o.region = REG_E1; o.de = de1; }
    } else if (de2 >= 0) {
        if ((q.x - f.v2) * f.bx + q.y * f.by <= 0) o.region = REG_V2;
        else if ((q.x - f.v3x) * f.bx + (q.y - f.v3y) * f.by >= 0) o.region = REG_V3;
        else { o.region = REG_E2; o.de = de2; }
    } else if (de3 >= 0) {
        if (q.x * f.cx + q.y * f.cy >= 0) o.region = REG_V1;
        else if ((q.x - f.v3x) * f.cx + (q.y - f.v3y) * f.cy <= 0) o.region = REG_V3;
        else { o.region = REG_E3; o.de = de3; }
    } else o.region = REG_FACE;
    return o;
}

__host__ __device__ __forceinline__ f3 localVector(const Classified& c, const TriFrame& f) {
    if (c.region == REG_V2) return c.q - mk3(f.v2, 0.0f, 0.0f);
    if (c.region == REG_V3) return c.q - mk3(f.v3x, f.v3y, 0.0f);
    return c.q;
}

__host__ __device__ __forceinline__ float sqDistOf(const Classified& c, const TriFrame& f) {
    if (c.region == REG_FACE) return c.q.z * c.q.z;
    if (c.region >= REG_E1) return c.de * c.de + c.q.z * c.q.z;
    const f3 l = localVector(c, f);
    return dot3(l, l);
}

// a3: squared distance (TriangleUtils.h:76-135)
__host__ __device__ __forceinline__ float sqDistPointTriangle(f3 p, const TriFrame& f) {
    return sqDistOf(classify(p, f), f);
}

__host__ __device__ __forceinline__ float regionSign(const Classified& c, const TriFrame& f, const TriData& d) {
    switch (c.region) {
        case REG_V1: return sign1(dot3(ld3(d.verticesNormal[0]), c.q));
        case REG_V2: return sign1(dot3(ld3(d.verticesNormal[1]), localVector(c, f)));
        case REG_V3: return sign1(dot3(ld3(d.verticesNormal[2]), localVector(c, f)));
        case REG_E1: return sign1(dot3(ld3(d.edgesNormal[0]), c.q));
        case REG_E2: return sign1(dot3(ld3(d.edgesNormal[1]), c.q - mk3(f.v2, 0.0f, 0.0f)));
        case REG_E3: return sign1(dot3(ld3(d.edgesNormal[2]), c.q));
        default: return 1.0f;
    }
}

__host__ __device__ __forceinline__ f3 edgePerpendicular(const Classified& c, const TriFrame& f) {
    const f3 q = c.q;
    if (c.region == REG_E1) return mk3(0.0f, q.y, q.z);
    if (c.region == REG_E2) {
        const float t = (q.x - f.v2) * f.bx + q.y * f.by;
        return mk3((q.x - f.v2) - t * f.bx, q.y - t * f.by, q.z);
    }
    const float t = q.x * f.cx + q.y * f.cy;
    return mk3(q.x - t * f.cx, q.y - t * f.cy, q.z);
}

// a4 (:137-196): signed distance only
__host__ __device__ __forceinline__ float signedDistPointTriangle(f3 p, const TriData& d) {
    const TriFrame f = frameOf(d);
    const Classified c = classify(p, f);
    if (c.region == REG_FACE) return c.q.z;
    return regionSign(c, f, d) * sqrtf(sqDistOf(c, f));
}

// a4 (:198-290): signed distance + unit gradient, vertex regions use the world-space vertices,
// NaN from normalize(0) falls back to the triangle normal. Used by the tri-cubic fit.
__host__ __device__ __forceinline__ float signedDistGradMesh(f3 p, const TriData& d, f3 w1, f3 w2, f3 w3, f3& outN) {
    const TriFrame f = frameOf(d);
    const Classified c = classify(p, f);
    if (c.region == REG_FACE) { outN = triNormal(d); return c.q.z; }
    const float s = regionSign(c, f, d);
    f3 v;
    if (c.region == REG_V1) v = p - w1;
    else if (c.region == REG_V2) v = p - w2;
    else if (c.region == REG_V3) v = p - w3;
    else v = matTMul(d.T, edgePerpendicular(c, f));
    f3 n = normalize3(v);
    if (isnan(n.x + n.y + n.z)) n = triNormal(d);
    outN = s * n;
    return s * sqrtf(sqDistOf(c, f));
}

// a4 (:292-376): self-contained gradient variant (exact-octree query); no NaN guard.
__host__ __device__ __forceinline__ float signedDistGradSelf(f3 p, const TriData& d, f3& outN) {
    const TriFrame f = frameOf(d);
    const Classified c = classify(p, f);
    if (c.region == REG_FACE) { outN = triNormal(d); return c.q.z; }
    const float s = regionSign(c, f, d);
    const f3 o = ld3(d.origin);
    f3 v;
    if (c.region == REG_V1) v = p - o;
    else if (c.region == REG_V2) v = p - o - matTMul(d.T, mk3(f.v2, 0.0f, 0.0f));
    else if (c.region == REG_V3) v = p - o - matTMul(d.T, mk3(f.v3x, f.v3y, 0.0f));
    else v = matTMul(d.T, edgePerpendicular(c, f));
    outN = s * normalize3(v);
    return s * sqrtf(sqDistOf(c, f));
}

// ---- float64 closest point on a triangle (squared distance only) --------------------------------
struct d3 { double x, y, z; };
__host__ __device__ __forceinline__ d3 mkd(double a, double b, double c) { d3 r; r.x = a; r.y = b; r.z = c; return r; }
__host__ __device__ __forceinline__ d3 operator-(d3 a, d3 b) { return mkd(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ double ddot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

__host__ __device__ inline double eberlySqDist(d3 p, d3 v0, d3 v1, d3 v2) {
    const d3 diff = v0 - p, e0 = v1 - v0, e1 = v2 - v0;
    const double a00 = ddot(e0, e0), a01 = ddot(e0, e1), a11 = ddot(e1, e1);
    const double b0 = ddot(diff, e0), b1 = ddot(diff, e1), c = ddot(diff, diff);
    const double det = fabs(a00 * a11 - a01 * a01);
    double s = a01 * b1 - a11 * b0;
    double t = a01 * b0 - a00 * b1;
    double d2;
#define SDFB_VERTEX1 (a00 + 2 * b0 + c)
#define SDFB_VERTEX2 (a11 + 2 * b1 + c)
#define SDFB_EDGE0 (b0 * (-b0 / a00) + c)
#define SDFB_EDGE1 (b1 * (-b1 / a11) + c)
#define SDFB_QUAD(ss, tt) ((ss) * (a00 * (ss) + a01 * (tt) + 2 * b0) + (tt) * (a01 * (ss) + a11 * (tt) + 2 * b1) + c)
    if (s + t <= det) {
        if (s < 0) {
            if (t < 0 && b0 < 0) d2 = (-b0 >= a00) ? SDFB_VERTEX1 : SDFB_EDGE0;   // region 4, edge 0 side
            else d2 = (b1 >= 0) ? c : ((-b1 >= a11) ? SDFB_VERTEX2 : SDFB_EDGE1);  // region 3 / 4
        } else if (t < 0) {
            d2 = (b0 >= 0) ? c : ((-b0 >= a00) ? SDFB_VERTEX1 : SDFB_EDGE0);        // region 5
        } else {                                                                   // region 0
            const double inv = 1 / det;
            s *= inv; t *= inv;
            d2 = SDFB_QUAD(s, t);
        }
    } else {
        if (s < 0) {   // region 2
            const double tmp0 = a01 + b0, tmp1 = a11 + b1;
            if (tmp1 > tmp0) {
                const double numer = tmp1 - tmp0, denom = a00 - 2 * a01 + a11;
                if (numer >= denom) d2 = SDFB_VERTEX1;
                else { s = numer / denom; t = 1 - s; d2 = SDFB_QUAD(s, t); }
            } else d2 = (tmp1 <= 0) ? SDFB_VERTEX2 : ((b1 >= 0) ? c : SDFB_EDGE1);
        } else if (t < 0) {   // region 6
            const double tmp0 = a01 + b1, tmp1 = a00 + b0;
            if (tmp1 > tmp0) {
                const double numer = tmp1 - tmp0, denom = a00 - 2 * a01 + a11;
                if (numer >= denom) d2 = SDFB_VERTEX2;
                else { t = numer / denom; s = 1 - t; d2 = SDFB_QUAD(s, t); }
            } else d2 = (tmp1 <= 0) ? SDFB_VERTEX1 : ((b0 >= 0) ? c : SDFB_EDGE0);
        } else {   // region 1
            const double numer = a11 + b1 - a01 - b0;
            if (numer <= 0) d2 = SDFB_VERTEX2;
            else {
                const double denom = a00 - 2 * a01 + a11;
                if (numer >= denom) d2 = SDFB_VERTEX1;
                else { s = numer / denom; t = 1 - s; d2 = SDFB_QUAD(s, t); }
            }
        }
    }
#undef SDFB_VERTEX1
#undef SDFB_VERTEX2
#undef SDFB_EDGE0
#undef SDFB_EDGE1
#undef SDFB_QUAD
    return d2 < 0 ? 0 : d2;
}

}  // namespace sdfb200
