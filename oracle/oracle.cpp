// TEST INFRASTRUCTURE — CPU restatement ("port") of the SdfLib hot paths. Not product code.
// See oracle.h for the role of this file. All citations are relative to /root/reference.
//
// Arithmetic contract: IEEE float32 (float64 inside the BVH), no FMA contraction (build with
// -ffp-contract=off; x86-64 baseline has no FMA anyway), and the operation ORDER of glm 0.9.8's
// scalar formulas (dot = products then left-to-right adds; normalize = v * (1/sqrt(dot));
// mat3*vec3 column-major). Those orders are what make the result bit-identical to the reference.
#include "oracle.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <vector>

namespace {

// ------------------------------------------------------------------------------------------
// small float3 algebra with glm's operation order
// ------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
inline V3 v3(float a, float b, float c) { return V3{a, b, c}; }
inline V3 v3(float s) { return V3{s, s, s}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
inline V3 operator+(V3 a, float s) { return V3{a.x + s, a.y + s, a.z + s}; }
inline V3 operator-(V3 a, float s) { return V3{a.x - s, a.y - s, a.z - s}; }
inline float dot(V3 a, V3 b) { float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }
inline V3 cross(V3 a, V3 b) { return V3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline V3 normalize(V3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }
inline float length(V3 v) { return std::sqrt(dot(v, v)); }
inline float fmin_(float a, float b) { return (b < a) ? b : a; }   // glm::min
inline float fmax_(float a, float b) { return (a < b) ? b : a; }   // glm::max
inline float fabs_(float a) { return a >= 0.0f ? a : -a; }         // glm::abs
inline float fsign(float x) { return float((0.0f < x) - (x < 0.0f)); }
inline float fract_(float x) { return x - std::floor(x); }

// Corner / child offsets: index c = x | y<<1 | z<<2 (include/SdfLib/TrianglesInfluence.h:25-36)
inline V3 cornerDir(uint32_t c) { return V3{(c & 1) ? 1.0f : -1.0f, (c & 2) ? 1.0f : -1.0f, (c & 4) ? 1.0f : -1.0f}; }

// The 19 mid-points of a node: lattice points (x,y,z) in {0,1,2}^3 that are not corners, in
// increasing x+3y+9z (src/sdf/OctreeSdfDepthFirst.h:139-162).
struct Lattice {
    int sampleOfLattice[27];   // -1 for corners
    int cornerOfLattice[27];   // -1 for mid-points
    V3 sampleRel[19];
    Lattice() {
        int s = 0;
        for (int L = 0; L < 27; L++) {
            int x = L % 3, y = (L / 3) % 3, z = L / 9;
            bool corner = (x != 1) && (y != 1) && (z != 1);
            cornerOfLattice[L] = corner ? ((x >> 1) | ((y >> 1) << 1) | ((z >> 1) << 2)) : -1;
            sampleOfLattice[L] = corner ? -1 : s;
            if (!corner) sampleRel[s++] = V3{float(x - 1), float(y - 1), float(z - 1)};
        }
    }
};
const Lattice kLattice;

// ------------------------------------------------------------------------------------------
// a1: TriangleData (include/SdfLib/utils/TriangleUtils.h:20-72). 37 floats, serialised in this
// field order (:53). T is column-major: T[c][r].
// ------------------------------------------------------------------------------------------
struct TriData {
    V3 origin;
    float T[3][3];
    float b[2], c[2];
    float v2;
    float v3[2];
    V3 edgesNormal[3];
    V3 verticesNormal[3];
};
static_assert(sizeof(TriData) == 37 * 4, "TriData must be 148 bytes");

inline V3 matMul(const float T[3][3], V3 v) {   // glm mat3 * vec3
    return V3{T[0][0] * v.x + T[1][0] * v.y + T[2][0] * v.z,
              T[0][1] * v.x + T[1][1] * v.y + T[2][1] * v.z,
              T[0][2] * v.x + T[1][2] * v.y + T[2][2] * v.z};
}
inline V3 matTMul(const float T[3][3], V3 v) {  // glm::transpose(T) * vec3
    return V3{T[0][0] * v.x + T[0][1] * v.y + T[0][2] * v.z,
              T[1][0] * v.x + T[1][1] * v.y + T[1][2] * v.z,
              T[2][0] * v.x + T[2][1] * v.y + T[2][2] * v.z};
}
inline V3 triNormal(const TriData& d) { return V3{d.T[0][2], d.T[1][2], d.T[2][2]}; }  // TriangleUtils.h:45-48

// TriangleData ctor, TriangleUtils.h:23-42
TriData makeTriData(V3 p1, V3 p2, V3 p3) {
    TriData d;
    d.origin = p1;
    V3 sx = normalize(p2 - p1);
    V3 sz = normalize(cross(p2 - p1, p3 - p1));
    V3 sy = cross(sz, sx);
    // glm::inverse(mat3(sx, sy, sz)): columns m[0]=sx, m[1]=sy, m[2]=sz
    const float m[3][3] = {{sx.x, sx.y, sx.z}, {sy.x, sy.y, sy.z}, {sz.x, sz.y, sz.z}};
    float inv = 1.0f / (+m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
                        - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
                        + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
    d.T[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * inv;
    d.T[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * inv;
    d.T[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * inv;
    d.T[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * inv;
    d.T[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * inv;
    d.T[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * inv;
    d.T[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * inv;
    d.T[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * inv;
    d.T[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * inv;
    auto norm2 = [](float x, float y, float* out) {
        float tx = x * x, ty = y * y;
        float s = 1.0f / std::sqrt(tx + ty);
        out[0] = x * s; out[1] = y * s;
    };
    V3 e = matMul(d.T, p3 - p2);
    norm2(e.x, e.y, d.b);
    e = matMul(d.T, p1 - p3);
    norm2(e.x, e.y, d.c);
    d.v2 = matMul(d.T, p2 - d.origin).x;
    e = matMul(d.T, p3 - d.origin);
    d.v3[0] = e.x; d.v3[1] = e.y;
    for (int k = 0; k < 3; k++) { d.edgesNormal[k] = v3(0.f, 0.f, 1.f); d.verticesNormal[k] = v3(0.f, 0.f, 1.f); }
    return d;
}

struct MeshView {
    const V3* verts; uint32_t nVerts;
    const uint32_t* idx; uint32_t nIdx;
};

// a2: calculateMeshTriangleData (src/utils/TriangleUtils.cpp:7-428). The degenerate-triangle branch
// is compiled out in the reference (`if(false && ...)`, :45), so :90-290 never run; what is live is
// the frame construction, the edge pairing through an ordered map (:63-83), the angle-weighted
// vertex normals in triangle order (:85-86), the non-manifold vertex merge (:292-420) and the final
// transform of vertex normals (:422-425).
std::vector<TriData> meshTriangleData(const MeshView& m) {
    const uint32_t nT = m.nIdx / 3;
    std::vector<TriData> tris(nT);
    for (uint32_t t = 0; t < nT; t++)
        tris[t] = makeTriData(m.verts[m.idx[3 * t]], m.verts[m.idx[3 * t + 1]], m.verts[m.idx[3 * t + 2]]);

    std::map<std::pair<uint32_t, uint32_t>, uint32_t> openEdges;
    std::vector<V3> vNormal(m.nVerts, v3(0.0f));
    for (uint32_t t = 0; t < nT; t++) {
        for (uint32_t k = 0; k < 3; k++) {
            const uint32_t a = m.idx[3 * t + k], b = m.idx[3 * t + (k + 1) % 3], c = m.idx[3 * t + (k + 2) % 3];
            auto key = std::make_pair(std::min(a, b), std::max(a, b));
            auto ins = openEdges.insert(std::make_pair(key, 3 * t + k));
            if (!ins.second) {
                const uint32_t other = ins.first->second, t2 = other / 3;
                V3 n = triNormal(tris[t]) + triNormal(tris[t2]);
                tris[t].edgesNormal[k] = matMul(tris[t].T, n);
                tris[t2].edgesNormal[other % 3] = matMul(tris[t2].T, n);
                openEdges.erase(ins.first);
            }
            float cosA = dot(normalize(m.verts[b] - m.verts[a]), normalize(m.verts[c] - m.verts[a]));
            const float angle = std::acos(fmin_(fmax_(cosA, -1.0f), 1.0f));
            V3 add = angle * triNormal(tris[t]);
            vNormal[a] = vNormal[a] + add;
        }
    }

    if (!openEdges.empty()) {   // :292-420, merge near-coincident vertices on a 2048^3 hash grid
        std::map<uint32_t, uint32_t> parentOf;
        auto findParent = [&](uint32_t v) {
            auto it = parentOf.find(v);
            while (it != parentOf.end() && it->second != v) { v = it->second; it = parentOf.find(v); }
            return v;
        };
        std::vector<uint32_t> nm;
        for (auto& e : openEdges) { nm.push_back(e.first.first); nm.push_back(e.first.second); }
        std::sort(nm.begin(), nm.end());
        nm.erase(std::unique(nm.begin(), nm.end()), nm.end());

        V3 mn = v3(INFINITY), mx = v3(-INFINITY);   // Mesh::computeBoundingBox, src/utils/Mesh.cpp:89-105
        for (uint32_t i = 0; i < m.nVerts; i++) {
            mn = V3{fmin_(mn.x, m.verts[i].x), fmin_(mn.y, m.verts[i].y), fmin_(mn.z, m.verts[i].z)};
            mx = V3{fmax_(mx.x, m.verts[i].x), fmax_(mx.y, m.verts[i].y), fmax_(mx.z, m.verts[i].z)};
        }
        const V3 bb = mx - mn;
        const uint32_t axisRes = 2048;
        const float maxExt = fmax_(bb.x, fmax_(bb.y, bb.z));
        const float gridScale = float(axisRes) / maxExt;
        const float threshold = float(1e-5 / double(maxExt));
        const float sqThreshold = threshold * threshold;
        std::map<uint64_t, std::vector<uint32_t>> set1, set2;
        auto cellId = [axisRes](int x, int y, int z) { return uint64_t(uint32_t(x + y * axisRes + z * axisRes * axisRes)); };
        auto cellOf = [&](V3 p, float off) {
            V3 q = (p - mn) * gridScale + off;
            return cellId(int(q.x), int(q.y), int(q.z));
        };
        for (uint32_t v : nm) {
            set1[cellOf(m.verts[v], 0.0f)].push_back(v);
            set2[cellOf(m.verts[v], 0.5f)].push_back(v);
        }
        std::map<uint64_t, std::vector<uint32_t>>* sets[2] = {&set1, &set2};
        for (uint32_t v : nm) {
            float off = 0.0f;
            for (auto* s : sets) {
                auto it = s->find(cellOf(m.verts[v], off));
                if (it != s->end()) {
                    for (uint32_t u : it->second) {
                        V3 diff = m.verts[v] - m.verts[u];
                        if (dot(diff, diff) < sqThreshold) {
                            uint32_t p1 = findParent(v), p2 = findParent(u);
                            if (v == p1) parentOf[p1] = p1;
                            parentOf[p2] = p1;
                            break;
                        }
                    }
                }
                off += 0.5f;
            }
        }
        std::map<std::pair<uint32_t, uint32_t>, uint32_t> merged;
        for (auto& e : openEdges) {
            uint32_t a = findParent(e.first.first), b = findParent(e.first.second);
            auto ins = merged.insert(std::make_pair(std::make_pair(std::min(a, b), std::max(a, b)), e.second));
            if (!ins.second) {
                const uint32_t t = e.second / 3, t2 = ins.first->second / 3;
                V3 n = triNormal(tris[t]) + triNormal(tris[t2]);
                tris[t].edgesNormal[e.second % 3] = matMul(tris[t].T, n);
                tris[t2].edgesNormal[ins.first->second % 3] = matMul(tris[t2].T, n);
                merged.erase(ins.first);
            }
        }
        for (uint32_t v : nm) { uint32_t p = findParent(v); if (p != v) vNormal[p] = vNormal[p] + vNormal[v]; }
        for (uint32_t v : nm) vNormal[v] = vNormal[findParent(v)];
    }

    for (uint32_t i = 0; i < m.nIdx; i++) tris[i / 3].verticesNormal[i % 3] = matMul(tris[i / 3].T, vNormal[m.idx[i]]);
    return tris;
}

// ------------------------------------------------------------------------------------------
// a3 / a4: point–triangle kernels (TriangleUtils.h:76-376). One classifier shared by all variants;
// each region evaluates exactly the float expressions of the reference's branch.
// ------------------------------------------------------------------------------------------
enum Region { R_V1, R_V2, R_V3, R_E1, R_E2, R_E3, R_FACE };

struct Classified {
    V3 q;          // point in the triangle frame
    Region r;
    float de;      // signed edge distance for edge regions
};

inline Classified classify(V3 p, const TriData& d) {
    Classified o;
    o.q = matMul(d.T, p - d.origin);
    const V3 q = o.q;
    const float de1 = -q.y;
    const float de2 = (q.x - d.v2) * d.b[1] - q.y * d.b[0];
    const float de3 = q.x * d.c[1] - q.y * d.c[0];
    o.de = 0.0f;
    if (de1 >= 0) {
        if (q.x <= 0) o.r = R_V1;
        else if (q.x >= d.v2) o.r = R_V2;
        else { o.r = R_E1; o.de = de1; }
    } else if (de2 >= 0) {
        if ((q.x - d.v2) * d.b[0] + q.y * d.b[1] <= 0) o.r = R_V2;
        else if ((q.x - d.v3[0]) * d.b[0] + (q.y - d.v3[1]) * d.b[1] >= 0) o.r = R_V3;
        else { o.r = R_E2; o.de = de2; }
    } else if (de3 >= 0) {
        if (q.x * d.c[0] + q.y * d.c[1] >= 0) o.r = R_V1;
        else if ((q.x - d.v3[0]) * d.c[0] + (q.y - d.v3[1]) * d.c[1] <= 0) o.r = R_V3;
        else { o.r = R_E3; o.de = de3; }
    } else o.r = R_FACE;
    return o;
}

inline V3 localVector(const Classified& c, const TriData& d) {   // q relative to the region's vertex
    switch (c.r) {
        case R_V2: return c.q - v3(d.v2, 0.0f, 0.0f);
        case R_V3: return c.q - v3(d.v3[0], d.v3[1], 0.0f);
        default: return c.q;
    }
}

inline float sqDistOf(const Classified& c, const TriData& d) {
    switch (c.r) {
        case R_V1: return dot(c.q, c.q);
        case R_V2: case R_V3: { V3 l = localVector(c, d); return dot(l, l); }
        case R_E1: case R_E2: case R_E3: return c.de * c.de + c.q.z * c.q.z;
        default: return c.q.z * c.q.z;
    }
}

float sqDistPointTriangle(V3 p, const TriData& d) { return sqDistOf(classify(p, d), d); }   // :76-135

inline float regionSign(const Classified& c, const TriData& d) {   // sign of pseudo-normal · local vector
    switch (c.r) {
        case R_V1: return fsign(dot(d.verticesNormal[0], c.q));
        case R_V2: return fsign(dot(d.verticesNormal[1], localVector(c, d)));
        case R_V3: return fsign(dot(d.verticesNormal[2], localVector(c, d)));
        case R_E1: return fsign(dot(d.edgesNormal[0], c.q));
        case R_E2: return fsign(dot(d.edgesNormal[1], c.q - v3(d.v2, 0.0f, 0.0f)));   // edge 2 is relative to v2 (:175)
        case R_E3: return fsign(dot(d.edgesNormal[2], c.q));
        default: return 1.0f;
    }
}

float signedDistPointTriangle(V3 p, const TriData& d) {   // :137-196
    Classified c = classify(p, d);
    if (c.r == R_FACE) return c.q.z;
    return regionSign(c, d) * std::sqrt(sqDistOf(c, d));
}

inline V3 edgePerpendicular(const Classified& c, const TriData& d) {   // q minus its component along the edge
    const V3 q = c.q;
    if (c.r == R_E1) return v3(0.0f, q.y, q.z);
    if (c.r == R_E2) {
        const float t = (q.x - d.v2) * d.b[0] + q.y * d.b[1];
        return v3((q.x - d.v2) - t * d.b[0], q.y - t * d.b[1], q.z);
    }
    const float t = q.x * d.c[0] + q.y * d.c[1];
    return v3(q.x - t * d.c[0], q.y - t * d.c[1], q.z);
}

// :198-290 — gradient variant used by the tri-cubic fit (needs the world-space vertices)
float signedDistGradMesh(V3 p, const TriData& d, V3 w1, V3 w2, V3 w3, V3& outN) {
    Classified c = classify(p, d);
    if (c.r == R_FACE) { outN = triNormal(d); return c.q.z; }
    auto guarded = [&d](V3 v) { V3 n = normalize(v); return std::isnan(n.x + n.y + n.z) ? triNormal(d) : n; };
    const float s = regionSign(c, d);
    switch (c.r) {
        case R_V1: outN = s * guarded(p - w1); break;
        case R_V2: outN = s * guarded(p - w2); break;
        case R_V3: outN = s * guarded(p - w3); break;
        default: outN = s * guarded(matTMul(d.T, edgePerpendicular(c, d))); break;
    }
    return s * std::sqrt(sqDistOf(c, d));
}

// :292-376 — self-contained gradient variant used by the exact-octree query
float signedDistGradSelf(V3 p, const TriData& d, V3& outN) {
    Classified c = classify(p, d);
    if (c.r == R_FACE) { outN = triNormal(d); return c.q.z; }
    const float s = regionSign(c, d);
    switch (c.r) {
        case R_V1: outN = s * normalize(p - d.origin); break;
        case R_V2: outN = s * normalize(p - d.origin - matTMul(d.T, v3(d.v2, 0.0f, 0.0f))); break;
        case R_V3: outN = s * normalize(p - d.origin - matTMul(d.T, v3(d.v3[0], d.v3[1], 0.0f))); break;
        default: outN = s * normalize(matTMul(d.T, edgePerpendicular(c, d))); break;
    }
    return s * std::sqrt(sqDistOf(c, d));
}

// ------------------------------------------------------------------------------------------
// a5: nearest triangle through a bounding-sphere BVH in float64
// (libs/InteractiveComputerGraphics/InteractiveComputerGraphics/TriangleMeshDistance.h:421-798).
// Third-party algorithm (ICG TriangleMeshDistance, MIT); the point-triangle routine is Eberly's
// published closest-point algorithm. SdfLib keeps only the triangle id (TrianglesInfluence.h:898-905).
// ------------------------------------------------------------------------------------------
struct D3 { double x, y, z; };
inline D3 operator+(D3 a, D3 b) { return D3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D3 operator-(D3 a, D3 b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline double ddot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double dnorm(D3 a) { return std::sqrt(ddot(a, a)); }
inline double dcomp(const D3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct Sphere { D3 c; double r; };
struct BvhNode { Sphere left, right; int l = -1, r = -1; };   // l == -1: leaf, r = triangle id

// Squared distance only (the reference also tracks the closest point/entity, unused by SdfLib).
double eberlySqDist(D3 p, D3 v0, D3 v1, D3 v2) {
    const D3 diff = v0 - p, e0 = v1 - v0, e1 = v2 - v0;
    const double a00 = ddot(e0, e0), a01 = ddot(e0, e1), a11 = ddot(e1, e1);
    const double b0 = ddot(diff, e0), b1 = ddot(diff, e1), c = ddot(diff, diff);
    const double det = std::abs(a00 * a11 - a01 * a01);
    double s = a01 * b1 - a11 * b0;
    double t = a01 * b0 - a00 * b1;
    double d2;
    // closest point on the edge through v0 along e0 / e1, clamped to the segment
    auto onEdge0 = [&]() { return (b0 >= 0) ? c : ((-b0 >= a00) ? a00 + 2 * b0 + c : b0 * (-b0 / a00) + c); };
    auto onEdge1 = [&]() { return (b1 >= 0) ? c : ((-b1 >= a11) ? a11 + 2 * b1 + c : b1 * (-b1 / a11) + c); };
    auto quad = [&](double ss, double tt) {
        return ss * (a00 * ss + a01 * tt + 2 * b0) + tt * (a01 * ss + a11 * tt + 2 * b1) + c;
    };
    if (s + t <= det) {
        if (s < 0) {
            if (t < 0) {   // region 4
                if (b0 < 0) d2 = (-b0 >= a00) ? a00 + 2 * b0 + c : b0 * (-b0 / a00) + c;
                else d2 = onEdge1();
            } else d2 = onEdge1();   // region 3
        } else if (t < 0) d2 = onEdge0();   // region 5
        else {   // region 0
            const double inv = 1 / det;
            s *= inv; t *= inv;
            d2 = quad(s, t);
        }
    } else {
        if (s < 0) {   // region 2
            const double tmp0 = a01 + b0, tmp1 = a11 + b1;
            if (tmp1 > tmp0) {
                const double numer = tmp1 - tmp0, denom = a00 - 2 * a01 + a11;
                if (numer >= denom) d2 = a00 + 2 * b0 + c;
                else { s = numer / denom; t = 1 - s; d2 = quad(s, t); }
            } else {
                if (tmp1 <= 0) d2 = a11 + 2 * b1 + c;
                else if (b1 >= 0) d2 = c;
                else d2 = b1 * (-b1 / a11) + c;
            }
        } else if (t < 0) {   // region 6
            const double tmp0 = a01 + b1, tmp1 = a00 + b0;
            if (tmp1 > tmp0) {
                const double numer = tmp1 - tmp0, denom = a00 - 2 * a01 + a11;
                if (numer >= denom) d2 = a11 + 2 * b1 + c;
                else { t = numer / denom; s = 1 - t; d2 = quad(s, t); }
            } else {
                if (tmp1 <= 0) d2 = a00 + 2 * b0 + c;
                else if (b0 >= 0) d2 = c;
                else d2 = b0 * (-b0 / a00) + c;
            }
        } else {   // region 1
            const double numer = a11 + b1 - a01 - b0;
            if (numer <= 0) d2 = a11 + 2 * b1 + c;
            else {
                const double denom = a00 - 2 * a01 + a11;
                if (numer >= denom) d2 = a00 + 2 * b0 + c;
                else { s = numer / denom; t = 1 - s; d2 = quad(s, t); }
            }
        }
    }
    return d2 < 0 ? 0 : d2;
}

struct Bvh {
    std::vector<D3> verts;
    std::vector<std::array<int, 3>> tris;
    std::vector<BvhNode> nodes;
    Sphere root;
    // traversal model only (tests/model_bvh_boxes.py): axis-aligned box of every subtree, by node id (float, exact: min / max of float coordinates)
    struct Box { float lo[3], hi[3]; };
    std::vector<Box> boxes;
    bool useBoxes = false;
    double boxMargin = 0.0;

    struct BuildTri { D3 v[3]; int id; };

    explicit Bvh(const MeshView& m) {
        verts.resize(m.nVerts);
        for (uint32_t i = 0; i < m.nVerts; i++) verts[i] = D3{double(m.verts[i].x), double(m.verts[i].y), double(m.verts[i].z)};
        tris.resize(m.nIdx / 3);
        std::vector<BuildTri> bt(tris.size());
        for (size_t t = 0; t < tris.size(); t++) {
            tris[t] = {int(m.idx[3 * t]), int(m.idx[3 * t + 1]), int(m.idx[3 * t + 2])};
            bt[t].id = int(t);
            for (int k = 0; k < 3; k++) bt[t].v[k] = verts[size_t(tris[t][size_t(k)])];
        }
        nodes.reserve(2 * tris.size());
        nodes.push_back(BvhNode());
        build(0, -1, 0, bt, 0, int(bt.size()));
    }
    // boxes of all subtrees, bottom-up over the finished tree (children have larger ids than their parent)
    void buildBoxes() {
        boxes.assign(nodes.size(), Box{{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}});
        for (size_t k = nodes.size(); k-- > 0;) {
            Box& b = boxes[k];
            if (nodes[k].l == -1) {
                for (int c = 0; c < 3; c++) {
                    const D3& p = verts[size_t(tris[size_t(nodes[k].r)][size_t(c)])];
                    const float q[3] = {float(p.x), float(p.y), float(p.z)};
                    for (int a = 0; a < 3; a++) { b.lo[a] = std::min(b.lo[a], q[a]); b.hi[a] = std::max(b.hi[a], q[a]); }
                }
            } else {
                for (int child : {nodes[k].l, nodes[k].r})
                    for (int a = 0; a < 3; a++) { b.lo[a] = std::min(b.lo[a], boxes[size_t(child)].lo[a]); b.hi[a] = std::max(b.hi[a], boxes[size_t(child)].hi[a]); }
            }
        }
        double s2 = 0;
        for (int a = 0; a < 3; a++) s2 += double(boxes[0].hi[a] - boxes[0].lo[a]) * double(boxes[0].hi[a] - boxes[0].lo[a]);
        boxMargin = 1e-10 * s2;
        useBoxes = true;
    }
    // conservative squared distance from p to the box of node k (float arithmetic, scaled down by more than its rounding error)
    bool boxBeyondBest(int k, D3 p, double best) const {
        const Box& b = boxes[size_t(k)];
        const float q[3] = {float(p.x), float(p.y), float(p.z)};
        float d2 = 0.f;
        for (int a = 0; a < 3; a++) { const float d = std::max(std::max(b.lo[a] - q[a], q[a] - b.hi[a]), 0.f); d2 += d * d; }
        return double(d2 * (1.0f - 1e-6f)) > best * best * (1.0 + 1e-9) + boxMargin;
    }

    // :421-490. `slot` says where this subtree's bounding sphere lives: 0 root, 1 parent.left, 2 parent.right.
    void build(int nodeId, int parent, int slot, std::vector<BuildTri>& bt, int begin, int end) {
        const int n = end - begin;
        Sphere sph;
        if (n == 1) {
            const BuildTri& tr = bt[size_t(begin)];
            D3 s = tr.v[0] + tr.v[1] + tr.v[2];
            D3 c = D3{s.x / 3.0, s.y / 3.0, s.z / 3.0};
            sph.c = c;
            sph.r = std::max(std::max(dnorm(tr.v[0] - c), dnorm(tr.v[1] - c)), dnorm(tr.v[2] - c));
            storeSphere(parent, slot, sph);
            nodes[size_t(nodeId)].l = -1;
            nodes[size_t(nodeId)].r = tr.id;
            return;
        }
        const double lo = std::numeric_limits<double>::lowest(), hi = std::numeric_limits<double>::max();
        D3 top{lo, lo, lo}, bottom{hi, hi, hi}, c{0, 0, 0};
        for (int i = begin; i < end; i++)
            for (int k = 0; k < 3; k++) {
                const D3& p = bt[size_t(i)].v[k];
                c = c + p;
                top = D3{std::max(top.x, p.x), std::max(top.y, p.y), std::max(top.z, p.z)};
                bottom = D3{std::min(bottom.x, p.x), std::min(bottom.y, p.y), std::min(bottom.z, p.z)};
            }
        const double cnt = double(3 * n);
        c = D3{c.x / cnt, c.y / cnt, c.z / cnt};
        const D3 diag = top - bottom;
        int dim = 0;   // first largest extent (std::max_element)
        if (diag.y > dcomp(diag, dim)) dim = 1;
        if (diag.z > dcomp(diag, dim)) dim = 2;
        double r2 = 0.0;
        for (int i = begin; i < end; i++)
            for (int k = 0; k < 3; k++) { D3 d = c - bt[size_t(i)].v[k]; r2 = std::max(r2, ddot(d, d)); }
        sph.c = c; sph.r = std::sqrt(r2);
        storeSphere(parent, slot, sph);
        // std::sort (introsort, not stable): ties between triangles that share their first vertex are
        // resolved by the library's comparison sequence, so the same call must be used here.
        std::sort(bt.begin() + begin, bt.begin() + end,
                  [dim](const BuildTri& a, const BuildTri& b) { return dcomp(a.v[0], dim) < dcomp(b.v[0], dim); });
        const int mid = int(0.5 * (begin + end));
        const int l = int(nodes.size());
        nodes[size_t(nodeId)].l = l;
        nodes.push_back(BvhNode());
        build(l, nodeId, 1, bt, begin, mid);
        const int r = int(nodes.size());
        nodes[size_t(nodeId)].r = r;
        nodes.push_back(BvhNode());
        build(r, nodeId, 2, bt, mid, end);
    }
    void storeSphere(int parent, int slot, const Sphere& s) {
        if (slot == 0) root = s;
        else if (slot == 1) nodes[size_t(parent)].left = s;
        else nodes[size_t(parent)].right = s;
    }

    // :492-540 — near child first, strict '<', the running best is sqrt'ed and re-squared.
    mutable uint32_t* visitCount = nullptr;   // optional counters {inner nodes, leaves} of the current query (traversal models in tests/)
    void query(const BvhNode& nd, D3 p, double& best, int& bestTri) const {
        if (visitCount) visitCount[nd.l == -1 ? 1 : 0]++;
        if (nd.l == -1) {
            const auto& tr = tris[size_t(nd.r)];
            const double d2 = eberlySqDist(p, verts[size_t(tr[0])], verts[size_t(tr[1])], verts[size_t(tr[2])]);
            if (d2 < best * best) { best = std::sqrt(d2); bestTri = nd.r; }
            return;
        }
        const double dl = dnorm(p - nd.left.c) - nd.left.r;
        const double dr = dnorm(p - nd.right.c) - nd.right.r;
        // model of the device's extra pruning: a child whose box lies beyond the running best (by a margin above every rounding
        // error of the leaf test) cannot change the result, so it is skipped although its sphere passes
        auto visit = [&](int child, double d) {
            if (!(d < best)) return;
            if (useBoxes && boxBeyondBest(child, p, best)) return;
            query(nodes[size_t(child)], p, best, bestTri);
        };
        if (dl < dr) { visit(nd.l, dl); visit(nd.r, dr); }
        else { visit(nd.r, dr); visit(nd.l, dl); }
    }
    uint32_t nearest(V3 p, double seed = std::numeric_limits<double>::max()) const {   // seed: traversal models only (an upper bound the search starts from)
        double best = seed;
        int tri = -1;
        query(nodes[0], D3{double(p.x), double(p.y), double(p.z)}, best, tri);
        return uint32_t(tri);
    }
};

// ------------------------------------------------------------------------------------------
// a10-a13: tri-cubic Hermite fit, evaluation and error integral
// (include/SdfLib/InterpolationMethods.h:267-498, include/SdfLib/OctreeSdfUtils.h:60-138,213-238)
// ------------------------------------------------------------------------------------------
typedef std::array<float, 8> PointValues;   // f, fx, fy, fz, fxy, fxz, fyz, fxyz
typedef std::array<float, 64> Coeffs;

const int kHermite[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {-3, -2, 3, -1}, {2, 1, -2, 1}};
const int kSlotOrder[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};

// calculateCoefficients (:292-378): derivatives scaled by nodeSize^order, then every coefficient is the
// left-to-right float sum of w*in[col] over the non-zero columns (corner*8+slot) in increasing order,
// w = H[i][2cx+dx]*H[j][2cy+dy]*H[k][2cz+dz].
void tricubicCoefficients(const std::array<PointValues, 8>& input, float nodeSize, Coeffs& out) {
    std::array<PointValues, 8> in = input;
    for (int c = 0; c < 8; c++) {
        in[c][1] *= nodeSize; in[c][2] *= nodeSize; in[c][3] *= nodeSize;
        const float sq = nodeSize * nodeSize;
        in[c][4] *= sq; in[c][5] *= sq; in[c][6] *= sq;
        in[c][7] *= sq * nodeSize;
    }
    for (int row = 0; row < 64; row++) {
        const int i = row & 3, j = (row >> 2) & 3, k = row >> 4;
        float acc = 0.0f;
        bool first = true;
        for (int c = 0; c < 8; c++)
            for (int s = 0; s < 8; s++) {
                const int w = kHermite[i][2 * (c & 1) + kSlotOrder[s][0]] * kHermite[j][2 * ((c >> 1) & 1) + kSlotOrder[s][1]] *
                              kHermite[k][2 * ((c >> 2) & 1) + kSlotOrder[s][2]];
                if (w == 0) continue;
                const float term = float(w) * in[c][s];
                acc = first ? term : acc + term;
                first = false;
            }
        out[size_t(row)] = acc;
    }
}

inline float monomial(float coef, int i, int j, int k, V3 f) {   // ((c*x..)*y..)*z.. left to right
    float t = coef;
    for (int a = 0; a < i; a++) t *= f.x;
    for (int a = 0; a < j; a++) t *= f.y;
    for (int a = 0; a < k; a++) t *= f.z;
    return t;
}

float tricubicValue(const float* c, V3 f) {   // interpolateValue scalar branch (:432-439)
    float acc = 0.0f;
    for (int n = 0; n < 64; n++) acc += monomial(c[n], n & 3, (n >> 2) & 3, n >> 4, f);
    return acc;
}

// generic derivative polynomial: d^(ox+oy+oz) / dx^ox dy^oy dz^oz, terms in increasing n,
// integer factor (falling factorial product) applied to the coefficient first.
float tricubicDerivative(const float* c, V3 f, int ox, int oy, int oz) {
    auto ff = [](int p, int o) { int r = 1; for (int a = 0; a < o; a++) r *= (p - a); return r; };
    float acc = 0.0f;
    bool first = true;
    for (int n = 0; n < 64; n++) {
        const int i = n & 3, j = (n >> 2) & 3, k = n >> 4;
        if (i < ox || j < oy || k < oz) continue;
        const int w = ff(i, ox) * ff(j, oy) * ff(k, oz);
        const float term = monomial(float(w) * c[n], i - ox, j - oy, k - oz, f);
        acc = first ? term : acc + term;
        first = false;
    }
    return acc;
}

V3 tricubicGradient(const float* c, V3 f) {   // interpolateGradient (:442-455)
    return V3{tricubicDerivative(c, f, 1, 0, 0), tricubicDerivative(c, f, 0, 1, 0), tricubicDerivative(c, f, 0, 0, 1)};
}

void tricubicVertexValues(const float* c, V3 f, float nodeSize, PointValues& out) {   // :457-497
    out[0] = tricubicValue(c, f);
    out[1] = tricubicDerivative(c, f, 1, 0, 0) / nodeSize;
    out[2] = tricubicDerivative(c, f, 0, 1, 0) / nodeSize;
    out[3] = tricubicDerivative(c, f, 0, 0, 1) / nodeSize;
    const float sq = nodeSize * nodeSize;
    out[4] = tricubicDerivative(c, f, 1, 1, 0) / sq;
    out[5] = tricubicDerivative(c, f, 1, 0, 1) / sq;
    out[6] = tricubicDerivative(c, f, 0, 1, 1) / sq;
    out[7] = tricubicDerivative(c, f, 1, 1, 1) / (sq * nodeSize);
}

// OctreeSdfUtils.h:60-85 (trapezoid), :87-138 (decay by distance). Weight 2^(#centred axes)/64.
float errorTrapezoid(const Coeffs& c, const std::array<PointValues, 19>& mid, bool decayRule, float decay) {
    float acc = 0.0f;
    for (int s = 0; s < 19; s++) {
        const V3 rel = kLattice.sampleRel[s];
        const int zeros = (rel.x == 0.0f) + (rel.y == 0.0f) + (rel.z == 0.0f);
        const float w = (zeros == 1 ? 2.0f : (zeros == 2 ? 4.0f : 8.0f)) / 64.0f;
        const V3 f = V3{0.5f * rel.x + 0.5f, 0.5f * rel.y + 0.5f, 0.5f * rel.z + 0.5f};
        const float v = tricubicValue(c.data(), f);
        float term;
        if (!decayRule) { const float d = mid[size_t(s)][0] - v; term = w * (d * d); }
        else { const float d = fmax_(fabs_(mid[size_t(s)][0] - v) - decay * fabs_(v), 0.0f); term = w * (d * d); }
        acc = (s == 0 && !decayRule) ? term : acc + term;
    }
    return acc;
}

// OctreeSdfUtils.h:213-238 (Simpson): weight 4^(#centred axes)/216, the float constant w/216 is
// formed first, then multiplied by the squared difference; terms summed left to right in sample order.
float errorSimpson(const Coeffs& c, const std::array<PointValues, 19>& mid) {
    float acc = 0.0f;
    for (int s = 0; s < 19; s++) {
        const V3 rel = kLattice.sampleRel[s];
        const int zeros = (rel.x == 0.0f) + (rel.y == 0.0f) + (rel.z == 0.0f);
        const float w = (zeros == 1 ? 4.0f : (zeros == 2 ? 16.0f : 64.0f)) / 216.0f;
        const V3 f = V3{0.5f * rel.x + 0.5f, 0.5f * rel.y + 0.5f, 0.5f * rel.z + 0.5f};
        const float d = mid[size_t(s)][0] - tricubicValue(c.data(), f);
        const float term = w * (d * d);
        acc = (s == 0) ? term : acc + term;
    }
    return acc;
}

// switch at OctreeSdfDepthFirst.h:194-212 / OctreeSdfBreadthFirstNoDelay.h:346-363
float errorByRule(int rule, const Coeffs& c, const std::array<PointValues, 19>& mid, float decay) {
    switch (rule) {
        case 1: return errorTrapezoid(c, mid, false, 0.0f);
        case 2: return errorSimpson(c, mid);
        case 3: return errorTrapezoid(c, mid, true, decay);
        default: return INFINITY;
    }
}

}  // namespace

namespace {

// ------------------------------------------------------------------------------------------
// a22: BoundingBox distance (include/SdfLib/utils/Mesh.h:42-63)
// ------------------------------------------------------------------------------------------
struct Box {
    V3 mn, mx;
    V3 size() const { return mx - mn; }
    V3 center() const { return mn + 0.5f * size(); }
    float distance(V3 p) const {
        V3 d = p - center();
        V3 q = V3{fabs_(d.x), fabs_(d.y), fabs_(d.z)} - 0.5f * size();
        V3 qp = V3{fmax_(q.x, 0.0f), fmax_(q.y, 0.0f), fmax_(q.z, 0.0f)};
        return length(qp) + fmin_(fmax_(q.x, fmax_(q.y, q.z)), 0.0f);
    }
    // the reference's gradient variant is not box-centred (uses |p| - size); kept as is.
    float distance(V3 p, V3& g) const {
        const V3 s = size();
        float a[3] = {fabs_(p.x) - s.x, fabs_(p.y) - s.y, fabs_(p.z) - s.z};
        float pp[3] = {p.x, p.y, p.z};
        float gg[3] = {g.x, g.y, g.z};
        int k = a[0] > a[1] ? 0 : 1;
        int l = a[2] > a[k] ? 2 : k;
        if (a[l] < 0) gg[l] = pp[l] / fabs_(pp[l]);
        else {
            float b[3] = {fmax_(a[0], 0.0f), fmax_(a[1], 0.0f), fmax_(a[2], 0.0f)};
            float c = std::sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
            for (int i = 0; i < 3; i++) gg[i] = a[i] > 0 ? b[i] / c * pp[i] / fabs_(pp[i]) : 0;
        }
        g = V3{gg[0], gg[1], gg[2]};
        return distance(p);
    }
};

// ------------------------------------------------------------------------------------------
// 32^3 direct-mapped vertex cache keyed by lattice coordinates at maxDepth resolution
// (TrianglesInfluence.h:676-690, :934-950)
// ------------------------------------------------------------------------------------------
struct VertexCache {
    struct Entry { uint32_t x, y, z, tri; };
    std::vector<Entry> entries;
    V3 coordToId, minPoint;
    bool enabled;
    void init(const Box& box, uint32_t maxDepth, bool on) {
        enabled = on;
        const uint32_t none = (1u << maxDepth) + 1;
        entries.assign(32 * 32 * 32, Entry{none, none, none, 0});
        const V3 s = box.size();
        const float n = float(1 << maxDepth);
        coordToId = V3{n / s.x, n / s.y, n / s.z};
        minPoint = box.mn;
    }
    // returns true on hit
    bool lookup(V3 p, uint32_t& slot, uint32_t key[3], uint32_t& tri) const {
        const V3 r = (p - minPoint) * coordToId;
        key[0] = uint32_t(std::round(r.x)); key[1] = uint32_t(std::round(r.y)); key[2] = uint32_t(std::round(r.z));
        slot = ((key[2] & 31u) << 10) | ((key[1] & 31u) << 5) | (key[0] & 31u);
        const Entry& e = entries[slot];
        if (enabled && e.x == key[0] && e.y == key[1] && e.z == key[2]) { tri = e.tri; return true; }
        return false;
    }
    void store(uint32_t slot, const uint32_t key[3], uint32_t tri) { entries[slot] = Entry{key[0], key[1], key[2], tri}; }
};

// ------------------------------------------------------------------------------------------
// the structure (both formats) + .bin layout
// ------------------------------------------------------------------------------------------
const uint32_t LEAF_BIT = 1u << 31;
const uint32_t OCT_INDEX_MASK = ~(3u << 30);       // OctreeSdf.h:53-55 (leaf + mark bits)
const uint32_t EXACT_INDEX_MASK = ~(1u << 31);     // ExactOctreeSdf.h:53-54

}  // namespace

struct OrcSdf {
    int format = 1;   // SdfFunction::SdfFormat: 1 OCTREE, 2 EXACT_OCTREE (SdfFunction.h:16-22)
    Box box;
    int startGridSize = 0;
    uint32_t maxDepth = 0;
    float cellSize = 0;
    // OCTREE
    float valueRange = 0, minBorderValue = 0;
    std::vector<uint32_t> octree;
    // EXACT_OCTREE
    uint32_t startDepth = 0, minTrisInLeafs = 0, maxTrisInLeafs = 0, maxTrisEncoded = 0, bitEncodingStartDepth = 0,
             bitsPerIndex = 0;
    std::vector<uint32_t> nodes;   // pairs (childrenIndex, trianglesArrayIndex)
    std::vector<uint32_t> sets;
    std::vector<uint8_t> masks;
    std::vector<TriData> tris;
};

namespace {

void cubify(OrcSdf& s, const float* box6, uint32_t startDepth) {   // OctreeSdf.cpp:43-51, ExactOctreeSdf.cpp:13-22
    Box in{v3(box6[0], box6[1], box6[2]), v3(box6[3], box6[4], box6[5])};
    const V3 sz = in.size();
    const float maxSize = fmax_(fmax_(sz.x, sz.y), sz.z);
    s.box.mn = in.center() - 0.5f * maxSize;
    s.box.mx = in.center() + 0.5f * maxSize;
    s.startGridSize = 1 << startDepth;
    s.cellSize = maxSize / float(s.startGridSize);
}

inline uint32_t startSlot(const OrcSdf& s, V3 center) {   // OctreeSdfDepthFirst.h:408-409
    V3 f = (center - s.box.mn) / s.cellSize;
    int x = int(std::floor(f.x)), y = int(std::floor(f.y)), z = int(std::floor(f.z));
    return uint32_t(z * s.startGridSize * s.startGridSize + y * s.startGridSize + x);
}

// ------------------------------------------------------------------------------------------
// a6 + a14: OctreeSdf depth-first build, NO_CONTINUITY (src/sdf/OctreeSdfDepthFirst.h:31-527)
// ------------------------------------------------------------------------------------------
struct OctBuildNode {
    uint32_t nodeIndex; uint32_t depth; V3 center; float half;
    std::array<PointValues, 8> values;
    std::array<uint32_t, 8> info;
};

struct OctBuilder {
    OrcSdf& out;
    MeshView mesh;
    std::vector<TriData> tris;
    std::unique_ptr<Bvh> bvh;
    VertexCache cache;
    uint32_t startDepth, maxDepth;
    int rule; float p0, p1;
    float valueRange = 0.0f;

    // VHQueries::calculateVerticesInfo (TrianglesInfluence.h:952-996) + calculatePointValues
    // (InterpolationMethods.h:273-290)
    void sample(V3 center, float half, V3 rel, PointValues& vals, uint32_t& tri) {
        const V3 p = center + rel * half;
        uint32_t slot, key[3];
        if (!cache.lookup(p, slot, key, tri)) { tri = bvh->nearest(p); cache.store(slot, key, tri); }
        V3 g;
        vals[0] = signedDistGradMesh(p, tris[tri], mesh.verts[mesh.idx[3 * tri]], mesh.verts[mesh.idx[3 * tri + 1]],
                                     mesh.verts[mesh.idx[3 * tri + 2]], g);
        vals[1] = g.x; vals[2] = g.y; vals[3] = g.z;
        vals[4] = vals[5] = vals[6] = vals[7] = 0.0f;
    }

    void emitLeaf(const OctBuildNode& n, std::vector<uint32_t>& oct) {   // :343-362, :371-390
        const uint32_t at = uint32_t(oct.size());
        oct[n.nodeIndex] = (at & OCT_INDEX_MASK) | LEAF_BIT;
        Coeffs c;
        tricubicCoefficients(n.values, 2.0f * n.half, c);
        oct.resize(oct.size() + 64);
        std::memcpy(&oct[at], c.data(), 64 * sizeof(float));
        for (int i = 0; i < 8; i++) valueRange = fmax_(valueRange, fabs_(n.values[size_t(i)][0]));
    }

    // processNode (:137-391). Children are pushed 0..7 so the stack pops 7 first.
    void process(const OctBuildNode& n, std::vector<OctBuildNode>& stack, std::vector<uint32_t>& oct) {
        if (n.depth >= maxDepth) { emitLeaf(n, oct); return; }
        std::array<PointValues, 19> mid;
        std::array<uint32_t, 19> midInfo;
        for (int s = 0; s < 19; s++) sample(n.center, n.half, kLattice.sampleRel[s], mid[size_t(s)], midInfo[size_t(s)]);
        bool terminal = false;
        if (n.depth >= startDepth) {
            Coeffs c;
            tricubicCoefficients(n.values, 2.0f * n.half, c);
            const float value = errorByRule(rule, c, mid, p1);
            terminal = value < p0 * p0;   // NONE: value = INFINITY (:206-208)
        }
        if (terminal) { emitLeaf(n, oct); return; }
        const float h = 0.5f * n.half;
        const bool real = n.depth >= startDepth;
        const uint32_t childIndex = real ? uint32_t(oct.size()) : 0xFFFFFFFFu;
        if (n.nodeIndex != 0xFFFFFFFFu) oct[n.nodeIndex] = childIndex & OCT_INDEX_MASK;
        if (real) oct.resize(oct.size() + 8);
        for (uint32_t c = 0; c < 8; c++) {
            OctBuildNode ch;
            ch.nodeIndex = real ? childIndex + c : 0xFFFFFFFFu;
            ch.depth = n.depth + 1;
            ch.center = n.center + cornerDir(c) * h;   // center + vec3(±h)
            ch.half = h;
            for (uint32_t k = 0; k < 8; k++) {
                const int L = int((c & 1) + (k & 1)) + 3 * int(((c >> 1) & 1) + ((k >> 1) & 1)) + 9 * int((c >> 2) + (k >> 2));
                if (kLattice.cornerOfLattice[L] >= 0) {
                    ch.values[k] = n.values[size_t(kLattice.cornerOfLattice[L])];
                    ch.info[k] = n.info[size_t(kLattice.cornerOfLattice[L])];
                } else {
                    ch.values[k] = mid[size_t(kLattice.sampleOfLattice[L])];
                    ch.info[k] = midInfo[size_t(kLattice.sampleOfLattice[L])];
                }
            }
            stack.push_back(ch);
        }
    }

    void run(uint32_t numThreads) {
        const uint32_t d0 = std::min(startDepth, 1u);
        const float h0 = float(0.5f * out.box.size().x * std::pow(0.5f, d0));   // std::pow(float, unsigned) -> double
        const V3 c0 = out.box.mn + h0;
        const uint32_t per = 1u << d0;
        std::vector<OctBuildNode> stack;
        for (uint32_t k = 0; k < per; k++)
            for (uint32_t j = 0; j < per; j++)
                for (uint32_t i = 0; i < per; i++) {
                    OctBuildNode n;
                    n.nodeIndex = 0xFFFFFFFFu; n.depth = d0; n.half = h0;
                    n.center = c0 + v3(float(i), float(j), float(k)) * 2.0f * h0;
                    for (uint32_t c = 0; c < 8; c++) sample(n.center, n.half, cornerDir(c), n.values[c], n.info[c]);
                    stack.push_back(n);
                }
        const uint32_t G = uint32_t(out.startGridSize);
        out.octree.assign(size_t(G) * G * G, 0u);
        while (!stack.empty()) {
            OctBuildNode n = stack.back();
            stack.pop_back();
            if (n.depth == startDepth) n.nodeIndex = startSlot(out, n.center);
            process(n, stack, out.octree);
        }
        out.valueRange = valueRange;
        if (numThreads >= 2) relayoutPerVoxel();
    }

    // Layout of the multi-threaded driver (:417-503): G^3 start slots, then every start voxel's
    // sub-octree (blocks in that voxel's own DFS order) in voxel-index order. The reference's racy MT
    // run is not reproduced; this is the single-thread content in the intended MT layout.
    void relayoutPerVoxel() {
        const std::vector<uint32_t> src = out.octree;
        const uint32_t G3 = uint32_t(out.startGridSize) * out.startGridSize * out.startGridSize;
        std::vector<uint32_t> dst(G3);
        for (uint32_t v = 0; v < G3; v++) {
            // explicit stack of (srcNodeSlot, dstNodeSlot); DFS pops child 7 first like the builder
            std::vector<std::pair<uint32_t, uint32_t>> st;
            st.push_back({v, v});
            while (!st.empty()) {
                auto [s, d] = st.back();
                st.pop_back();
                const uint32_t w = src[s];
                const uint32_t at = uint32_t(dst.size());
                if (w & LEAF_BIT) {
                    dst[d] = (at & OCT_INDEX_MASK) | LEAF_BIT;
                    dst.insert(dst.end(), src.begin() + (w & OCT_INDEX_MASK), src.begin() + (w & OCT_INDEX_MASK) + 64);
                } else {
                    dst[d] = at & OCT_INDEX_MASK;
                    dst.resize(dst.size() + 8);
                    for (uint32_t c = 0; c < 8; c++) st.push_back({(w & OCT_INDEX_MASK) + c, at + c});
                }
            }
        }
        out.octree.swap(dst);
    }
};

// ------------------------------------------------------------------------------------------
// a15: OctreeSdf breadth-first build with C1 continuity across T-junctions, InitAlgorithm::CONTINUITY
// (src/sdf/OctreeSdfBreadthFirstNoDelay.h:84-1224; helpers src/sdf/OctreeSdfBreadthFirst.h:35-89).
//
// Vocabulary used below. A node tracks six "outer" neighbour links, one per non-empty proper axis
// set n in 1..6 (bit0 x, bit1 y, bit2 z): the region reached by leaving the node's PARENT through
// the faces the node touches (sign of each axis = the node's own child bit). A link is either
//   * bit31 set: word index of a coarser LEAF covering that region,
//   * bit30 set: outside the start grid,
//   * else: index of an 8-word children block (its depth is linkDepth), or, right after creation,
//     the word of the parent-sized neighbour.
// The 18 face/edge neighbours of a node are probed through (link of the outward axes | own sibling
// block) + child id (dir ^ c). Which of the 19 mid-point samples lie on the face/edge shared with a
// neighbour is a fixed bit table (:139-176): bit (18 - s) for sample s, entry 4*(dir-1) + sign, where
// sign packs "positive side?" of dir's axes, lowest axis in bit 0.
// ------------------------------------------------------------------------------------------
struct ContNode {
    uint64_t path = 0;                 // childIndices: 3 bits per level, own child id in the low 3
    uint32_t parentBlock = 0xFFFFFFFFu;
    bool terminal = false, ignore = false;
    uint8_t linkDepth[6] = {0, 0, 0, 0, 0, 0};
    uint32_t link[6] = {0, 0, 0, 0, 0, 0};
    V3 center{0, 0, 0};
    float half = 0;
    std::array<PointValues, 8> values{};
    Coeffs coeff{};
    std::array<PointValues, 19> mid{};
};

struct ContBuilder {
    OrcSdf& out;
    MeshView mesh;
    std::vector<TriData> tris;
    std::unique_ptr<Bvh> bvh;
    VertexCache cacheMain;   // `trianglesInfluence`: start nodes + fix-up pass (:193, :917)
    VertexCache cacheIter;   // `threadTrianglesInfluence[0]`: copy taken after the start nodes (:253), Iter 1
    uint32_t startDepth, maxDepth;
    int rule; float p0, p1;

    static constexpr uint32_t B31 = 1u << 31, B30 = 1u << 30, MARK = 1u << 30;
    uint32_t faceMask[24];

    static inline int axisSign(uint32_t dir, uint32_t sign, int axis) {   // +1 / -1 side of `axis` (in dir)
        int bit = 0;
        for (int a = 0; a < axis; a++) if (dir & (1u << a)) bit++;
        return ((sign >> bit) & 1u) ? 1 : -1;
    }
    void buildMaskTable() {
        for (uint32_t dir = 1; dir <= 6; dir++)
            for (uint32_t sign = 0; sign < 4; sign++) {
                uint32_t m = 0;
                const int nAxes = int((dir & 1) + ((dir >> 1) & 1) + ((dir >> 2) & 1));
                if (sign < (1u << nAxes))
                    for (int s = 0; s < 19; s++) {
                        const V3 r = kLattice.sampleRel[s];
                        const float rr[3] = {r.x, r.y, r.z};
                        bool on = true;
                        for (int a = 0; a < 3; a++)
                            if ((dir & (1u << a)) && rr[a] != float(axisSign(dir, sign, a))) on = false;
                        if (on) m |= 1u << (18 - s);
                    }
                faceMask[4 * (dir - 1) + sign] = m;
            }
    }
    // sign index of direction `dir` for a node with child id c: axes in `outward` go to the node's own
    // side (bit of c), the others to the opposite side (:419-438)
    static inline uint32_t signOf(uint32_t dir, uint32_t outward, uint32_t c) {
        uint32_t sign = 0, bit = 0;
        for (uint32_t a = 0; a < 3; a++)
            if (dir & (1u << a)) {
                const uint32_t pos = (outward & (1u << a)) ? ((c >> a) & 1u) : (((~c) >> a) & 1u);
                sign |= pos << bit;
                bit++;
            }
        return sign;
    }

    inline bool isLeaf(uint32_t w) const { return (out.octree[w] & B31) != 0; }
    inline bool isMarked(uint32_t w) const { return (out.octree[w] & MARK) != 0; }
    inline uint32_t childrenOf(uint32_t w) const { return out.octree[w] & OCT_INDEX_MASK; }
    inline void setWord(uint32_t w, bool leaf, uint32_t index) { out.octree[w] = (index & OCT_INDEX_MASK) | (leaf ? B31 : 0u); }

    void trueSample(VertexCache& cache, V3 p, PointValues& vals) {   // TrianglesInfluence.h:971-992
        uint32_t slot, key[3], tri;
        if (!cache.lookup(p, slot, key, tri)) { tri = bvh->nearest(p); cache.store(slot, key, tri); }
        V3 g;
        vals[0] = signedDistGradMesh(p, tris[tri], mesh.verts[mesh.idx[3 * tri]], mesh.verts[mesh.idx[3 * tri + 1]],
                                     mesh.verts[mesh.idx[3 * tri + 2]], g);
        vals[1] = g.x; vals[2] = g.y; vals[3] = g.z;
        vals[4] = vals[5] = vals[6] = vals[7] = 0.0f;
    }
    static inline V3 sampleFrac(int s) { return 0.5f * kLattice.sampleRel[s] + 0.5f; }
    // calculateVerticesInfo<19> with an interpolation mask (TrianglesInfluence.h:952-996)
    void sampleMid(VertexCache& cache, ContNode& n, uint32_t interpolateMask) {
        for (int s = 0; s < 19; s++) {
            if (interpolateMask & (1u << (18 - s))) tricubicVertexValues(n.coeff.data(), sampleFrac(s), 2.0f * n.half, n.mid[size_t(s)]);
            else trueSample(cache, n.center + kLattice.sampleRel[s] * n.half, n.mid[size_t(s)]);
        }
    }

    void gridPos(V3 center, int pos[3]) const {   // :282, :389
        V3 f = (center - out.box.mn) / out.cellSize;
        pos[0] = int(std::floor(f.x)); pos[1] = int(std::floor(f.y)); pos[2] = int(std::floor(f.z));
    }
    inline bool inGrid(const int p[3]) const {
        const int G = out.startGridSize;
        return p[0] >= 0 && p[0] < G && p[1] >= 0 && p[1] < G && p[2] >= 0 && p[2] < G;
    }
    inline uint32_t gridSlot(const int p[3]) const {
        const int G = out.startGridSize;
        return uint32_t(p[2] * G * G + p[1] * G + p[0]);
    }
    uint32_t wordOf(const ContNode& n, uint32_t depth) const {
        if (depth > startDepth) return n.parentBlock + uint32_t(n.path & 7u);
        int pos[3];
        gridPos(n.center, pos);
        return gridSlot(pos);
    }

    // the 8 children of `n` (depth `depth`): inherited lattice values (:532-707) and neighbour links
    // (getNeighboursVector / getNeighboursVectorInUniformGrid, OctreeSdfBreadthFirst.h:47-89)
    void makeChildren(const ContNode& n, uint32_t depth, uint32_t childBlock, std::vector<ContNode>& dst,
                      std::vector<uint32_t>* dstDepth) {
        const float h = 0.5f * n.half;
        const uint32_t c = uint32_t(n.path & 7u);
        int pos[3] = {0, 0, 0};
        if (depth == startDepth) gridPos(n.center, pos);
        for (uint32_t o = 0; o < 8; o++) {
            ContNode ch;
            ch.parentBlock = childBlock;
            ch.path = (n.path << 3) | o;
            ch.center = n.center + cornerDir(o) * h;   // center + vec3(±h)
            ch.half = h;
            for (uint32_t k = 0; k < 8; k++) {
                const int L = int((o & 1) + (k & 1)) + 3 * int(((o >> 1) & 1) + ((k >> 1) & 1)) + 9 * int((o >> 2) + (k >> 2));
                ch.values[k] = (kLattice.cornerOfLattice[L] >= 0) ? n.values[size_t(kLattice.cornerOfLattice[L])]
                                                                  : n.mid[size_t(kLattice.sampleOfLattice[L])];
            }
            if (depth == startDepth) {
                for (uint32_t m = 1; m <= 6; m++) {
                    int q[3];
                    for (int a = 0; a < 3; a++) q[a] = pos[a] + ((m & (1u << a)) ? ((o & (1u << a)) ? 1 : -1) : 0);
                    ch.link[m - 1] = inGrid(q) ? gridSlot(q) : B30;
                    ch.linkDepth[m - 1] = uint8_t(depth);
                }
            } else {
                for (uint32_t m = 1; m <= 6; m++) {
                    const uint32_t same = (~(o ^ c)) & m;   // axes on which the child also leaves n's parent
                    if (same != 0) {
                        const uint32_t pl = n.link[same - 1];
                        ch.link[m - 1] = pl + (m ^ c) * (1u - (pl >> 31));
                        ch.linkDepth[m - 1] = n.linkDepth[same - 1];
                    } else {
                        ch.link[m - 1] = n.parentBlock + (m ^ c);
                        ch.linkDepth[m - 1] = uint8_t(depth);
                    }
                }
            }
            dst.push_back(ch);
            if (dstDepth) dstDepth->push_back(depth + 1);
        }
    }

    void run() {
        buildMaskTable();
        std::vector<uint32_t>& oct = out.octree;
        const float sqThr = p0 * p0;
        const uint32_t d0 = std::min(startDepth, 1u);
        const uint32_t G = uint32_t(out.startGridSize);
        oct.assign(size_t(G) * G * G, 0u);
        std::vector<std::vector<ContNode>> level(maxDepth + 2);
        {
            const float h0 = float(0.5f * out.box.size().x * std::pow(0.5f, d0));
            const V3 c0 = out.box.mn + h0;
            const uint32_t per = 1u << d0;
            for (uint32_t k = 0; k < per; k++)
                for (uint32_t j = 0; j < per; j++)
                    for (uint32_t i = 0; i < per; i++) {
                        ContNode n;
                        n.center = c0 + v3(float(i), float(j), float(k)) * 2.0f * h0;
                        n.half = h0;
                        for (uint32_t c = 0; c < 8; c++) trueSample(cacheMain, n.center + cornerDir(c) * n.half, n.values[c]);
                        level[d0].push_back(n);
                    }
        }
        cacheIter = cacheMain;
        std::vector<uint32_t> toSubdivide;
        std::map<uint32_t, std::pair<uint32_t, uint32_t>> leaves;   // word -> (depth, index in level[depth])
        float valueRange = 0.0f;   // the reference leaves mValueRange uninitialised on this path (OctreeSdf.h:251)

        for (uint32_t depth = d0; depth <= maxDepth; depth++) {
            // ---- Iter 1 (:258-369): advance links, fit, sample, decide ------------------------------
            if (depth < maxDepth) {
                for (size_t id = 0; id < level[depth].size(); id++) {
                    ContNode& n = level[depth][id];
                    if (n.ignore) continue;
                    if (depth > startDepth) {
                        for (uint32_t m = 1; m <= 6; m++) {
                            uint32_t& l = n.link[m - 1];
                            if (((l >> 30) & 1u) != 0) continue;
                            if (isLeaf(l & ~B31)) { l |= B31; continue; }
                            l = childrenOf(l & ~B31);
                            n.linkDepth[m - 1]++;
                            while (n.linkDepth[m - 1] < depth) {
                                const uint32_t diff = depth - n.linkDepth[m - 1];
                                const uint32_t cid = uint32_t((n.path >> (3 * diff)) & 7u);
                                l += (m ^ cid);
                                if (isLeaf(l & ~B31)) { l |= B31; break; }
                                l = childrenOf(l & ~B31);
                                n.linkDepth[m - 1]++;
                            }
                        }
                    }
                    if (depth >= startDepth) tricubicCoefficients(n.values, 2.0f * n.half, n.coeff);
                    sampleMid(cacheIter, n, 0u);
                    bool terminal = false;
                    if (depth >= startDepth) terminal = errorByRule(rule, n.coeff, n.mid, p1) < sqThr;
                    n.terminal = terminal;
                    if (depth >= startDepth) setWord(wordOf(n, depth), terminal, 0xFFFFFFFFu);
                }
            }
            // ---- Iter 2 (:372-734): serial; T-junction samples, children / leaf emission -------------
            toSubdivide.clear();
            for (size_t id = 0; id < level[depth].size(); id++) {
                if (level[depth][id].ignore) continue;
                const uint32_t word = depth >= startDepth ? wordOf(level[depth][id], depth) : 0xFFFFFFFFu;
                if (!level[depth][id].terminal && depth < maxDepth) {
                    ContNode& n = level[depth][id];
                    const uint32_t c = uint32_t(n.path & 7u);
                    uint32_t samplesMask = 0;
                    uint32_t neighbourWord[24];
                    for (int i = 0; i < 24; i++) neighbourWord[i] = 0xFFFFFFFFu;
                    if (depth > startDepth) {
                        for (uint32_t dir = 1; dir <= 6; dir++)
                            for (uint32_t outward = 0; outward <= dir; outward++) {
                                if ((outward & dir) != outward) continue;   // subsets of dir
                                const uint32_t ptr = outward ? n.link[outward - 1] : n.parentBlock;
                                const uint32_t entry = 4 * (dir - 1) + signOf(dir, outward, c);
                                if (ptr >> 31) { neighbourWord[entry] = ptr & ~B31; samplesMask |= faceMask[entry]; }
                                else if (!(ptr >> 30) && isLeaf(ptr + (dir ^ c))) {
                                    neighbourWord[entry] = ptr + (dir ^ c);
                                    samplesMask |= faceMask[entry];
                                }
                            }
                    } else if (depth == startDepth) {
                        int pos[3];
                        gridPos(n.center, pos);
                        for (uint32_t dir = 1; dir <= 6; dir++) {
                            const uint32_t nAxes = (dir & 1) + ((dir >> 1) & 1) + ((dir >> 2) & 1);
                            for (uint32_t sign = 0; sign < (1u << nAxes); sign++) {
                                int q[3];
                                for (int a = 0; a < 3; a++) q[a] = pos[a] + ((dir & (1u << a)) ? axisSign(dir, sign, a) : 0);
                                if (inGrid(q) && isLeaf(gridSlot(q))) {
                                    neighbourWord[4 * (dir - 1) + sign] = gridSlot(q);
                                    samplesMask |= faceMask[4 * (dir - 1) + sign];
                                }
                            }
                        }
                    }
                    uint32_t subdivisionMask = 0;
                    for (int s = 0; s < 19; s++)
                        if (samplesMask & (1u << (18 - s))) {
                            const float inter = tricubicValue(n.coeff.data(), sampleFrac(s));
                            const float d = n.mid[size_t(s)][0] - inter;
                            if (d * d > sqThr) subdivisionMask |= samplesMask & (1u << (18 - s));
                            else tricubicVertexValues(n.coeff.data(), sampleFrac(s), 2.0f * n.half, n.mid[size_t(s)]);
                        }
                    for (int i = 0; i < 24; i++)
                        if ((subdivisionMask & faceMask[i]) && !(neighbourWord[i] >> 30)) toSubdivide.push_back(neighbourWord[i]);

                    uint32_t childBlock = 0xFFFFFFFFu;
                    if (depth >= startDepth) {
                        childBlock = uint32_t(oct.size());
                        setWord(word, false, childBlock);
                        oct.resize(oct.size() + 8, ~(7u << 29));
                    }
                    const ContNode parent = n;   // level[depth + 1] may alias nothing, but keep a stable copy
                    makeChildren(parent, depth, childBlock, level[depth + 1], nullptr);
                } else {
                    ContNode& n = level[depth][id];
                    const uint32_t at = uint32_t(oct.size());
                    setWord(word, true, at);
                    oct.resize(oct.size() + 64);
                    if (depth >= maxDepth) tricubicCoefficients(n.values, 2.0f * n.half, n.coeff);
                    std::memcpy(&oct[at], n.coeff.data(), 64 * sizeof(float));
                    for (int i = 0; i < 8; i++) valueRange = fmax_(valueRange, fabs_(n.values[size_t(i)][0]));
                    leaves.insert(std::make_pair(word, std::make_pair(depth, uint32_t(id))));
                }
            }
            // ---- fix-up (:741-1181): re-open queued coarser leaves ---------------------------------
            for (size_t q = 0; q < toSubdivide.size(); q++) {
                auto it = leaves.find(toSubdivide[q]);
                if (it == leaves.end()) continue;   // the reference prints "Leaf data not found" and is undefined here
                std::vector<ContNode> queue;
                std::vector<uint32_t> queueDepth;
                queue.push_back(level[it->second.first][it->second.second]);
                queueDepth.push_back(it->second.first);
                const uint32_t rootWord = wordOf(queue[0], queueDepth[0]);
                if (!isLeaf(rootWord)) continue;
                bool recycled = false;
                const uint32_t oldCoefficients = childrenOf(rootWord);
                bool first = true;
                for (size_t qi = 0; qi < queue.size(); qi++) {
                    ContNode n = queue[qi];
                    const uint32_t nd = queueDepth[qi];
                    const uint32_t c = uint32_t(n.path & 7u);
                    const uint32_t word = wordOf(n, nd);
                    uint32_t subdivided = 0;
                    if (nd > startDepth) {
                        for (uint32_t m = 1; m <= 6; m++) {
                            uint32_t& l = n.link[m - 1];
                            if (((l >> 30) & 1u) != 0) continue;
                            const bool follow = !first || (l >> 31);
                            if (follow && isLeaf(l & ~B31)) { l |= B31; continue; }
                            if (follow) { l = childrenOf(l & ~B31); n.linkDepth[m - 1]++; }
                            while (n.linkDepth[m - 1] < nd && n.linkDepth[m - 1] < depth) {
                                const uint32_t diff = nd - n.linkDepth[m - 1];
                                const uint32_t cid = uint32_t((n.path >> (3 * diff)) & 7u);
                                l += (m ^ cid);
                                if (isLeaf(l & ~B31)) { l |= B31; break; }
                                l = childrenOf(l & ~B31);
                                n.linkDepth[m - 1]++;
                            }
                            if (depth >= nd && !(l >> 31)) {
                                const uint32_t nw = (l & ~B31) + (m ^ c);
                                if (!(isLeaf(nw) || isMarked(nw))) subdivided |= faceMask[4 * (m - 1) + signOf(m, m, c)];
                            }
                        }
                    }
                    bool split = false;
                    uint32_t interpolateMask = 0;
                    if (depth >= nd) {
                        if (nd > startDepth) {
                            for (uint32_t dir = 1; dir <= 6; dir++)
                                for (uint32_t outward = 0; outward < dir; outward++) {
                                    if ((outward & dir) != outward) continue;   // proper subsets of dir
                                    const uint32_t ptr = outward ? n.link[outward - 1] : n.parentBlock;
                                    const bool closed = (ptr >> 31) || (ptr >> 30) || isLeaf(ptr + (dir ^ c)) || isMarked(ptr + (dir ^ c));
                                    if (!closed) subdivided |= faceMask[4 * (dir - 1) + signOf(dir, outward, c)];
                                }
                        } else if (nd == startDepth) {
                            int pos[3];
                            gridPos(n.center, pos);
                            for (uint32_t dir = 1; dir <= 6; dir++) {
                                const uint32_t nAxes = (dir & 1) + ((dir >> 1) & 1) + ((dir >> 2) & 1);
                                for (uint32_t sign = 0; sign < (1u << nAxes); sign++) {
                                    int p[3];
                                    for (int a = 0; a < 3; a++) p[a] = pos[a] + ((dir & (1u << a)) ? axisSign(dir, sign, a) : 0);
                                    if (inGrid(p) && !(isLeaf(gridSlot(p)) || isMarked(gridSlot(p)))) subdivided |= faceMask[4 * (dir - 1) + sign];
                                }
                            }
                        }
                        interpolateMask = ~subdivided;
                        split = interpolateMask != ~0u;
                    }
                    if (split) {
                        const bool recycleMid = first && !n.ignore;
                        if (!recycleMid) {
                            tricubicCoefficients(n.values, 2.0f * n.half, n.coeff);
                            sampleMid(cacheMain, n, interpolateMask);
                        }
                        for (int s = 0; s < 19; s++) {
                            if ((interpolateMask & (1u << (18 - s))) == 0) {
                                const float inter = tricubicValue(n.coeff.data(), sampleFrac(s));
                                const float d = n.mid[size_t(s)][0] - inter;
                                if (d * d < sqThr) tricubicVertexValues(n.coeff.data(), sampleFrac(s), 2.0f * n.half, n.mid[size_t(s)]);
                            } else if (recycleMid) {
                                tricubicVertexValues(n.coeff.data(), sampleFrac(s), 2.0f * n.half, n.mid[size_t(s)]);
                            }
                        }
                        const uint32_t childBlock = uint32_t(oct.size());
                        setWord(word, false, childBlock);
                        oct[word] |= MARK;
                        oct.resize(oct.size() + 8, B31);   // setValues(true, 0)
                        makeChildren(n, nd, childBlock, queue, &queueDepth);
                    } else {
                        uint32_t at = uint32_t(oct.size());
                        if (recycled) { setWord(word, true, at); oct.resize(oct.size() + 64); }
                        else { at = oldCoefficients; setWord(word, true, at); recycled = true; }
                        tricubicCoefficients(n.values, 2.0f * n.half, n.coeff);
                        std::memcpy(&oct[at], n.coeff.data(), 64 * sizeof(float));
                        n.terminal = true;
                        n.ignore = true;
                        level[nd].push_back(n);
                        leaves.insert(std::make_pair(n.parentBlock + c, std::make_pair(nd, uint32_t(level[nd].size() - 1))));
                    }
                    first = false;
                }
            }
        }
        // final un-mark (:1191-1217): every word reachable from the start grid
        {
            std::vector<uint32_t> st;
            for (uint32_t v = 0; v < G * G * G; v++) st.push_back(v);
            while (!st.empty()) {
                const uint32_t w = st.back();
                st.pop_back();
                oct[w] &= ~MARK;
                if (!(oct[w] & B31)) for (uint32_t i = 0; i < 8; i++) st.push_back((oct[w] & OCT_INDEX_MASK) + i);
            }
        }
        out.valueRange = valueRange;
    }
};

// a19: computeMinBorderValue (src/sdf/OctreeSdf.cpp:155-230)
float minBorderRec(const OrcSdf& s, uint32_t slot, V3 pos, float half) {
    const uint32_t w = s.octree[slot];
    float best = INFINITY;
    if (!(w & LEAF_BIT)) {
        for (uint32_t i = 0; i < 8; i++) {
            const V3 cp = pos + 0.5f * half * cornerDir(i);
            if (cp.x < half || cp.y < half || cp.z < half || cp.x > (1.0f - half) || cp.y > (1.0f - half) || cp.z > (1.0f - half))
                best = fmin_(best, minBorderRec(s, (w & OCT_INDEX_MASK) + i, cp, 0.5f * half));
        }
    } else {
        const float* coeff = reinterpret_cast<const float*>(&s.octree[w & OCT_INDEX_MASK]);
        for (uint32_t i = 0; i < 8; i++) {
            const V3 sp = pos + half * cornerDir(i);
            if (sp.x < 1e-4 || sp.y < 1e-4 || sp.z < 1e-4 || sp.x > (1.0f - 1e-4) || sp.y > (1.0f - 1e-4) || sp.z > (1.0f - 1e-4))
                best = fmin_(best, tricubicValue(coeff, 0.5f * cornerDir(i) + v3(0.5f)));
        }
    }
    return best;
}

void computeMinBorder(OrcSdf& s) {
    const uint32_t G = uint32_t(s.startGridSize);
    const float cell = 1.0f / float(G);
    float best = INFINITY;
    for (uint32_t k = 0; k < G; k++)
        for (uint32_t j = 0; j < G; j++)
            for (uint32_t i = 0; i < G; i++)
                best = fmin_(best, minBorderRec(s, k * G * G + j * G + i,
                                                v3((float(i) + 0.5f) * cell, (float(j) + 0.5f) * cell, (float(k) + 0.5f) * cell),
                                                0.5f * cell));
    s.minBorderValue = best;
}

// ------------------------------------------------------------------------------------------
// a9: Frank-Wolfe proximity test (src/utils/GJK.cpp:830-866, :715-738, :644-652)
// ------------------------------------------------------------------------------------------
bool isNearMinimize(float half, const float radius[8], const V3 tri[3], float thr, uint32_t* outIter) {
    uint32_t iter = 0;
    float distToP, distToO;
    bool isNear = false;
    const float sqThr = thr * thr;
    V3 x = -tri[0];
    bool result;
    for (;;) {
        const V3 g = normalize(-x);
        // support of the 8 corner spheres along g (first strict maximum)
        float best = dot(v3(-half), g) + radius[0];
        uint32_t bi = 0;
        for (uint32_t i = 1; i < 8; i++) {
            const float v = dot(cornerDir(i) * half, g) + radius[i];
            if (v > best) { best = v; bi = i; }
        }
        const V3 boxPoint = cornerDir(bi) * half + radius[bi] * g;
        // support of the triangle along -g
        const V3 ng = -g;
        const float d1 = dot(tri[0], ng), d2 = dot(tri[1], ng), d3 = dot(tri[2], ng);
        const V3 triPoint = (d1 > d2) ? ((d1 > d3) ? tri[0] : tri[2]) : ((d2 > d3) ? tri[1] : tri[2]);
        const V3 p = boxPoint - triPoint;
        distToP = dot(g, p - x);
        distToO = dot(g, -x);
        const V3 dir = p - x;
        const float d = dot(dir, -x);
        if (double(d) < 1.0e-5) { result = distToO <= distToP + thr; if (outIter) *outIter = iter; return result; }
        x = x + dir * fmin_(d / dot(dir, dir), 1.0f);
        isNear = dot(x, x) < sqThr;
        if (!(!isNear && distToO <= distToP + thr && ++iter < 15)) break;
    }
    if (outIter) *outIter = iter;
    return isNear || iter >= 15;
}

// a8: PerNodeRegionTrianglesInfluence::filterTriangles (TrianglesInfluence.h:767-860)
void filterTriangles(const MeshView& mesh, const std::vector<TriData>& tris, V3 center, float half,
                     const std::vector<uint32_t>& in, const uint32_t cornerTri[8], std::vector<uint32_t>& out) {
    out.clear();
    float region[8][8], minDist[8];
    for (uint32_t i = 0; i < 8; i++) {
        minDist[i] = INFINITY;
        for (uint32_t c = 0; c < 8; c++) {
            region[i][c] = std::sqrt(sqDistPointTriangle(center + cornerDir(c) * half, tris[cornerTri[i]]));
            minDist[i] = fmin_(minDist[i], region[i][c]);
        }
        for (uint32_t c = 0; c < 8; c++) region[i][c] -= minDist[i];
    }
    for (uint32_t t : in) {
        V3 tri[3] = {mesh.verts[mesh.idx[3 * t]] - center, mesh.verts[mesh.idx[3 * t + 1]] - center,
                     mesh.verts[mesh.idx[3 * t + 2]] - center};
        const V3 g = 0.3333333f * (tri[0] + tri[1] + tri[2]);
        const uint32_t v = ((g.z > 0) ? 4u : 0u) + ((g.y > 0) ? 2u : 0u) + ((g.x > 0) ? 1u : 0u);
        if (cornerTri[v] == t || isNearMinimize(half, region[v], tri, minDist[v], nullptr)) out.push_back(t);
    }
}

// ------------------------------------------------------------------------------------------
// a7 + a16: ExactOctreeSdf depth-first build (include/SdfLib/ExactOctreeSdfDepthFirst.h:28-651)
// ------------------------------------------------------------------------------------------
struct ExactBuilder {
    OrcSdf& out;
    MeshView mesh;
    VertexCache cache;
    uint32_t startDepth, maxDepth, minTris, bitEnc, bits;
    std::vector<uint32_t>* nodesOut; std::vector<uint32_t>* setsOut; std::vector<uint8_t>* masksOut;

    // PerNodeRegionTrianglesInfluence::calculateVerticesInfo (TrianglesInfluence.h:693-765):
    // cache hit, else first strict minimum over the (ascending) list.
    uint32_t nearestInList(V3 p, const std::vector<uint32_t>& list, uint32_t stale) {
        uint32_t slot, key[3], tri = stale;
        if (cache.lookup(p, slot, key, tri)) return tri;
        float best = INFINITY;
        tri = stale;   // the reference leaves the output untouched when the list is empty
        for (uint32_t t : list) {
            const float d = sqDistPointTriangle(p, out.tris[t]);
            if (d < best) { tri = t; best = d; }
        }
        cache.store(slot, key, tri);
        return tri;
    }

    void appendSet(const std::vector<uint32_t>& list, std::vector<uint32_t>& sets) {   // :263-280, :450-467
        const uint32_t n = uint32_t(list.size());
        const uint32_t words = (n * bits + 31) / 32;
        size_t at = sets.size();
        sets.resize(sets.size() + words + 2, 0u);
        sets[at++] = n;
        uint32_t bIdx = 0;
        for (uint32_t t = 0; t < n; t++, bIdx += bits) {
            const uint32_t index = list[t], w = bIdx >> 5, bit = bIdx & 31u;
            sets[at + w] |= (index << (32 - bits)) >> bit;
            sets[at + w + 1] |= uint32_t(uint64_t(index) << (64 - (bit + bits)));
        }
    }

    struct Node { uint32_t nodeIndex, depth; V3 center; float half; uint32_t info[8]; };

    // One node, both visits of the reference folded into a recursion with the same append order:
    // first visit (:288-478) -> children 7..0 -> second visit (:189-286). Returns the node's final list.
    void process(const Node& n, const std::vector<uint32_t>& parentList, std::vector<uint32_t>& list) {
        std::vector<uint32_t>& oct = *nodesOut;
        filterTriangles(mesh, out.tris, n.center, n.half, parentList, n.info, list);
        bool terminal = false;
        if (n.depth >= startDepth) terminal = list.size() <= minTris;
        if (!terminal && n.depth < maxDepth) {
            uint32_t midInfo[19];
            for (int s = 0; s < 19; s++) {
                midInfo[s] = 0;   // deterministic stand-in for the reference's uninitialised slot on an empty list
                midInfo[s] = nearestInList(n.center + kLattice.sampleRel[s] * n.half, list, midInfo[s]);
            }
            const float h = 0.5f * n.half;
            const bool real = n.depth >= startDepth;
            const uint32_t childIndex = real ? uint32_t(oct.size() / 2) : 0xFFFFFFFFu;
            if (n.nodeIndex != 0xFFFFFFFFu) oct[2 * n.nodeIndex] = childIndex & EXACT_INDEX_MASK;
            if (real) oct.resize(oct.size() + 16, 0u);
            std::array<std::vector<uint32_t>, 8> childLists;
            for (int c = 7; c >= 0; c--) {
                Node ch;
                ch.nodeIndex = real ? childIndex + uint32_t(c) : 0xFFFFFFFFu;
                ch.depth = n.depth + 1;
                ch.center = n.center + cornerDir(uint32_t(c)) * h;
                ch.half = h;
                for (uint32_t k = 0; k < 8; k++) {
                    const int L = int((c & 1) + (k & 1)) + 3 * int(((c >> 1) & 1) + ((k >> 1) & 1)) + 9 * int((c >> 2) + (k >> 2));
                    ch.info[k] = kLattice.cornerOfLattice[L] >= 0 ? n.info[kLattice.cornerOfLattice[L]]
                                                                  : midInfo[kLattice.sampleOfLattice[L]];
                }
                if (ch.depth == startDepth) ch.nodeIndex = startSlot(out, ch.center);
                process(ch, list, childLists[size_t(c)]);
            }
            if (n.depth >= bitEnc) {   // second visit: 8-way merge, masks, (at bitEnc) the union set
                const size_t preMergeBytes = (list.size() + 7) / 8;
                std::array<std::vector<uint8_t>, 8> maskOf;
                for (auto& mk : maskOf) mk.assign(preMergeBytes, 0);
                list.clear();
                size_t pos[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (;;) {
                    uint32_t mn = 0xFFFFFFFFu;
                    bool any = false;
                    for (int c = 0; c < 8; c++)
                        if (pos[c] < childLists[size_t(c)].size()) { mn = std::min(mn, childLists[size_t(c)][pos[c]]); any = true; }
                    if (!any) break;
                    const size_t j = list.size();
                    list.push_back(mn);
                    for (int c = 0; c < 8; c++)
                        if (pos[c] < childLists[size_t(c)].size() && childLists[size_t(c)][pos[c]] == mn) {
                            pos[c]++;
                            maskOf[size_t(c)][j / 8] |= uint8_t(1u << (7 - (j & 7)));
                        }
                }
                const size_t bytes = (list.size() + 7) / 8;
                const uint32_t firstChild = oct[2 * n.nodeIndex] & EXACT_INDEX_MASK;
                for (uint32_t c = 0; c < 8; c++) {
                    oct[2 * (firstChild + c) + 1] = uint32_t(masksOut->size());
                    masksOut->insert(masksOut->end(), maskOf[c].begin(), maskOf[c].begin() + long(bytes));
                }
                if (n.depth == bitEnc) {
                    oct[2 * n.nodeIndex + 1] = uint32_t(setsOut->size());
                    appendSet(list, *setsOut);
                    out.maxTrisEncoded = std::max(out.maxTrisEncoded, uint32_t(list.size()));
                }
            }
        } else {   // leaf (:442-478)
            oct[2 * n.nodeIndex] = 0xFFFFFFFFu;
            if (n.depth <= bitEnc) {
                oct[2 * n.nodeIndex + 1] = uint32_t(setsOut->size());
                appendSet(list, *setsOut);
            }
            out.maxTrisInLeafs = std::max(out.maxTrisInLeafs, uint32_t(list.size()));
        }
    }

    void run() {
        const uint32_t d0 = std::min(startDepth, 1u);
        std::vector<uint32_t> all;
        for (uint32_t t = 0; t < out.tris.size(); t++) {
            V3 nrm = triNormal(out.tris[t]);
            if (dot(nrm, nrm) > 1e-3f) all.push_back(t);
        }
        const float h0 = float(0.5f * out.box.size().x * std::pow(0.5f, d0));
        const V3 c0 = out.box.mn + h0;
        const uint32_t per = 1u << d0;
        std::vector<Node> seeds;
        for (uint32_t k = 0; k < per; k++)
            for (uint32_t j = 0; j < per; j++)
                for (uint32_t i = 0; i < per; i++) {
                    Node n;
                    n.nodeIndex = 0xFFFFFFFFu; n.depth = d0; n.half = h0;
                    n.center = c0 + v3(float(i), float(j), float(k)) * 2.0f * h0;
                    for (uint32_t c = 0; c < 8; c++) n.info[c] = nearestInList(n.center + cornerDir(c) * n.half, all, 0);
                    seeds.push_back(n);
                }
        const uint32_t G = uint32_t(out.startGridSize);
        out.nodes.assign(size_t(2) * G * G * G, 0u);
        nodesOut = &out.nodes; setsOut = &out.sets; masksOut = &out.masks;
        for (size_t i = seeds.size(); i-- > 0;) {   // the stack pops the last-pushed seed first
            Node n = seeds[i];
            if (n.depth == startDepth) n.nodeIndex = startSlot(out, n.center);
            std::vector<uint32_t> list;
            process(n, all, list);
        }
    }
};

// ------------------------------------------------------------------------------------------
// a18 / a21: queries
// ------------------------------------------------------------------------------------------
inline bool locateStart(const OrcSdf& s, V3 p, V3& frac, uint32_t& slot) {
    V3 f = (p - s.box.mn) / s.cellSize;
    const int x = int(std::floor(f.x)), y = int(std::floor(f.y)), z = int(std::floor(f.z));
    frac = V3{fract_(f.x), fract_(f.y), fract_(f.z)};
    const int G = s.startGridSize;
    if (x < 0 || x >= G || y < 0 || y >= G || z < 0 || z >= G) return false;
    slot = uint32_t(z * G * G + y * G + x);
    return true;
}

float queryOctree(const OrcSdf& s, V3 p, V3* grad) {   // src/sdf/OctreeSdf.cpp:93-152
    V3 frac; uint32_t slot;
    if (!locateStart(s, p, frac, slot)) {
        if (grad) return s.box.distance(p, *grad) + s.minBorderValue;
        return s.box.distance(p) + s.minBorderValue;
    }
    uint32_t w = s.octree[slot];
    while (!(w & LEAF_BIT)) {
        const uint32_t child = ((frac.z >= 0.5f) ? 4u : 0u) + ((frac.y >= 0.5f) ? 2u : 0u) + ((frac.x >= 0.5f) ? 1u : 0u);
        w = s.octree[(w & OCT_INDEX_MASK) + child];
        frac = V3{fract_(2.0f * frac.x), fract_(2.0f * frac.y), fract_(2.0f * frac.z)};
    }
    const float* c = reinterpret_cast<const float*>(&s.octree[w & OCT_INDEX_MASK]);
    if (grad) *grad = normalize(tricubicGradient(c, frac));
    return tricubicValue(c, frac);
}

inline uint32_t unpackIndex(const uint32_t* words, uint32_t bIdx, uint32_t bits) {   // ExactOctreeSdf.cpp:73-77
    const uint32_t w = bIdx >> 5, bit = bIdx & 31u;
    return ((words[w] << bit) >> (32 - bits)) | uint32_t(uint64_t(words[w + 1]) >> (64 - (bit + bits)));
}

float queryExact(const OrcSdf& s, V3 p, V3* grad, std::vector<uint32_t> scratch[2]) {   // src/sdf/ExactOctreeSdf.cpp:38-320
    V3 frac; uint32_t slot;
    if (!locateStart(s, p, frac, slot)) return s.box.distance(p) + std::sqrt(3.0f) * s.box.size().x;
    auto childOf = [&](V3 f) { return ((f.z > 0.5f) ? 4u : 0u) + ((f.y > 0.5f) ? 2u : 0u) + ((f.x > 0.5f) ? 1u : 0u); };
    auto halve = [&](V3 f) { return V3{fract_(2.0f * f.x), fract_(2.0f * f.y), fract_(2.0f * f.z)}; };
    const uint32_t* nd = &s.nodes[2 * slot];
    uint32_t depth = s.startDepth;
    while (!(nd[0] & LEAF_BIT) && depth < s.bitEncodingStartDepth) {
        nd = &s.nodes[2 * ((nd[0] & EXACT_INDEX_MASK) + childOf(frac))];
        frac = halve(frac);
        depth++;
    }
    float best = INFINITY;
    uint32_t bestTri = 0;
    auto finish = [&]() { return grad ? signedDistGradSelf(p, s.tris[bestTri], *grad) : signedDistPointTriangle(p, s.tris[bestTri]); };
    if (nd[0] & LEAF_BIT) {
        const uint32_t* set = &s.sets[nd[1]];
        const uint32_t n = set[0];
        for (uint32_t t = 0, b = 0; t < n; t++, b += s.bitsPerIndex) {
            const uint32_t tri = unpackIndex(set + 1, b, s.bitsPerIndex);
            const float d = sqDistPointTriangle(p, s.tris[tri]);
            if (d < best) { bestTri = tri; best = d; }
        }
        return finish();
    }
    const uint32_t* set = &s.sets[nd[1]];
    nd = &s.nodes[2 * ((nd[0] & EXACT_INDEX_MASK) + childOf(frac))];
    frac = halve(frac);
    uint32_t n = set[0];
    uint32_t* in = scratch[0].data();
    uint32_t* outp = scratch[1].data();
    {
        const uint8_t* mask = s.masks.data() + nd[1];
        uint32_t kept = 0;
        for (uint32_t t = 0; t < n; t++)
            if (mask[t >> 3] & (0x80u >> (t & 7))) in[kept++] = unpackIndex(set + 1, t * s.bitsPerIndex, s.bitsPerIndex);
        n = kept;
    }
    while (!(nd[0] & LEAF_BIT)) {
        nd = &s.nodes[2 * ((nd[0] & EXACT_INDEX_MASK) + childOf(frac))];
        frac = halve(frac);
        const uint8_t* mask = s.masks.data() + nd[1];
        uint32_t kept = 0;
        for (uint32_t t = 0; t < n; t++)
            if (mask[t >> 3] & (0x80u >> (t & 7))) outp[kept++] = in[t];
        n = kept;
        std::swap(in, outp);
    }
    for (uint32_t t = 0; t < n; t++) {
        const float d = sqDistPointTriangle(p, s.tris[in[t]]);
        if (d < best) { bestTri = in[t]; best = d; }
    }
    return finish();
}

// ------------------------------------------------------------------------------------------
// f-1: .bin layout = cereal PortableBinary, little endian (src/sdf/SdfFunction.cpp:9-79; field order
// OctreeSdf.h:225, ExactOctreeSdf.h:141, TriangleUtils.h:53, Mesh.h:68)
// ------------------------------------------------------------------------------------------
template <class T> void put(std::ostream& os, const T& v) { os.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <class T> bool get(std::istream& is, T& v) { is.read(reinterpret_cast<char*>(&v), sizeof(T)); return bool(is); }
template <class T> void putVec(std::ostream& os, const T* p, uint64_t n) {
    put(os, n);
    os.write(reinterpret_cast<const char*>(p), std::streamsize(n * sizeof(T)));
}

}  // namespace

// ==========================================================================================
// C API
// ==========================================================================================
extern "C" {

void orc_isosphere(uint32_t subdivisions, float* outVerts, uint32_t* outIdx, uint32_t* nVerts, uint32_t* nIdx) {
    // src/utils/PrimitivesFactory.cpp:19-104: icosahedron + midpoint subdivision; the centre triangle
    // replaces the parent in place, the three corner triangles are appended; midpoints are shared
    // through a per-level edge map and pushed onto the unit sphere with normalize().
    const float X = 0.525731112119133606f, Z = 0.850650808352039932f;
    std::vector<V3> v = {{-X, 0, Z}, {X, 0, Z}, {-X, 0, -Z}, {X, 0, -Z}, {0, Z, X}, {0, Z, -X},
                         {0, -Z, X}, {0, -Z, -X}, {Z, X, 0}, {-Z, X, 0}, {Z, -X, 0}, {-Z, -X, 0}};
    std::vector<uint32_t> f = {0, 4, 1, 0, 9, 4, 9, 5, 4, 4, 5, 8, 4, 8, 1, 8, 10, 1, 8, 3, 10, 5, 3, 8, 5, 2, 3, 2, 7, 3,
                               7, 10, 3, 7, 6, 10, 7, 11, 6, 11, 0, 6, 0, 1, 6, 6, 1, 10, 9, 0, 11, 9, 11, 2, 9, 2, 5, 7, 2, 11};
    for (uint32_t s = 0; s < subdivisions; s++) {
        std::map<std::pair<uint32_t, uint32_t>, uint32_t> mids;
        auto mid = [&](uint32_t a, uint32_t b) {
            auto key = std::make_pair(std::min(a, b), std::max(a, b));
            auto it = mids.find(key);
            if (it != mids.end()) { uint32_t r = it->second; mids.erase(it); return r; }
            v.push_back(normalize(0.5f * (v[a] + v[b])));
            mids[key] = uint32_t(v.size() - 1);
            return uint32_t(v.size() - 1);
        };
        const size_t old = f.size();
        for (size_t t = 0; t < old; t += 3) {
            const uint32_t a = f[t], b = f[t + 1], c = f[t + 2];
            const uint32_t ab = mid(a, b), bc = mid(b, c), ca = mid(c, a);
            const uint32_t add[9] = {a, ab, ca, ab, b, bc, bc, c, ca};
            f.insert(f.end(), add, add + 9);
            f[t] = ab; f[t + 1] = bc; f[t + 2] = ca;
        }
    }
    *nVerts = uint32_t(v.size());
    *nIdx = uint32_t(f.size());
    if (outVerts) std::memcpy(outVerts, v.data(), v.size() * sizeof(V3));
    if (outIdx) std::memcpy(outIdx, f.data(), f.size() * sizeof(uint32_t));
}

void orc_triangle_data(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, float* out37) {
    MeshView m{reinterpret_cast<const V3*>(verts), nVerts, idx, nIdx};
    auto td = meshTriangleData(m);
    std::memcpy(out37, td.data(), td.size() * sizeof(TriData));
}

void orc_sq_dist(const float* tri37, const float* pts, uint64_t n, float* out) {
    const TriData& d = *reinterpret_cast<const TriData*>(tri37);
    for (uint64_t i = 0; i < n; i++) out[i] = sqDistPointTriangle(v3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), d);
}

void orc_signed_dist(const float* tri37, const float* w, const float* pts, uint64_t n, int mode, float* outDist,
                     float* outGrad) {
    const TriData& d = *reinterpret_cast<const TriData*>(tri37);
    for (uint64_t i = 0; i < n; i++) {
        V3 p = v3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), g = v3(0.0f);
        if (mode == 0) outDist[i] = signedDistPointTriangle(p, d);
        else if (mode == 1) outDist[i] = signedDistGradMesh(p, d, v3(w[0], w[1], w[2]), v3(w[3], w[4], w[5]), v3(w[6], w[7], w[8]), g);
        else outDist[i] = signedDistGradSelf(p, d, g);
        if (outGrad) { outGrad[3 * i] = g.x; outGrad[3 * i + 1] = g.y; outGrad[3 * i + 2] = g.z; }
    }
}

void orc_tricubic_coefficients(const float* values8x8, float nodeSize, float* out64) {
    std::array<PointValues, 8> in;
    std::memcpy(in.data(), values8x8, sizeof(in));
    Coeffs c;
    tricubicCoefficients(in, nodeSize, c);
    std::memcpy(out64, c.data(), sizeof(c));
}

void orc_tricubic_eval(const float* coeff64, const float* frac, uint64_t n, float* outValue, float* outGrad,
                       float* outVertexValues, float nodeSize) {
    for (uint64_t i = 0; i < n; i++) {
        V3 f = v3(frac[3 * i], frac[3 * i + 1], frac[3 * i + 2]);
        if (outValue) outValue[i] = tricubicValue(coeff64, f);
        if (outGrad) { V3 g = tricubicGradient(coeff64, f); outGrad[3 * i] = g.x; outGrad[3 * i + 1] = g.y; outGrad[3 * i + 2] = g.z; }
        if (outVertexValues) { PointValues vv; tricubicVertexValues(coeff64, f, nodeSize, vv); std::memcpy(outVertexValues + 8 * i, vv.data(), sizeof(vv)); }
    }
}

float orc_error_estimate(const float* coeff64, const float* mid19x8, int rule, float decay) {
    Coeffs c;
    std::memcpy(c.data(), coeff64, sizeof(c));
    std::array<PointValues, 19> mid;
    std::memcpy(mid.data(), mid19x8, sizeof(mid));
    return errorByRule(rule, c, mid, decay);
}

int orc_is_near_minimize(float half, const float* r8, const float* tri9, float thr, uint32_t* outIter) {
    V3 t[3] = {v3(tri9[0], tri9[1], tri9[2]), v3(tri9[3], tri9[4], tri9[5]), v3(tri9[6], tri9[7], tri9[8])};
    return isNearMinimize(half, r8, t, thr, outIter) ? 1 : 0;
}

uint32_t orc_filter_triangles(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx,
                              const float* c3, float half, const uint32_t* inTris, uint32_t nIn,
                              const uint32_t* cornerTris8, uint32_t* outTris) {
    MeshView m{reinterpret_cast<const V3*>(verts), nVerts, idx, nIdx};
    auto td = meshTriangleData(m);
    std::vector<uint32_t> in(inTris, inTris + nIn), out;
    filterTriangles(m, td, v3(c3[0], c3[1], c3[2]), half, in, cornerTris8, out);
    std::memcpy(outTris, out.data(), out.size() * sizeof(uint32_t));
    return uint32_t(out.size());
}

void orc_nearest_triangle(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, const float* pts,
                          uint64_t n, uint32_t* outTri) {
    MeshView m{reinterpret_cast<const V3*>(verts), nVerts, idx, nIdx};
    Bvh bvh(m);
    for (uint64_t i = 0; i < n; i++) outTri[i] = bvh.nearest(v3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
}

// Same queries, plus how many inner nodes and leaves each traversal visited (2 counters per point): input of the
// warp-occupancy model of the device sampler, tests/model_bvh_traversal.py.
void orc_nearest_triangle_visits(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, const float* pts,
                                 uint64_t n, uint32_t* outTri, uint32_t* outVisits2) {
    MeshView m{reinterpret_cast<const V3*>(verts), nVerts, idx, nIdx};
    Bvh bvh(m);
    for (uint64_t i = 0; i < n; i++) {
        outVisits2[2 * i] = outVisits2[2 * i + 1] = 0;
        bvh.visitCount = outVisits2 + 2 * i;
        outTri[i] = bvh.nearest(v3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
    }
    bvh.visitCount = nullptr;
}

// Traversal model of a SEEDED search (tests/model_bvh_seed.py): the running best starts from seeds[i] instead of DBL_MAX.
void orc_nearest_triangle_visits_seeded(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, const float* pts,
                                        uint64_t n, const double* seeds, uint32_t* outTri, uint32_t* outVisits2) {
    MeshView m{reinterpret_cast<const V3*>(verts), nVerts, idx, nIdx};
    Bvh bvh(m);
    if (seeds == nullptr) bvh.buildBoxes();   // no seeds: the box-pruning model instead
    for (uint64_t i = 0; i < n; i++) {
        outVisits2[2 * i] = outVisits2[2 * i + 1] = 0;
        bvh.visitCount = outVisits2 + 2 * i;
        outTri[i] = bvh.nearest(v3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), seeds ? seeds[i] : std::numeric_limits<double>::max());
    }
    bvh.visitCount = nullptr;
}

OrcSdf* orc_build_octree(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, const float* box6,
                         uint32_t depth, uint32_t startDepth, int rule, float param0, float param1, int algorithm,
                         uint32_t numThreads, int useCache) {
    if (algorithm != 1 && algorithm != 2) return nullptr;   // UNIFORM ("for testing", OctreeSdf.h:286): not restated
    auto* s = new OrcSdf();
    s->format = 1;
    s->maxDepth = depth;
    cubify(*s, box6, startDepth);
    MeshView m{reinterpret_cast<const V3*>(verts), nVerts, idx, nIdx};
    if (algorithm == 2) {
        ContBuilder b{*s, m, meshTriangleData(m), nullptr, VertexCache(), VertexCache(), startDepth, depth, rule, param0, param1, {}};
        b.bvh.reset(new Bvh(m));
        b.cacheMain.init(s->box, depth, useCache != 0);
        b.run();
        computeMinBorder(*s);
        return s;
    }
    OctBuilder b{*s, m, meshTriangleData(m), nullptr, VertexCache(), startDepth, depth, rule, param0, param1};
    b.bvh.reset(new Bvh(m));
    b.cache.init(s->box, depth, useCache != 0);
    b.run(numThreads);
    computeMinBorder(*s);
    return s;
}

OrcSdf* orc_build_exact(const float* verts, uint32_t nVerts, const uint32_t* idx, uint32_t nIdx, const float* box6,
                        uint32_t maxDepth, uint32_t startDepth, uint32_t minTris, uint32_t numThreads, int useCache) {
    (void)numThreads;   // single-DFS layout only; see header comment on the reference's MT merge defect
    if (maxDepth < startDepth + 2) return nullptr;   // the reference dereferences a null node otherwise (:193)
    auto* s = new OrcSdf();
    s->format = 2;
    s->maxDepth = maxDepth;
    s->startDepth = startDepth;
    cubify(*s, box6, startDepth);
    MeshView m{reinterpret_cast<const V3*>(verts), nVerts, idx, nIdx};
    s->tris = meshTriangleData(m);
    s->minTrisInLeafs = minTris;
    s->bitEncodingStartDepth = maxDepth - 2;
    s->bitsPerIndex = uint32_t(int32_t(std::ceil(std::log2(float(s->tris.size())))));
    ExactBuilder b{*s, m, VertexCache(), startDepth, maxDepth, minTris, maxDepth - 2, s->bitsPerIndex, nullptr, nullptr, nullptr};
    b.cache.init(s->box, maxDepth, useCache != 0);
    b.run();
    return s;
}

void orc_delete(OrcSdf* s) { delete s; }
int orc_format(const OrcSdf* s) { return s->format; }

int orc_save(const OrcSdf* s, const char* path) {
    std::ofstream os(path, std::ios::out | std::ios::binary);
    if (!os.is_open()) return 0;
    put(os, uint8_t(1));
    put(os, uint32_t(s->format));
    put(os, s->box);
    put(os, int32_t(s->startGridSize));
    if (s->format == 1) {
        put(os, s->maxDepth); put(os, s->valueRange); put(os, s->minBorderValue);
        putVec(os, s->octree.data(), uint64_t(s->octree.size()));
    } else {
        put(os, s->startDepth); put(os, s->minTrisInLeafs); put(os, s->maxTrisInLeafs); put(os, s->maxTrisEncoded);
        put(os, s->bitEncodingStartDepth); put(os, s->bitsPerIndex); put(os, s->maxDepth);
        put(os, uint64_t(s->nodes.size() / 2));
        os.write(reinterpret_cast<const char*>(s->nodes.data()), std::streamsize(s->nodes.size() * 4));
        putVec(os, s->sets.data(), uint64_t(s->sets.size()));
        putVec(os, s->masks.data(), uint64_t(s->masks.size()));
        putVec(os, s->tris.data(), uint64_t(s->tris.size()));
    }
    return bool(os) ? 1 : 0;
}

OrcSdf* orc_load(const char* path) {
    std::ifstream is(path, std::ios::binary);
    if (!is.is_open()) return nullptr;
    uint8_t le; uint32_t fmt; int32_t sg; uint64_t n;
    if (!get(is, le) || !get(is, fmt) || (fmt != 1 && fmt != 2)) return nullptr;
    auto s = std::unique_ptr<OrcSdf>(new OrcSdf());
    s->format = int(fmt);
    get(is, s->box); get(is, sg);
    s->startGridSize = sg;
    if (fmt == 1) {
        get(is, s->maxDepth); get(is, s->valueRange); get(is, s->minBorderValue);
        if (!get(is, n)) return nullptr;
        s->octree.resize(n);
        is.read(reinterpret_cast<char*>(s->octree.data()), std::streamsize(n * 4));
    } else {
        get(is, s->startDepth); get(is, s->minTrisInLeafs); get(is, s->maxTrisInLeafs); get(is, s->maxTrisEncoded);
        get(is, s->bitEncodingStartDepth); get(is, s->bitsPerIndex); get(is, s->maxDepth);
        if (!get(is, n)) return nullptr;
        s->nodes.resize(2 * n);
        is.read(reinterpret_cast<char*>(s->nodes.data()), std::streamsize(n * 8));
        get(is, n); s->sets.resize(n);
        is.read(reinterpret_cast<char*>(s->sets.data()), std::streamsize(n * 4));
        get(is, n); s->masks.resize(n);
        is.read(reinterpret_cast<char*>(s->masks.data()), std::streamsize(n));
        get(is, n); s->tris.resize(n);
        is.read(reinterpret_cast<char*>(s->tris.data()), std::streamsize(n * sizeof(TriData)));
    }
    if (!is) return nullptr;
    s->cellSize = s->box.size().x / float(s->startGridSize);   // OctreeSdf.h:233, ExactOctreeSdf.h:149
    return s.release();
}

void orc_sample_area(const OrcSdf* s, float* o) {
    o[0] = s->box.mn.x; o[1] = s->box.mn.y; o[2] = s->box.mn.z; o[3] = s->box.mx.x; o[4] = s->box.mx.y; o[5] = s->box.mx.z;
}
uint64_t orc_octree_data_size(const OrcSdf* s) { return s->format == 1 ? s->octree.size() : s->nodes.size() / 2; }
void orc_octree_data(const OrcSdf* s, uint32_t* out) {
    const auto& v = s->format == 1 ? s->octree : s->nodes;
    std::memcpy(out, v.data(), v.size() * 4);
}
void orc_octree_header(const OrcSdf* s, int* sg, uint32_t* md, float* a, float* b) {
    *sg = s->startGridSize; *md = s->maxDepth;
    if (s->format == 1) { *a = s->valueRange; *b = s->minBorderValue; }
    else { *a = float(s->minTrisInLeafs); *b = float(s->maxTrisInLeafs); }
}
uint64_t orc_exact_sizes(const OrcSdf* s, uint64_t* nSets, uint64_t* nMasks, uint64_t* nTris) {
    *nSets = s->sets.size(); *nMasks = s->masks.size(); *nTris = s->tris.size();
    return s->nodes.size() / 2;
}
void orc_exact_arrays(const OrcSdf* s, uint32_t* sets, uint8_t* masks, float* tris37) {
    if (sets) std::memcpy(sets, s->sets.data(), s->sets.size() * 4);
    if (masks) std::memcpy(masks, s->masks.data(), s->masks.size());
    if (tris37) std::memcpy(tris37, s->tris.data(), s->tris.size() * sizeof(TriData));
}
void orc_exact_header(const OrcSdf* s, uint32_t* o) {
    o[0] = uint32_t(s->startGridSize); o[1] = s->startDepth; o[2] = s->minTrisInLeafs; o[3] = s->maxTrisInLeafs;
    o[4] = s->maxTrisEncoded; o[5] = s->bitEncodingStartDepth; o[6] = s->bitsPerIndex; o[7] = s->maxDepth;
}

double orc_query(const OrcSdf* s, const float* pts, uint64_t n, float* outDist, float* outGrad, int numThreads) {
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel num_threads(numThreads > 1 ? numThreads : 1)
    {
        std::vector<uint32_t> scratch[2];
        if (s->format == 2) { scratch[0].resize(s->maxTrisEncoded + 8); scratch[1].resize(s->maxTrisEncoded + 8); }
#pragma omp for schedule(static)
        for (int64_t i = 0; i < int64_t(n); i++) {
            V3 p = v3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
            V3 g = v3(0.0f);
            float d = s->format == 1 ? queryOctree(*s, p, outGrad ? &g : nullptr) : queryExact(*s, p, outGrad ? &g : nullptr, scratch);
            outDist[i] = d;
            if (outGrad) { outGrad[3 * i] = g.x; outGrad[3 * i + 1] = g.y; outGrad[3 * i + 2] = g.z; }
        }
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
