// Per-triangle part of TriangleData shared by the host path (mesh_host.cpp) and the device path (mesh_device.cu):
// the TriangleData constructor and the libm-exact acosf of the corner angles. Both translation units are compiled
// without FMA contraction, so the two paths produce the same bits (tests/test_gpu_mesh.py).
#pragma once
#include <cstring>
#include "tri_math.cuh"

namespace sdfb200 {

__host__ __device__ inline TriData makeTriData(f3 p1, f3 p2, f3 p3) {   // TriangleData ctor, include/SdfLib/utils/TriangleUtils.h:23-42
    TriData d;
    d.origin[0] = p1.x; d.origin[1] = p1.y; d.origin[2] = p1.z;
    const f3 e12 = p2 - p1, e13 = p3 - p1;
    const f3 sx = normalize3(e12);
    const f3 crs = mk3(e12.y * e13.z - e13.y * e12.z, e12.z * e13.x - e13.z * e12.x, e12.x * e13.y - e13.x * e12.y);
    const f3 sz = normalize3(crs);
    const f3 sy = mk3(sz.y * sx.z - sx.y * sz.z, sz.z * sx.x - sx.z * sz.x, sz.x * sx.y - sx.x * sz.y);
    // inverse of the matrix with columns (sx, sy, sz): cofactors times 1/det, det along the first row
    const float m[3][3] = {{sx.x, sx.y, sx.z}, {sy.x, sy.y, sy.z}, {sz.x, sz.y, sz.z}};
    const float c00 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    const float c10 = m[0][1] * m[2][2] - m[2][1] * m[0][2];
    const float c20 = m[0][1] * m[1][2] - m[1][1] * m[0][2];
    const float inv = 1.0f / (+m[0][0] * c00 - m[1][0] * c10 + m[2][0] * c20);
    d.T[0][0] = +c00 * inv;
    d.T[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * inv;
    d.T[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * inv;
    d.T[0][1] = -c10 * inv;
    d.T[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * inv;
    d.T[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * inv;
    d.T[0][2] = +c20 * inv;
    d.T[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * inv;
    d.T[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * inv;
    {
        const f3 v = matMul(d.T, p3 - p2);
        const float tx = v.x * v.x, ty = v.y * v.y;
        const float s = 1.0f / sqrtf(tx + ty);
        d.b[0] = v.x * s; d.b[1] = v.y * s;
    }
    {
        const f3 v = matMul(d.T, p1 - p3);
        const float tx = v.x * v.x, ty = v.y * v.y;
        const float s = 1.0f / sqrtf(tx + ty);
        d.c[0] = v.x * s; d.c[1] = v.y * s;
    }
    d.v2 = matMul(d.T, p2 - p1).x;
    const f3 l3 = matMul(d.T, p3 - p1);
    d.v3[0] = l3.x; d.v3[1] = l3.y;
    for (int k = 0; k < 3; k++) {
        d.edgesNormal[k][0] = 0.f; d.edgesNormal[k][1] = 0.f; d.edgesNormal[k][2] = 1.f;
        d.verticesNormal[k][0] = 0.f; d.verticesNormal[k][1] = 0.f; d.verticesNormal[k][2] = 1.f;
    }
    return d;
}

// acosf with the bits of the host libm the reference is linked against (glibc 2.39: fdlibm's e_acosf.c, a float
// rational approximation; no FMA). The vertex pseudo-normals are sums of angle * normal, so the last bit of the angle
// reaches TriangleData and the .bin bytes. tests/cpp/acosf_libm_main.cpp compares it with the host's acosf on EVERY float
// of [-1, 1] (2 130 706 434 values, 0 differences in this image); tests/test_capi_host.py runs that check.
__host__ __device__ inline float acosfLibm(float x) {
    const float one = 1.0f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f,
                pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
                pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f,
                qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
    uint32_t ux;
    memcpy(&ux, &x, 4);
    const int32_t hx = int32_t(ux), ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) return hx > 0 ? 0.0f : pi + 2.0f * pio2_lo;
    if (ix > 0x3f800000) return (x - x) / (x - x);
    if (ix < 0x3f000000) {   // |x| < 0.5
        if (ix <= 0x32800000) return pio2_hi + pio2_lo;
        const float z = x * x;
        const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const float r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    if (hx < 0) {            // x < -0.5
        const float z = (one + x) * 0.5f;
        const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const float s = sqrtf(z);
        const float r = p / q;
        const float w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    }
    const float z = (one - x) * 0.5f;   // x > 0.5
    const float s = sqrtf(z);
    uint32_t us;
    memcpy(&us, &s, 4);
    us &= 0xfffff000u;
    float df;
    memcpy(&df, &us, 4);
    const float c = (z - df * df) / (s + df);
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const float r = p / q;
    const float w = r * s + c;
    return 2.0f * (df + w);
}

}  // namespace sdfb200
