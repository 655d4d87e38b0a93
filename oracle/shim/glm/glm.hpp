// TEST INFRASTRUCTURE — not product code.
//
// Minimal stand-in for g-truc/glm 0.9.8 (pinned by the reference at
// libs/CMakeLists.txt:5-18, commit 89e52e3; NOT vendored under /root/reference and
// not available offline). It exists only so that the UNMODIFIED reference sources can be
// compiled into oracle/_ref (see oracle/Makefile). Every function restates glm 0.9.8's
// published scalar formula and operation order:
//   dot(a,b)      = a.x*b.x + a.y*b.y + a.z*b.z   (products first, left-to-right adds)
//   normalize(v)  = v * (1 / sqrt(dot(v,v)))
//   length(v)     = sqrt(dot(v,v))
//   cross(x,y)    = (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
//   min(a,b)      = (b < a) ? b : a ;  max(a,b) = (a < b) ? b : a ;  clamp = min(max(x,lo),hi)
//   sign(x)       = (0 < x) - (x < 0) ;  abs(x) = x >= 0 ? x : -x ;  fract(x) = x - floor(x)
//   inverse(mat3) = cofactors * (1/det), det expanded along the first row
//   mat3 * vec3   = column-major: m[0][r]*v.x + m[1][r]*v.y + m[2][r]*v.z
// Parity note: glm itself is absent, so bit-level claims that depend on glm's exact
// operation order are "parity unpinned" (DESIGN.md §oracle).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <cassert>
#include <cstring>
#include <type_traits>
#include <limits>

namespace glm {

template <typename T> struct tvec2;
template <typename T> struct tvec3;
template <typename T> struct tvec4;

template <typename T> struct tvec2 {
    T x, y;
    tvec2() : x(0), y(0) {}
    tvec2(T a, T b) : x(a), y(b) {}
    explicit tvec2(T s) : x(s), y(s) {}
    template <typename U> tvec2(const tvec3<U>& v);
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};

template <typename T> struct tvec3 {
    T x, y, z;
    tvec3() : x(0), y(0), z(0) {}
    template <typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    tvec3(S s) : x(T(s)), y(T(s)), z(T(s)) {}
    template <typename A, typename B, typename C> tvec3(A a, B b, C c) : x(T(a)), y(T(b)), z(T(c)) {}
    template <typename U> tvec3(const tvec3<U>& v) : x(T(v.x)), y(T(v.y)), z(T(v.z)) {}
    template <typename U> tvec3(const tvec4<U>& v);
    template <typename U, typename S> tvec3(const tvec2<U>& v, S s) : x(T(v.x)), y(T(v.y)), z(T(s)) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    tvec3& operator-=(const tvec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    tvec3& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
};

template <typename T> struct tvec4 {
    T x, y, z, w;
    tvec4() : x(0), y(0), z(0), w(0) {}
    tvec4(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
    template <typename S> tvec4(const tvec3<T>& v, S s) : x(v.x), y(v.y), z(v.z), w(T(s)) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};

template <typename T> template <typename U> tvec2<T>::tvec2(const tvec3<U>& v) : x(T(v.x)), y(T(v.y)) {}
template <typename T> template <typename U> tvec3<T>::tvec3(const tvec4<U>& v) : x(T(v.x)), y(T(v.y)), z(T(v.z)) {}

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec3<int> ivec3;
typedef tvec3<unsigned> uvec3;
typedef tvec3<double> dvec3;

#define SHIM_V3_BINOP(op)                                                                          \
    template <typename T> inline tvec3<T> operator op(const tvec3<T>& a, const tvec3<T>& b) {     \
        return tvec3<T>(a.x op b.x, a.y op b.y, a.z op b.z); }                                     \
    template <typename T> inline tvec3<T> operator op(const tvec3<T>& a, T s) {                   \
        return tvec3<T>(a.x op s, a.y op s, a.z op s); }                                           \
    template <typename T> inline tvec3<T> operator op(T s, const tvec3<T>& a) {                   \
        return tvec3<T>(s op a.x, s op a.y, s op a.z); }
SHIM_V3_BINOP(+)
SHIM_V3_BINOP(-)
SHIM_V3_BINOP(*)
SHIM_V3_BINOP(/)
#undef SHIM_V3_BINOP

inline ivec3 operator-(const ivec3& a, int s) { return ivec3(a.x - s, a.y - s, a.z - s); }
template <typename T> inline tvec3<T> operator-(const tvec3<T>& a) { return tvec3<T>(-a.x, -a.y, -a.z); }
template <typename T> inline bool operator==(const tvec3<T>& a, const tvec3<T>& b) {
    return a.x == b.x && a.y == b.y && a.z == b.z; }
template <typename T> inline bool operator!=(const tvec3<T>& a, const tvec3<T>& b) { return !(a == b); }
template <typename T> inline tvec2<T> operator*(const tvec2<T>& a, T s) { return tvec2<T>(a.x * s, a.y * s); }

using std::sqrt;
using std::pow;
using std::acos;
using std::tan;
using std::log2;
using std::ceil;
using std::floor;
using std::isnan;
using std::round;  // glm 0.9.8 func_common.inl: `using ::std::round` when the STL is C++11

template <typename T> inline T abs(T a) { return a >= T(0) ? a : -a; }
template <typename T> inline T min(T a, T b) { return (b < a) ? b : a; }
template <typename T> inline T max(T a, T b) { return (a < b) ? b : a; }
template <typename T> inline T sign(T x) { return T((T(0) < x) - (x < T(0))); }
template <typename T> inline T clamp(T x, T lo, T hi) { return min(max(x, lo), hi); }
inline float fract(float x) { return x - std::floor(x); }
inline double radians(double d) { return d * 0.01745329251994329576923690768489; }

template <typename T> inline tvec3<T> abs(const tvec3<T>& v) { return tvec3<T>(abs(v.x), abs(v.y), abs(v.z)); }
template <typename T> inline tvec3<T> sign(const tvec3<T>& v) { return tvec3<T>(sign(v.x), sign(v.y), sign(v.z)); }
inline vec3 floor(const vec3& v) { return vec3(std::floor(v.x), std::floor(v.y), std::floor(v.z)); }
inline vec3 ceil(const vec3& v) { return vec3(std::ceil(v.x), std::ceil(v.y), std::ceil(v.z)); }
inline vec3 fract(const vec3& v) { return vec3(fract(v.x), fract(v.y), fract(v.z)); }
inline vec3 round(const vec3& v) { return vec3(std::round(v.x), std::round(v.y), std::round(v.z)); }
template <typename T> inline tvec3<T> max(const tvec3<T>& a, const tvec3<T>& b) {
    return tvec3<T>(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
template <typename T> inline tvec3<T> min(const tvec3<T>& a, const tvec3<T>& b) {
    return tvec3<T>(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }

template <typename T> inline T dot(const tvec3<T>& a, const tvec3<T>& b) {
    T tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z;
    return tx + ty + tz;
}
template <typename T> inline T dot(const tvec2<T>& a, const tvec2<T>& b) {
    T tx = a.x * b.x, ty = a.y * b.y;
    return tx + ty;
}
template <typename T> inline tvec3<T> cross(const tvec3<T>& x, const tvec3<T>& y) {
    return tvec3<T>(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
template <typename T> inline T length(const tvec3<T>& v) { return std::sqrt(dot(v, v)); }
template <typename T> inline tvec3<T> normalize(const tvec3<T>& v) { return v * (T(1) / std::sqrt(dot(v, v))); }
template <typename T> inline tvec2<T> normalize(const tvec2<T>& v) { return v * (T(1) / std::sqrt(dot(v, v))); }

struct bvec3 { bool x, y, z; };
template <typename T> inline bvec3 greaterThan(const tvec3<T>& a, const tvec3<T>& b) {
    return bvec3{a.x > b.x, a.y > b.y, a.z > b.z}; }
inline bool any(const bvec3& b) { return b.x || b.y || b.z; }

struct mat3 {
    vec3 c[3];
    mat3() { c[0] = vec3(1, 0, 0); c[1] = vec3(0, 1, 0); c[2] = vec3(0, 0, 1); }
    mat3(const vec3& a, const vec3& b, const vec3& d) { c[0] = a; c[1] = b; c[2] = d; }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
typedef mat3 mat3x3;

inline vec3 operator*(const mat3& m, const vec3& v) {
    return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
                m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
                m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}
inline mat3 transpose(const mat3& m) {
    mat3 r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[i][j] = m[j][i];
    return r;
}
inline mat3 inverse(const mat3& m) {
    float inv = 1.0f / (+m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2])
                        - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2])
                        + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
    mat3 I;
    I[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * inv;
    I[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * inv;
    I[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * inv;
    I[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * inv;
    I[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * inv;
    I[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * inv;
    I[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * inv;
    I[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * inv;
    I[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * inv;
    return I;
}

struct mat4 {
    vec4 c[4];
    mat4() {}
    explicit mat4(float d) { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) c[i][j] = (i == j) ? d : 0.0f; }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline vec4 operator*(const mat4& m, const vec4& v) {
    vec4 r;
    for (int i = 0; i < 4; i++) r[i] = m[0][i] * v.x + m[1][i] * v.y + m[2][i] * v.z + m[3][i] * v.w;
    return r;
}

}  // namespace glm
