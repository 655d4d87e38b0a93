// acosfLibm (sdflib_b200/csrc/tri_data_build.cuh) against the host libm's acosf on EVERY float of [-1, 1]:
// the device computes the corner angles of TriangleData with it, the reference with std::acos(float).
//   acosf_libm_main  ->  "ok <values>" or the first differences
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#define __host__
#define __device__
#define __forceinline__ inline
#include "tri_data_build.cuh"

int main() {
    long bad = 0, n = 0;
#pragma omp parallel for reduction(+ : bad, n) schedule(static)
    for (long u = 0; u <= 0x3f800000L; u++)
        for (int sgn = 0; sgn < 2; sgn++) {
            const uint32_t b = uint32_t(u) | (uint32_t(sgn) << 31);
            float x;
            std::memcpy(&x, &b, 4);
            volatile float vx = x;
            const float a = std::acos(float(vx)), m = sdfb200::acosfLibm(x);
            uint32_t ba, bm;
            std::memcpy(&ba, &a, 4);
            std::memcpy(&bm, &m, 4);
            n++;
            if (ba != bm && bad++ < 5) std::printf("x=%a libm=%a acosfLibm=%a\n", x, a, m);
        }
    if (bad) { std::printf("%ld of %ld values differ\n", bad, n); return 1; }
    std::printf("ok %ld\n", n);
    return 0;
}
