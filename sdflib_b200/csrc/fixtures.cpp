// Synthetic mesh fixture of the benchmark configs: the subdivided icosahedron of the reference's
// PrimitivesFactory::getIsosphere (src/utils/PrimitivesFactory.cpp:19-104). Same vertex and triangle
// ORDER as the reference (the centre triangle replaces its parent in place, the three corner
// triangles are appended; edge mid-points are created on first use and pushed onto the unit sphere),
// so config 1 ("icosphere 320-tri" = 2 subdivisions) is reproduced verbatim. Host-side utility, not
// on the hot path.
#include <cmath>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "sdf_internal.h"

using namespace sdfb200;

extern "C" int sdfb200_make_isosphere(uint32_t subdivisions, float* outVertices, uint32_t* outIndices,
                                      uint32_t* numVertices, uint32_t* numIndices) {
    if (!numVertices || !numIndices || subdivisions > 12) { setLastError("bad isosphere arguments"); return SDFB200_ERR_INVALID; }
    uint64_t nTri = 20, nVert = 12;
    for (uint32_t s = 0; s < subdivisions; s++) { nVert += nTri * 3 / 2; nTri *= 4; }
    *numVertices = uint32_t(nVert);
    *numIndices = uint32_t(nTri * 3);
    if (!outVertices || !outIndices) return SDFB200_OK;   // size query

    const float X = 0.525731112119133606f, Z = 0.850650808352039932f;
    const f3 base[12] = {{-X, 0, Z}, {X, 0, Z}, {-X, 0, -Z}, {X, 0, -Z}, {0, Z, X}, {0, Z, -X},
                         {0, -Z, X}, {0, -Z, -X}, {Z, X, 0}, {-Z, X, 0}, {Z, -X, 0}, {-Z, -X, 0}};
    static const uint32_t faces[60] = {0, 4, 1, 0, 9, 4, 9, 5, 4, 4, 5, 8, 4, 8, 1, 8, 10, 1, 8, 3, 10, 5, 3, 8, 5, 2, 3,
                                       2, 7, 3, 7, 10, 3, 7, 6, 10, 7, 11, 6, 11, 0, 6, 0, 1, 6, 6, 1, 10, 9, 0, 11,
                                       9, 11, 2, 9, 2, 5, 7, 2, 11};
    std::vector<f3> v(base, base + 12);
    std::vector<uint32_t> f(faces, faces + 60);
    v.reserve(nVert);
    f.reserve(nTri * 3);
    for (uint32_t s = 0; s < subdivisions; s++) {
        std::unordered_map<uint64_t, uint32_t> mids;
        mids.reserve(f.size());
        auto mid = [&](uint32_t a, uint32_t b) -> uint32_t {
            const uint64_t key = (uint64_t(a < b ? a : b) << 32) | (a < b ? b : a);
            auto it = mids.find(key);
            if (it != mids.end()) return it->second;   // each edge is met exactly twice on a closed mesh
            v.push_back(normalize3(0.5f * (v[a] + v[b])));
            const uint32_t id = uint32_t(v.size() - 1);
            mids.emplace(key, id);
            return id;
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
    std::memcpy(outVertices, v.data(), v.size() * sizeof(f3));
    std::memcpy(outIndices, f.data(), f.size() * sizeof(uint32_t));
    return SDFB200_OK;
}
