// sdflib::TriangleUtils::TriangleData — the 37-float record of include/SdfLib/utils/TriangleUtils.h:20-72, same
// field order as the reference serialises it (:53) — and calculateMeshTriangleData (src/utils/TriangleUtils.cpp:7-428),
// forwarded to the library's multi-threaded host implementation. The point-triangle distance functions run on the GPU
// (bulk entry: sdfb200_point_triangle in sdfb200.h).
#ifndef SDFB200_SDFLIB_TRIANGLE_UTILS_H
#define SDFB200_SDFLIB_TRIANGLE_UTILS_H

#include <array>
#include <stdexcept>
#include <string>
#include <vector>
#include <glm/glm.hpp>

#include "../../sdfb200.h"
#include "Mesh.h"

namespace sdflib
{
namespace TriangleUtils
{
    struct TriangleData
    {
        glm::vec3 origin;
        glm::mat3 transform;
        glm::vec2 b;
        glm::vec2 c;
        float v2;
        glm::vec2 v3;
        std::array<glm::vec3, 3> edgesNormal;
        std::array<glm::vec3, 3> verticesNormal;

        glm::vec3 getTriangleNormal() const { return glm::vec3(transform[0][2], transform[1][2], transform[2][2]); }
    };
    static_assert(sizeof(TriangleData) == 37 * sizeof(float), "TriangleData must be 37 packed floats");

    // local frames, edge pseudo-normals, angle-weighted vertex pseudo-normals, non-manifold vertex merging; bit-identical
    // to the reference's output (tests/test_capi_host.py)
    inline std::vector<TriangleData> calculateMeshTriangleData(const Mesh& mesh)
    {
        std::vector<TriangleData> out(mesh.getIndices().size() / 3);
        if (sdfb200_triangle_data(reinterpret_cast<const float*>(mesh.getVertices().data()), uint32_t(mesh.getVertices().size()),
                                  mesh.getIndices().data(), uint32_t(mesh.getIndices().size()), reinterpret_cast<float*>(out.data())) != SDFB200_OK)
            throw std::runtime_error(std::string("sdfb200: ") + sdfb200_last_error());
        return out;
    }
}
}

#endif
