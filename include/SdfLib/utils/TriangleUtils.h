// sdflib::TriangleUtils::TriangleData — the 37-float record of include/SdfLib/utils/TriangleUtils.h:20-72, same
// field order as the reference serialises it (:53). Data only: the distance functions run on the GPU.
#ifndef SDFB200_SDFLIB_TRIANGLE_UTILS_H
#define SDFB200_SDFLIB_TRIANGLE_UTILS_H

#include <array>
#include <glm/glm.hpp>

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
}
}

#endif
