// sdflib::BoundingBox / sdflib::Mesh — drop-in mirror of include/SdfLib/utils/Mesh.h:16-106 of the reference for the
// B200-native library. Header-only; needs <glm/glm.hpp> exactly like the reference's public headers do.
// File loading through assimp (Mesh(std::string), Mesh.h:76-79) is out of scope: construct from arrays.
#ifndef SDFB200_SDFLIB_MESH_H
#define SDFB200_SDFLIB_MESH_H

#include <cmath>
#include <cstdint>
#include <string>
#include <vector>
#include <glm/glm.hpp>

namespace sdflib
{
struct BoundingBox
{
    BoundingBox() : min(INFINITY), max(-INFINITY) {}
    BoundingBox(glm::vec3 min, glm::vec3 max) : min(min), max(max) {}
    glm::vec3 min;
    glm::vec3 max;

    glm::vec3 getSize() const { return max - min; }
    glm::vec3 getCenter() const { return min + 0.5f * getSize(); }
    void addMargin(float margin) { min -= glm::vec3(margin); max += glm::vec3(margin); }

    // Mesh.h:42-46
    float getDistance(glm::vec3 point) const
    {
        glm::vec3 q = glm::abs(point - getCenter()) - 0.5f * getSize();
        return glm::length(glm::max(q, glm::vec3(0.0f))) + glm::min(glm::max(q.x, glm::max(q.y, q.z)), 0.0f);
    }
};

class Mesh
{
public:
    Mesh() {}
    // src/utils/Mesh.cpp:34-42: the arrays are copied
    Mesh(glm::vec3* vertices, uint32_t numVertices, uint32_t* indices, uint32_t numIndices)
        : mVertices(vertices, vertices + numVertices), mIndices(indices, indices + numIndices)
    {
        computeBoundingBox();
    }

    std::vector<glm::vec3>& getVertices() { return mVertices; }
    const std::vector<glm::vec3>& getVertices() const { return mVertices; }
    std::vector<uint32_t>& getIndices() { return mIndices; }
    const std::vector<uint32_t>& getIndices() const { return mIndices; }
    const BoundingBox& getBoundingBox() const { return mBBox; }

    void computeBoundingBox()   // src/utils/Mesh.cpp:66-76
    {
        glm::vec3 min(INFINITY), max(-INFINITY);
        for (const glm::vec3& v : mVertices)
        {
            min.x = glm::min(min.x, v.x); max.x = glm::max(max.x, v.x);
            min.y = glm::min(min.y, v.y); max.y = glm::max(max.y, v.y);
            min.z = glm::min(min.z, v.z); max.z = glm::max(max.z, v.z);
        }
        mBBox = BoundingBox(min, max);
    }

    void applyTransform(glm::mat4 trans)   // src/utils/Mesh.cpp:78-87
    {
        for (glm::vec3& v : mVertices) v = glm::vec3(trans * glm::vec4(v, 1.0f));
        computeBoundingBox();
    }

private:
    std::vector<glm::vec3> mVertices;
    std::vector<uint32_t> mIndices;
    BoundingBox mBBox;
};
}

#endif
