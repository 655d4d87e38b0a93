// sdflib::ExactOctreeSdf — drop-in mirror of include/SdfLib/ExactOctreeSdf.h:17-214 on top of the C-ABI.
// Same constructor (src/sdf/ExactOctreeSdf.cpp:7-31) and getters; construction and queries run on the GPU.
// Unlike the reference's (mutable scratch, ExactOctreeSdf.h:204-205) the query methods are re-entrant.
#ifndef SDFB200_SDFLIB_EXACT_OCTREE_SDF_H
#define SDFB200_SDFLIB_EXACT_OCTREE_SDF_H

#include <vector>

#include "SdfFunction.h"
#include "utils/TriangleUtils.h"

namespace sdflib
{
class ExactOctreeSdf : public SdfFunction
{
public:
    // ExactOctreeSdf::OctreeNode (ExactOctreeSdf.h:35-77)
    struct OctreeNode
    {
        static constexpr uint32_t IS_LEAF_MASK = 1u << 31;
        static constexpr uint32_t CHILDREN_INDEX_MASK = ~IS_LEAF_MASK;
        uint32_t childrenIndex;
        uint32_t trianglesArrayIndex;
        bool isLeaf() const { return childrenIndex & IS_LEAF_MASK; }
        uint32_t getChildrenIndex() const { return childrenIndex & CHILDREN_INDEX_MASK; }
    };

    ExactOctreeSdf(const Mesh& mesh, BoundingBox box, uint32_t maxDepth, uint32_t startDepth = 1,
                   uint32_t minTrianglesPerNode = 128, uint32_t numThreads = 1)
    {
        const float b[6] = {box.min.x, box.min.y, box.min.z, box.max.x, box.max.y, box.max.z};
        const std::vector<int>& devices = getDevices();
        if (devices.size() > 1)
        {
            std::vector<sdfb200_sdf*> handles(devices.size(), nullptr);
            check(sdfb200_build_exact_multi(reinterpret_cast<const float*>(mesh.getVertices().data()), uint32_t(mesh.getVertices().size()),
                                            mesh.getIndices().data(), uint32_t(mesh.getIndices().size()), b, maxDepth, startDepth,
                                            minTrianglesPerNode, numThreads, devices.data(), uint32_t(devices.size()), handles.data()));
            adopt(handles);
        }
        else
        {
            if (devices.size() == 1) check(sdfb200_set_device(devices[0]));
            check(sdfb200_build_exact(reinterpret_cast<const float*>(mesh.getVertices().data()), uint32_t(mesh.getVertices().size()),
                                      mesh.getIndices().data(), uint32_t(mesh.getIndices().size()), b, maxDepth, startDepth,
                                      minTrianglesPerNode, numThreads, &mHandle));
        }
        fetch();
    }

    glm::ivec3 getStartGridSize() const { return glm::ivec3(mInfo.start_grid_size); }
    const BoundingBox& getGridBoundingBox() const { return mBox; }
    BoundingBox getSampleArea() const override { return mBox; }
    uint32_t getMaxTrianglesInLeafs() const { return mInfo.max_triangles_in_leafs; }
    uint32_t getMinTrianglesInLeafs() const { return mInfo.min_triangles_in_leafs; }
    uint32_t getOctreeMaxDepth() const { return mInfo.max_depth; }
    const std::vector<OctreeNode>& getOctreeData() const { return mOctreeData; }
    const std::vector<uint32_t>& getTrianglesSets() const { return mTrianglesSets; }
    const std::vector<uint8_t>& getTrianglesMasks() const { return mTrianglesMasks; }
    const std::vector<TriangleUtils::TriangleData>& getTrianglesData() { return mTrianglesData; }
    SdfFormat getFormat() const override { return SdfFormat::EXACT_OCTREE; }

private:
    friend class SdfFunction;
    explicit ExactOctreeSdf(sdfb200_sdf* h) : SdfFunction(h) { fetch(); }
    void fetch()
    {
        mInfo = info();
        mBox = SdfFunction::getSampleArea();
        mOctreeData.resize(size_t(mInfo.octree_words));
        mTrianglesSets.resize(size_t(mInfo.triangle_sets_words));
        mTrianglesMasks.resize(size_t(mInfo.triangle_masks_bytes));
        mTrianglesData.resize(size_t(mInfo.num_triangles));
        check(sdfb200_get_octree_data(mHandle, reinterpret_cast<uint32_t*>(mOctreeData.data()), 2 * mInfo.octree_words));
        check(sdfb200_get_exact_arrays(mHandle, mTrianglesSets.data(), mTrianglesMasks.data(), reinterpret_cast<float*>(mTrianglesData.data())));
    }
    sdfb200_info mInfo;
    BoundingBox mBox;
    std::vector<OctreeNode> mOctreeData;
    std::vector<uint32_t> mTrianglesSets;
    std::vector<uint8_t> mTrianglesMasks;
    std::vector<TriangleUtils::TriangleData> mTrianglesData;
};
}

#include "OctreeSdf.h"

namespace sdflib
{
// src/sdf/SdfFunction.cpp:45-79: nullptr when the file cannot be loaded
inline std::unique_ptr<SdfFunction> SdfFunction::loadFromFile(const std::string& inputPath)
{
    sdfb200_sdf* h = nullptr;
    if (sdfb200_load(inputPath.c_str(), &h) != SDFB200_OK || !h) return nullptr;
    sdfb200_info i;
    if (sdfb200_get_info(h, &i) != SDFB200_OK) { sdfb200_free(h); return nullptr; }
    if (i.format == SDFB200_FORMAT_OCTREE) return std::unique_ptr<SdfFunction>(new OctreeSdf(h));
    return std::unique_ptr<SdfFunction>(new ExactOctreeSdf(h));
}
}

#endif
