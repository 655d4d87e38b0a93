// sdflib::OctreeSdf — drop-in mirror of include/SdfLib/OctreeSdf.h:20-292 on top of the C-ABI.
// Same constructors (src/sdf/OctreeSdf.cpp:18-86) and getters; construction and queries run on the GPU.
// numThreads keeps its layout-selecting meaning (< 2: single depth-first layout, >= 2: per-start-voxel layout).
#ifndef SDFB200_SDFLIB_OCTREE_SDF_H
#define SDFB200_SDFLIB_OCTREE_SDF_H

#include <array>
#include <optional>
#include <utility>
#include <vector>

#include "SdfFunction.h"

namespace sdflib
{
class OctreeSdf : public SdfFunction
{
public:
    enum InitAlgorithm { UNIFORM, NO_CONTINUITY, CONTINUITY };
    enum TerminationRule { NONE, TRAPEZOIDAL_RULE, SIMPSONS_RULE, BY_DISTANCE_RULE };

    static std::optional<InitAlgorithm> stringToInitAlgorithm(std::string text)   // OctreeSdf.h:30-37
    {
        if (text == "uniform" || text == "UNIFORM") return InitAlgorithm::UNIFORM;
        if (text == "no_continuity" || text == "NO_CONTINUITY") return InitAlgorithm::NO_CONTINUITY;
        if (text == "continuity" || text == "CONTINUITY") return InitAlgorithm::CONTINUITY;
        return std::optional<InitAlgorithm>();
    }
    static std::optional<TerminationRule> stringToTerminationRule(std::string text)   // OctreeSdf.h:128-148
    {
        if (text == "none" || text == "NONE") return TerminationRule::NONE;
        if (text == "trapezoidal_rule" || text == "TRAPEZOIDAL_RULE") return TerminationRule::TRAPEZOIDAL_RULE;
        if (text == "simpsons_rule" || text == "SIMPSONS_RULE") return TerminationRule::SIMPSONS_RULE;
        if (text == "by_distance_rule" || text == "BY_DISTANCE_RULE") return TerminationRule::BY_DISTANCE_RULE;
        return std::optional<TerminationRule>();
    }

    class TerminationRuleParams
    {
    public:
        std::array<float, 2> params;
        static TerminationRuleParams setNoneRuleParams() { return TerminationRuleParams(); }
        static TerminationRuleParams setTrapezoidalRuleParams(float expectedError) { return TerminationRuleParams{{expectedError, 0.0f}}; }
        static TerminationRuleParams setSimpsonRuleParams(float expectedError) { return TerminationRuleParams{{expectedError, 0.0f}}; }
        static TerminationRuleParams setByDistanceRuleParams(float baseError, float errorDecayByDistance) { return TerminationRuleParams{{baseError, errorDecayByDistance}}; }
        float& operator[](int p) { return params[p]; }
    };

    // OctreeSdf::OctreeNode (OctreeSdf.h:39-98): bit 31 leaf, bit 30 mark, low 30 bits children / coefficient index
    struct OctreeNode
    {
        static constexpr uint32_t IS_LEAF_MASK = 1u << 31;
        static constexpr uint32_t MARK_MASK = 1u << 30;
        static constexpr uint32_t CHILDREN_INDEX_MASK = ~(IS_LEAF_MASK | MARK_MASK);
        union { uint32_t childrenIndex; float value; };
        bool isLeaf() const { return childrenIndex & IS_LEAF_MASK; }
        uint32_t getChildrenIndex() const { return childrenIndex & CHILDREN_INDEX_MASK; }
    };

    OctreeSdf(const Mesh& mesh, BoundingBox box, uint32_t depth, uint32_t startDepth, float maxError = 1e-3,
              InitAlgorithm initAlgorithm = InitAlgorithm::NO_CONTINUITY, uint32_t numThreads = 1)
        : OctreeSdf(mesh, box, depth, startDepth, TerminationRule::TRAPEZOIDAL_RULE,
                    TerminationRuleParams::setTrapezoidalRuleParams(maxError), initAlgorithm, numThreads) {}

    OctreeSdf(const Mesh& mesh, BoundingBox box, uint32_t depth, uint32_t startDepth, TerminationRule terminationRule,
              TerminationRuleParams params, InitAlgorithm initAlgorithm, uint32_t numThreads = 1)
    {
        const float b[6] = {box.min.x, box.min.y, box.min.z, box.max.x, box.max.y, box.max.z};
        const std::vector<int>& devices = getDevices();
        if (devices.size() > 1)
        {
            std::vector<sdfb200_sdf*> handles(devices.size(), nullptr);
            check(sdfb200_build_octree_multi(reinterpret_cast<const float*>(mesh.getVertices().data()), uint32_t(mesh.getVertices().size()),
                                             mesh.getIndices().data(), uint32_t(mesh.getIndices().size()), b, depth, startDepth,
                                             int(terminationRule), params.params[0], params.params[1], int(initAlgorithm), numThreads,
                                             devices.data(), uint32_t(devices.size()), handles.data()));
            adopt(handles);
        }
        else
        {
            if (devices.size() == 1) check(sdfb200_set_device(devices[0]));
            check(sdfb200_build_octree(reinterpret_cast<const float*>(mesh.getVertices().data()), uint32_t(mesh.getVertices().size()),
                                       mesh.getIndices().data(), uint32_t(mesh.getIndices().size()), b, depth, startDepth,
                                       int(terminationRule), params.params[0], params.params[1], int(initAlgorithm), numThreads, &mHandle));
        }
        fetch();
    }

    // Sphere tracing of n rays (the raycast loop of the reference's viewer, src/render_engine/shaders/sdfOctreeRender.comp:392-410,
    // one ray per GPU thread): outHit = last evaluated position, outTravelled = marched distance or -1 when no surface was reached.
    void sphereTrace(const glm::vec3* origins, const glm::vec3* directions, size_t n, float farDistance, glm::vec3* outHit, float* outTravelled,
                     uint32_t* outIterations = nullptr, float epsilon = 1e-5f, uint32_t maxIterations = 1024, bool devicePointers = false,
                     void* cudaStream = nullptr, bool referenceOperationOrder = false) const
    {
        const int flags = (devicePointers ? SDFB200_QUERY_DEVICE_POINTERS : 0) | (referenceOperationOrder ? SDFB200_QUERY_EXACT_ORDER : 0);
        check(sdfb200_sphere_trace(mHandle, reinterpret_cast<const float*>(origins), reinterpret_cast<const float*>(directions), n, epsilon,
                                   farDistance, maxIterations, reinterpret_cast<float*>(outHit), outTravelled, outIterations, flags, cudaStream));
    }

    float getOctreeValueRange() const { return mInfo.value_range; }
    float getOctreeMinBorderValue() const { return mInfo.min_border_value; }
    glm::ivec3 getStartGridSize() const { return glm::ivec3(mInfo.start_grid_size); }
    const BoundingBox& getGridBoundingBox() const { return mBox; }
    BoundingBox getSampleArea() const override { return mBox; }
    uint32_t getOctreeMaxDepth() const { return mInfo.max_depth; }
    const std::vector<OctreeNode>& getOctreeData() const { return mOctreeData; }
    std::vector<OctreeNode>& getOctreeData() { return mOctreeData; }   // host mirror; the device copy is what queries read
    SdfFunction::SdfFormat getFormat() const override { return SdfFunction::SdfFormat::OCTREE; }

    // Volume fraction of the (unit) octree covered by the leaves of every depth (src/sdf/OctreeSdf.cpp:232-277):
    // #leaves(depth) * 8^-depth. Walks the host mirror with an explicit stack.
    void getDepthDensity(std::vector<float>& depthsDensity)
    {
        depthsDensity.assign(size_t(mInfo.max_depth) + 1, 0.0f);
        std::vector<uint32_t> leavesPerDepth(depthsDensity.size(), 0u);
        uint32_t startDepth = 0;
        while ((1 << startDepth) < mInfo.start_grid_size) startDepth++;
        const size_t startSlots = size_t(mInfo.start_grid_size) * mInfo.start_grid_size * mInfo.start_grid_size;
        std::vector<std::pair<uint32_t, uint32_t>> open;   // (word, depth)
        for (size_t slot = 0; slot < startSlots; slot++) open.emplace_back(uint32_t(slot), startDepth);
        while (!open.empty())
        {
            const std::pair<uint32_t, uint32_t> at = open.back();
            open.pop_back();
            const OctreeNode& node = mOctreeData[at.first];
            if (node.isLeaf()) { if (at.second < leavesPerDepth.size()) leavesPerDepth[at.second]++; }
            else for (uint32_t c = 0; c < 8; c++) open.emplace_back(node.getChildrenIndex() + c, at.second + 1);
        }
        float cellVolume = 1.0f;
        for (size_t d = 0; d < depthsDensity.size(); d++)
        {
            depthsDensity[d] = cellVolume * static_cast<float>(leavesPerDepth[d]);
            cellVolume *= 0.125f;
        }
    }

private:
    friend class SdfFunction;
    explicit OctreeSdf(sdfb200_sdf* h) : SdfFunction(h) { fetch(); }
    void fetch()
    {
        mInfo = info();
        mBox = SdfFunction::getSampleArea();
        mOctreeData.resize(size_t(mInfo.octree_words));
        check(sdfb200_get_octree_data(mHandle, reinterpret_cast<uint32_t*>(mOctreeData.data()), mInfo.octree_words));
    }
    sdfb200_info mInfo;
    BoundingBox mBox;
    std::vector<OctreeNode> mOctreeData;
};
}

#include "ExactOctreeSdf.h"
#endif
