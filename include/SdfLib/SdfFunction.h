// sdflib::SdfFunction — drop-in mirror of include/SdfLib/SdfFunction.h:12-58 on top of the C-ABI (include/sdfb200.h).
//
// Same virtuals, same saveToFile / loadFromFile (same .bin bytes). Additive: getDistances(), the bulk entry that is
// the actual hot path (one kernel launch for the whole array; host or device pointers). The scalar virtuals forward
// to a 1-element launch — correct, but a per-point round trip to the GPU; port loops over points to getDistances().
// Errors: like the reference, file errors return false / nullptr; construction errors (which the reference answers
// with undefined behaviour) throw std::runtime_error carrying sdfb200_last_error().
#ifndef SDFB200_SDFLIB_SDF_FUNCTION_H
#define SDFB200_SDFLIB_SDF_FUNCTION_H

#include <cstddef>
#include <memory>
#include <stdexcept>
#include <string>
#include <glm/glm.hpp>

#include "../sdfb200.h"
#include "utils/Mesh.h"

namespace sdflib
{
class SdfFunction
{
public:
    enum SdfFormat { GRID, OCTREE, EXACT_OCTREE, NONE };

    virtual ~SdfFunction() { if (mHandle) sdfb200_free(mHandle); }
    SdfFunction(const SdfFunction&) = delete;
    SdfFunction& operator=(const SdfFunction&) = delete;

    virtual float getDistance(glm::vec3 sample) const
    {
        float d = 0.0f;
        check(sdfb200_query(mHandle, &sample.x, 1, &d, nullptr, 0, nullptr));
        return d;
    }
    virtual float getDistance(glm::vec3 sample, glm::vec3& outGradient) const
    {
        float d = 0.0f;
        glm::vec3 g(0.0f);
        check(sdfb200_query(mHandle, &sample.x, 1, &d, &g.x, 0, nullptr));
        outGradient = g;
        return d;
    }
    // Bulk getDistance (hot path 2): n points in, n distances (and n gradients if outGradients != nullptr) out.
    // devicePointers: the three arrays live on the structure's GPU and the call only enqueues the kernel on `stream`.
    void getDistances(const glm::vec3* samples, size_t n, float* outDistances, glm::vec3* outGradients = nullptr,
                      bool devicePointers = false, void* cudaStream = nullptr, bool referenceOperationOrder = false) const
    {
        const int flags = (devicePointers ? SDFB200_QUERY_DEVICE_POINTERS : 0) | (referenceOperationOrder ? SDFB200_QUERY_EXACT_ORDER : 0);
        check(sdfb200_query(mHandle, reinterpret_cast<const float*>(samples), n, outDistances, reinterpret_cast<float*>(outGradients), flags, cudaStream));
    }

    virtual BoundingBox getSampleArea() const
    {
        const sdfb200_info i = info();
        return BoundingBox(glm::vec3(i.box_min[0], i.box_min[1], i.box_min[2]), glm::vec3(i.box_max[0], i.box_max[1], i.box_max[2]));
    }
    virtual SdfFormat getFormat() const { return SdfFormat::NONE; }

    bool saveToFile(const std::string& outputPath) { return sdfb200_save(mHandle, outputPath.c_str()) == SDFB200_OK; }
    static std::unique_ptr<SdfFunction> loadFromFile(const std::string& inputPath);   // defined in ExactOctreeSdf.h, after both subclasses

    sdfb200_sdf* handle() const { return mHandle; }

protected:
    SdfFunction() {}
    explicit SdfFunction(sdfb200_sdf* h) : mHandle(h) {}
    sdfb200_sdf* mHandle = nullptr;

    static void check(int code)
    {
        if (code != SDFB200_OK) throw std::runtime_error(std::string("sdfb200: ") + sdfb200_last_error());
    }
    sdfb200_info info() const
    {
        sdfb200_info i;
        check(sdfb200_get_info(mHandle, &i));
        return i;
    }
};
}

#include "OctreeSdf.h"   // brings both subclasses and the definition of loadFromFile (README.md:53-72 includes only this header)
#endif
