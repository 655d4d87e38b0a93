// sdflib::SdfFunction — drop-in mirror of include/SdfLib/SdfFunction.h:12-58 on top of the C-ABI (include/sdfb200.h).
//
// Same virtuals, same saveToFile / loadFromFile (same .bin bytes). Additive: getDistances(), the bulk entry that is
// the actual hot path (one kernel launch for the whole array; host or device pointers). The scalar virtuals forward
// to a 1-element launch — correct, but a per-point round trip to the GPU; port loops over points to getDistances().
// Errors: like the reference, file errors return false / nullptr; construction errors (which the reference answers
// with undefined behaviour) throw std::runtime_error carrying sdfb200_last_error().
// Several GPUs: SdfFunction::setDevices({0, 1, ...}) (or SDFB200_DEVICES=0,1,... in the environment) makes the
// constructors of OctreeSdf / ExactOctreeSdf build over those devices of the box (sdfb200_build_*_multi: start-depth voxels
// partitioned, one NCCL all-gather) and keep one replica of the structure per device; getDistances() on host arrays then
// spreads a large batch over the replicas. Nothing else changes for the caller.
#ifndef SDFB200_SDFLIB_SDF_FUNCTION_H
#define SDFB200_SDFLIB_SDF_FUNCTION_H

#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <glm/glm.hpp>

#include "../sdfb200.h"
#include "utils/Mesh.h"

namespace sdflib
{
class SdfFunction
{
public:
    enum SdfFormat { GRID, OCTREE, EXACT_OCTREE, NONE };

    virtual ~SdfFunction()
    {
        if (mHandle) sdfb200_free(mHandle);
        for (sdfb200_sdf* r : mReplicas) sdfb200_free(r);
    }
    // devices the constructors build on and bulk host-pointer queries are spread over (empty: the current device only)
    static void setDevices(const std::vector<int>& devices) { deviceList() = devices; }
    static const std::vector<int>& getDevices() { return deviceList(); }
    size_t numReplicas() const { return 1 + mReplicas.size(); }
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
        if (!devicePointers && !mReplicas.empty() && n >= (size_t(1) << 20))
        {
            // one contiguous slab per replica, one host thread per device
            std::vector<sdfb200_sdf*> all(1, mHandle);
            all.insert(all.end(), mReplicas.begin(), mReplicas.end());
            std::vector<int> codes(all.size(), SDFB200_OK);
            std::vector<std::string> messages(all.size());
            std::vector<std::thread> workers;
            const size_t per = (n + all.size() - 1) / all.size();
            for (size_t k = 0; k < all.size(); k++)
                workers.emplace_back([&, k] {
                    const size_t first = std::min(n, k * per), count = std::min(per, n - first);
                    if (!count) return;
                    codes[k] = sdfb200_query(all[k], reinterpret_cast<const float*>(samples + first), count, outDistances + first,
                                             outGradients ? reinterpret_cast<float*>(outGradients + first) : nullptr, flags, nullptr);
                    if (codes[k] != SDFB200_OK) messages[k] = sdfb200_last_error();   // thread-local message
                });
            for (std::thread& t : workers) t.join();
            for (size_t k = 0; k < all.size(); k++)
                if (codes[k] != SDFB200_OK) throw std::runtime_error("sdfb200: " + messages[k]);
            return;
        }
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
    std::vector<sdfb200_sdf*> mReplicas;   // the same structure on the other devices of setDevices()

    static std::vector<int>& deviceList()
    {
        static std::vector<int> devices = [] {
            std::vector<int> d;
            if (const char* e = std::getenv("SDFB200_DEVICES"))
                for (const char* p = e; *p;)
                {
                    char* end = nullptr;
                    const long v = std::strtol(p, &end, 10);
                    if (end == p) break;
                    d.push_back(int(v));
                    p = (*end == ',') ? end + 1 : end;
                }
            return d;
        }();
        return devices;
    }
    // adopts the handles of a multi-device build: the first is the primary, the others are replicas
    void adopt(const std::vector<sdfb200_sdf*>& handles)
    {
        mHandle = handles.at(0);
        mReplicas.assign(handles.begin() + 1, handles.end());
    }

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
