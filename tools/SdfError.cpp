// SdfError — same command line and report as the reference's error meter (src/tools/SdfError/main.cpp:11-95):
//   SdfError sdf_path exact_sdf_path [num_samples_in_millions]
// loads two .bin structures, draws the same rand()-driven samples inside the first one's sample area, queries both
// and prints microseconds per query, RMSE, MAE and the maximum error. The two per-point loops of the reference are one
// bulk getDistances() call each (the additive array entry of SdfFunction.h), i.e. one kernel launch per structure.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <vector>

#include "SdfLib/SdfFunction.h"

using namespace sdflib;

// The reference's sample stream: three rand() draws per point (x, y, z in that order), mapped into the sample area
// shrunk by 1e-5 (main.cpp:44-57).
static std::vector<glm::vec3> drawSamples(const BoundingBox& area, size_t count)
{
    const glm::vec3 middle = area.getCenter();
    const glm::vec3 extent = area.getSize() - glm::vec3(1e-5);
    std::vector<glm::vec3> points(count);
    for (glm::vec3& p : points)
    {
        const float u = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
        const float v = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
        const float w = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
        p = middle + (glm::vec3(u, v, w) - 0.5f) * extent;
    }
    return points;
}

// one bulk call instead of the reference's per-point loop; returns microseconds per query
static float timedQuery(const SdfFunction& f, const std::vector<glm::vec3>& points, std::vector<float>& out)
{
    const auto t0 = std::chrono::steady_clock::now();
    f.getDistances(points.data(), points.size(), out.data());
    const float seconds = std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count();
    return seconds * 1.0e6f / static_cast<float>(points.size());
}

int main(int argc, char** argv)
{
    if (argc < 3 || std::string(argv[1]) == "-h" || std::string(argv[1]) == "--help")
    {
        std::cerr << "  SdfError sdf_path exact_sdf_path [num_samples_in_millions]\n\n    Calculate the error of a sdf\n";
        return argc < 3 ? 1 : 0;
    }
    try
    {
        std::unique_ptr<SdfFunction> sdf = SdfFunction::loadFromFile(argv[1]);
        std::unique_ptr<SdfFunction> exactSdf = SdfFunction::loadFromFile(argv[2]);
        if (!sdf || !exactSdf) { std::cerr << "[error] Cannot load the models: " << sdfb200_last_error() << std::endl; return 1; }
        std::cout << "[info] Models Loaded" << std::endl;

        const uint32_t millions = argc > 3 ? uint32_t(std::strtoul(argv[3], nullptr, 10)) : 1u;
        const size_t count = size_t(1000000) * millions;
        const BoundingBox area = sdf->getSampleArea();
        const std::vector<glm::vec3> points = drawSamples(area, count);

        std::vector<float> approx(count), exact(count);
        const float usApprox = timedQuery(*sdf, points, approx);
        std::cout << "[info] Sdf us per query: " << usApprox << std::endl;
        const float usExact = timedQuery(*exactSdf, points, exact);
        std::cout << "[info] Exact Sdf us per query: " << usExact << std::endl;

        // sums in double, differences in float, like the reference's report (main.cpp:76-93)
        double sumSq = 0.0, sumAbs = 0.0;
        float worst = 0.0f;
        for (size_t k = 0; k < count; k++)
        {
            const float diff = approx[k] - exact[k];
            const float mag = glm::abs(diff);
            sumSq += static_cast<double>(diff * diff);
            sumAbs += static_cast<double>(mag);
            worst = glm::max(worst, mag);
        }
        std::cout << "[info] RMSE: " << std::sqrt(sumSq / static_cast<double>(count)) << std::endl;
        std::cout << "[info] MAE: " << sumAbs / static_cast<double>(count) << std::endl;
        std::cout << "[info] Max error: " << worst << std::endl;
    }
    catch (const std::exception& e)
    {
        std::cerr << "[error] " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
