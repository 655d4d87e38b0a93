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

        const uint32_t numSamples = 1000000u * (argc > 3 ? uint32_t(std::strtoul(argv[3], nullptr, 10)) : 1u);
        std::vector<glm::vec3> samples(numSamples);
        const glm::vec3 center = sdf->getSampleArea().getCenter();
        const glm::vec3 size = sdf->getSampleArea().getSize() - glm::vec3(1e-5);
        auto getRandomSample = [&]() -> glm::vec3
        {
            // one rand() per component, in x, y, z order (the reference's braced initialiser is sequenced left to right)
            const float x = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
            const float y = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
            const float z = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
            return center + (glm::vec3(x, y, z) - 0.5f) * size;
        };
        std::generate(samples.begin(), samples.end(), getRandomSample);

        std::vector<float> sdfDist(numSamples), exactSdfDist(numSamples);
        auto t0 = std::chrono::steady_clock::now();
        sdf->getDistances(samples.data(), numSamples, sdfDist.data());
        float seconds = std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count();
        std::cout << "[info] Sdf us per query: " << seconds * 1.0e6f / static_cast<float>(numSamples) << std::endl;
        t0 = std::chrono::steady_clock::now();
        exactSdf->getDistances(samples.data(), numSamples, exactSdfDist.data());
        seconds = std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count();
        std::cout << "[info] Exact Sdf us per query: " << seconds * 1.0e6f / static_cast<float>(numSamples) << std::endl;

        double rmseError = 0.0, maeError = 0.0;
        float maxError = 0.0f;
        for (uint32_t s = 0; s < numSamples; s++)
        {
            const float d = sdfDist[s] - exactSdfDist[s];
            rmseError += static_cast<double>(d * d);
            maeError += static_cast<double>(glm::abs(d));
            maxError = glm::max(maxError, glm::abs(d));
        }
        rmseError = std::sqrt(rmseError / static_cast<double>(numSamples));
        maeError = maeError / static_cast<double>(numSamples);
        std::cout << "[info] RMSE: " << rmseError << std::endl;
        std::cout << "[info] MAE: " << maeError << std::endl;
        std::cout << "[info] Max error: " << maxError << std::endl;
    }
    catch (const std::exception& e)
    {
        std::cerr << "[error] " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
