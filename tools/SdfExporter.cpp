// SdfExporter — same command line as the reference's exporter (src/tools/SdfExporter/main.cpp:20-171), built on the
// drop-in classes of include/SdfLib (construction runs on the GPU):
//   SdfExporter model_path output_path [-d depth] [--start_depth n] [--sdf_format octree|exact_octree]
//       [--algorithm uniform|no_continuity|continuity] [--termination_rule trapezoidal_rule|simpsons_rule|by_distance_rule|none]
//       [--termination_threshold t] [--termination_threshold_by_distance t] [--min_triangles_per_node n]
//       [-n|--normalize] [--bb_margin percent] [--num_threads n]
// Defaults are the reference tool's (not the constructors'): octree depth 8 / start depth 1 / continuity / 1e-3,
// exact_octree depth 5 / start depth 1 / 32 triangles per node, margin 20 %. `--sdf_format grid` (UniformGridSdf) is
// outside the two hot paths and is refused.
#include <chrono>
#include <iostream>
#include <memory>

#include "SdfLib/OctreeSdf.h"
#include "SdfLib/ExactOctreeSdf.h"
#include "SdfLib/utils/Mesh.h"
#include "CliArgs.h"

using namespace sdflib;

static const char* kUsage =
    "  SdfExporter model_path output_path {OPTIONS}\n\n    SdfExporter export an sdf\n\n  OPTIONS:\n"
    "      -h, --help                        Display help menu\n"
    "      -d[depth], --depth=[depth]        The octree subdivision depth\n"
    "      --start_depth=[start_depth]       The octree start depth\n"
    "      --termination_rule=[rule]         trapezoidal_rule, simpsons_rule, by_distance_rule, none\n"
    "      --termination_threshold=[t]       Octree generation termination threshold\n"
    "      --termination_threshold_by_distance=[t]\n"
    "      --min_triangles_per_node=[n]      The minimum acceptable number of triangles per leaf in the octree\n"
    "      --sdf_format=[sdf_format]         octree, exact_octree\n"
    "      --algorithm=[algorithm]           uniform, no_continuity, continuity\n"
    "      -n, --normalize                   Normalize the model coordinates\n"
    "      --bb_margin=[bb_margin]           Percentage of margin added between the structure BB and the model BB\n"
    "      --num_threads=[num_threads]       Keeps the reference's layout-selecting meaning (< 2 / >= 2)\n";

int main(int argc, char** argv)
{
    CliArgs a;
    if (!a.parse(argc, argv,
                 {"--cell_size", "--depth", "--start_depth", "--termination_rule", "--termination_threshold", "--termination_threshold_by_distance",
                  "--min_triangles_per_node", "--sdf_format", "--algorithm", "--bb_margin", "--num_threads"},
                 {"--normalize"}, {{"-h", "--help"}, {"-d", "--depth"}, {"-c", "--cell_size"}, {"-n", "--normalize"}}))
    {
        std::cerr << kUsage;
        return 1;
    }
    if (a.help) { std::cerr << kUsage; return 0; }
    if (a.positionals.empty())
    {
        std::cerr << "Error: No model_path specified" << std::endl << kUsage;
        return 1;
    }
    const std::string sdfFormat = a.str("--sdf_format", "octree");
    const std::string modelPath = a.positionals[0];
    const std::string outputPath = a.positionals.size() > 1 ? a.positionals[1] : "../output/sdfOctreeBunny.bin";
    try
    {
        Mesh mesh(modelPath);
        BoundingBox box = mesh.getBoundingBox();
        if (a.flags.count("--normalize"))
        {
            // glm::scale(mat4(1), vec3(2 / maxSize)) * glm::translate(mat4(1), -centre), written out (main.cpp:84-92)
            const glm::vec3 boxSize = box.getSize();
            const float maxSize = glm::max(glm::max(boxSize.x, boxSize.y), boxSize.z);
            const float s = 2.0f / maxSize;
            const glm::vec3 c = box.getCenter();
            glm::mat4 m(1.0f);
            m[0][0] = s; m[1][1] = s; m[2][2] = s;
            m[3][0] = s * -c.x; m[3][1] = s * -c.y; m[3][2] = s * -c.z;
            mesh.applyTransform(m);
            box = mesh.getBoundingBox();
        }
        const glm::vec3 modelBBSize = box.getSize();
        const float margin = a.num("--bb_margin", 20.0f) / 100.0f;
        box.addMargin(margin * glm::max(glm::max(modelBBSize.x, modelBBSize.y), modelBBSize.z));

        const auto t0 = std::chrono::steady_clock::now();
        std::unique_ptr<SdfFunction> sdfFunc;
        if (sdfFormat == "octree")
        {
            const std::string algorithm = a.str("--algorithm", "continuity");
            const auto initAlgorithm = OctreeSdf::stringToInitAlgorithm(algorithm);
            if (!initAlgorithm)
            {
                std::cerr << algorithm << " is not a valid supported octree generation algorithm" << std::endl;
                return 0;
            }
            const auto rule = OctreeSdf::stringToTerminationRule(a.str("--termination_rule", "trapezoidal_rule"));
            if (!rule)
            {
                std::cerr << a.str("--termination_rule", "") << " is not a valid supported termination rule" << std::endl;
                return 0;
            }
            OctreeSdf::TerminationRuleParams params = OctreeSdf::TerminationRuleParams::setNoneRuleParams();
            if (*rule == OctreeSdf::TerminationRule::TRAPEZOIDAL_RULE || *rule == OctreeSdf::TerminationRule::SIMPSONS_RULE)
                params = OctreeSdf::TerminationRuleParams::setTrapezoidalRuleParams(a.num("--termination_threshold", 1e-3f));
            else if (*rule == OctreeSdf::TerminationRule::BY_DISTANCE_RULE)
                params = OctreeSdf::TerminationRuleParams::setByDistanceRuleParams(a.num("--termination_threshold", 1e-3f),
                                                                                  a.num("--termination_threshold_by_distance", 0.0f));
            sdfFunc.reset(new OctreeSdf(mesh, box, a.uint("--depth", 8), a.uint("--start_depth", 1), *rule, params, *initAlgorithm,
                                        a.uint("--num_threads", 1)));
        }
        else if (sdfFormat == "exact_octree")
        {
            sdfFunc.reset(new ExactOctreeSdf(mesh, box, a.uint("--depth", 5), a.uint("--start_depth", 1), a.uint("--min_triangles_per_node", 32),
                                             a.uint("--num_threads", 1)));
        }
        else
        {
            std::cerr << "The sdf_format can only be octree or exact_octree (grid: UniformGridSdf is not part of this library)" << std::endl;
            return 1;
        }
        std::cout << "[info] Computation time " << std::chrono::duration<float>(std::chrono::steady_clock::now() - t0).count() << "s" << std::endl;
        std::cout << "[info] Saving the model" << std::endl;
        if (!sdfFunc->saveToFile(outputPath)) { std::cerr << "[error] Cannot write " << outputPath << std::endl; return 1; }
    }
    catch (const std::exception& e)
    {
        std::cerr << "[error] " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
