// Prints what sdflib::Mesh(path) read: "<numVertices> <numIndices>" then every vertex and every index, so that the
// test can compare with what it wrote (tests/test_tools.py). Host-only: no device needed.
#include <cstdio>
#include <SdfLib/utils/Mesh.h>

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    try
    {
        sdflib::Mesh mesh{std::string(argv[1])};
        std::printf("%zu %zu\n", mesh.getVertices().size(), mesh.getIndices().size());
        for (const glm::vec3& v : mesh.getVertices()) std::printf("%.9g %.9g %.9g\n", v.x, v.y, v.z);
        for (uint32_t i : mesh.getIndices()) std::printf("%u\n", i);
        const sdflib::BoundingBox b = mesh.getBoundingBox();
        std::printf("%.9g %.9g %.9g %.9g %.9g %.9g\n", b.min.x, b.min.y, b.min.z, b.max.x, b.max.y, b.max.z);
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
