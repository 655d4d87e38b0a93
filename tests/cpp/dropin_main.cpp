// Drop-in check of the C++ surface (include/SdfLib/*.h): the reference's README usage (README.md:53-133) compiled
// against our headers and libsdfb200.so. Writes what it computed to a flat binary file that the Python test compares
// with the ctypes binding's results for the same inputs.
//   usage: dropin_main <out.bin> <tmpdir>
#include <cstdio>
#include <cstring>
#include <vector>

#include <SdfLib/SdfFunction.h>
#include <SdfLib/ExactOctreeSdf.h>
#include <SdfLib/OctreeSdf.h>

using namespace sdflib;

template <class T> static void put(FILE* f, const std::vector<T>& v) {
    const uint64_t n = v.size();
    fwrite(&n, 8, 1, f);
    fwrite(v.data(), sizeof(T), v.size(), f);
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    try {
        // mesh: the reference's icosphere (PrimitivesFactory::getIsosphere(2)) pushed off-centre
        uint32_t nv = 0, ni = 0;
        sdfb200_make_isosphere(2, nullptr, nullptr, &nv, &ni);
        std::vector<glm::vec3> verts(nv);
        std::vector<uint32_t> idx(ni);
        sdfb200_make_isosphere(2, &verts[0].x, idx.data(), &nv, &ni);
        for (glm::vec3& v : verts) v = v * glm::vec3(1.0f, 0.9f, 1.1f) + glm::vec3(0.013f, -0.007f, 0.003f);
        Mesh mesh(verts.data(), nv, idx.data(), ni);

        BoundingBox box = mesh.getBoundingBox();
        const glm::vec3 modelBBSize = box.getSize();
        box.addMargin(0.2f * glm::max(glm::max(modelBBSize.x, modelBBSize.y), modelBBSize.z));

        ExactOctreeSdf exactSdf(mesh, box, 5, 2, 16, 8);
        OctreeSdf octreeSdf(mesh, box, 5, 2, 1e-3, OctreeSdf::InitAlgorithm::NO_CONTINUITY, 8);

        const std::string dir(argv[2]);
        if (!exactSdf.saveToFile(dir + "/exact.bin") || !octreeSdf.saveToFile(dir + "/octree.bin")) return 3;
        std::unique_ptr<SdfFunction> loaded = SdfFunction::loadFromFile(dir + "/octree.bin");
        if (!loaded || loaded->getFormat() != SdfFunction::SdfFormat::OCTREE) return 4;
        if (SdfFunction::loadFromFile(dir + "/does_not_exist.bin") != nullptr) return 5;

        // queries: scalar virtuals and the bulk entry must agree
        std::vector<glm::vec3> pts;
        const BoundingBox area = octreeSdf.getGridBoundingBox();
        for (int k = 0; k < 12; k++)
            for (int j = 0; j < 12; j++)
                for (int i = 0; i < 12; i++)
                    pts.push_back(area.min + glm::vec3((i + 0.37f) / 11.5f, (j + 0.41f) / 11.5f, (k + 0.29f) / 11.5f) * area.getSize() - glm::vec3(0.02f));
        std::vector<float> dOct(pts.size()), dEx(pts.size());
        std::vector<glm::vec3> gOct(pts.size()), gEx(pts.size());
        octreeSdf.getDistances(pts.data(), pts.size(), dOct.data(), gOct.data());
        exactSdf.getDistances(pts.data(), pts.size(), dEx.data(), gEx.data());
        for (size_t q = 0; q < pts.size(); q += 97) {
            glm::vec3 g;
            if (octreeSdf.getDistance(pts[q]) != dOct[q] || loaded->getDistance(pts[q], g) != dOct[q]) return 6;
            if (std::memcmp(&g, &gOct[q], 12) != 0) return 7;
            if (exactSdf.getDistance(pts[q], g) != dEx[q] || std::memcmp(&g, &gEx[q], 12) != 0) return 8;
        }

        // TriangleUtils::calculateMeshTriangleData through the drop-in header equals what ExactOctreeSdf keeps
        const std::vector<TriangleUtils::TriangleData> td = TriangleUtils::calculateMeshTriangleData(mesh);
        if (td.size() != exactSdf.getTrianglesData().size() ||
            std::memcmp(td.data(), exactSdf.getTrianglesData().data(), td.size() * sizeof(TriangleUtils::TriangleData)) != 0) return 11;

        // getDepthDensity (OctreeSdf.cpp:232-277): the leaves tile the unit cube, so the densities sum to exactly 1
        std::vector<float> density;
        octreeSdf.getDepthDensity(density);
        double covered = 0.0;
        for (float d : density) covered += d;
        if (density.size() != octreeSdf.getOctreeMaxDepth() + 1 || covered < 0.999999 || covered > 1.000001) return 10;

        FILE* f = fopen(argv[1], "wb");
        if (!f) return 9;
        put(f, verts); put(f, idx);
        std::vector<float> b = {box.min.x, box.min.y, box.min.z, box.max.x, box.max.y, box.max.z};
        put(f, b); put(f, pts); put(f, dOct); put(f, gOct); put(f, dEx); put(f, gEx);
        put(f, octreeSdf.getOctreeData()); put(f, exactSdf.getOctreeData()); put(f, exactSdf.getTrianglesSets());
        put(f, exactSdf.getTrianglesMasks());
        std::vector<float> hdr = {octreeSdf.getOctreeValueRange(), octreeSdf.getOctreeMinBorderValue(), float(octreeSdf.getStartGridSize().x),
                                  float(octreeSdf.getOctreeMaxDepth()), float(exactSdf.getMaxTrianglesInLeafs()), float(exactSdf.getMinTrianglesInLeafs())};
        put(f, hdr);
        fclose(f);
        std::printf("dropin ok: %zu octree words, %zu exact nodes, %zu queries\n", octreeSdf.getOctreeData().size(),
                    exactSdf.getOctreeData().size(), pts.size());
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "dropin failed: %s\n", e.what());
        return 1;
    }
}
