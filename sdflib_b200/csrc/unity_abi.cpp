// The Unity plugin's C exports (include/sdfb200_unity.h) on top of the C-ABI. Behaviour follows
// src/tools/SdfLibUnity/SdfExportFunc.cpp:43-182 of the reference, including the parts that matter to callers:
// createOctreeSdf asks for InitAlgorithm::CONTINUITY (:101-103), getStartGridSize / getOctreeDataSize /
// getOctreeData answer for OctreeSdf only. Differences: deleteSdf frees every format (the reference leaks all but
// EXACT_OCTREE), and errors return NULL / 0 instead of crashing.
#include <cstdio>

#include "../../include/sdfb200.h"
#include "../../include/sdfb200_unity.h"

extern "C" {

void saveSdf(SdfFunctionHandle* sdf, char* path) {
    if (sdf && path && sdfb200_save(sdf, path) != SDFB200_OK) std::fprintf(stderr, "[sdfb200] saveSdf: %s\n", sdfb200_last_error());
}

SdfFunctionHandle* loadSdf(char* path) {
    sdfb200_sdf* h = nullptr;
    if (!path || sdfb200_load(path, &h) != SDFB200_OK) {
        std::fprintf(stderr, "[sdfb200] loadSdf: %s\n", sdfb200_last_error());
        return nullptr;
    }
    return h;
}

SdfFunctionHandle* createExactOctreeSdf(sdfb200_vec3* vertices, uint32_t numVertices, uint32_t* indices, uint32_t numIndices,
                                        float bbMinX, float bbMinY, float bbMinZ, float bbMaxX, float bbMaxY, float bbMaxZ,
                                        uint32_t startOctreeDepth, uint32_t maxOctreeDepth, uint32_t minTrianglesPerNode,
                                        uint32_t numThreads) {
    const float box[6] = {bbMinX, bbMinY, bbMinZ, bbMaxX, bbMaxY, bbMaxZ};
    sdfb200_sdf* h = nullptr;
    if (sdfb200_build_exact(&vertices->x, numVertices, indices, numIndices, box, maxOctreeDepth, startOctreeDepth,
                            minTrianglesPerNode, numThreads, &h) != SDFB200_OK) {
        std::fprintf(stderr, "[sdfb200] createExactOctreeSdf: %s\n", sdfb200_last_error());
        return nullptr;
    }
    return h;
}

SdfFunctionHandle* createOctreeSdf(sdfb200_vec3* vertices, uint32_t numVertices, uint32_t* indices, uint32_t numIndices,
                                   float bbMinX, float bbMinY, float bbMinZ, float bbMaxX, float bbMaxY, float bbMaxZ,
                                   uint32_t startOctreeDepth, uint32_t maxOctreeDepth, float maxError, uint32_t numThreads) {
    const float box[6] = {bbMinX, bbMinY, bbMinZ, bbMaxX, bbMaxY, bbMaxZ};
    sdfb200_sdf* h = nullptr;
    int code = sdfb200_build_octree(&vertices->x, numVertices, indices, numIndices, box, maxOctreeDepth, startOctreeDepth,
                                    SDFB200_RULE_TRAPEZOIDAL, maxError, 0.0f, SDFB200_ALG_CONTINUITY, numThreads, &h);
    if (code != SDFB200_OK) {
        std::fprintf(stderr, "[sdfb200] createOctreeSdf: %s\n", sdfb200_last_error());
        return nullptr;
    }
    return h;
}

float getDistance(SdfFunctionHandle* sdf, float x, float y, float z) {
    const float p[3] = {x, y, z};
    float d = 0.0f;
    sdfb200_query(sdf, p, 1, &d, nullptr, 0, nullptr);
    return d;
}

float getDistanceAndGradient(SdfFunctionHandle* sdf, float x, float y, float z, sdfb200_vec3* outGradient) {
    const float p[3] = {x, y, z};
    float d = 0.0f, g[3] = {0.0f, 0.0f, 0.0f};
    sdfb200_query(sdf, p, 1, &d, g, 0, nullptr);
    if (outGradient) { outGradient->x = g[0]; outGradient->y = g[1]; outGradient->z = g[2]; }
    return d;
}

sdfb200_vec3 getBBMinPoint(SdfFunctionHandle* sdf) {
    sdfb200_info i;
    sdfb200_vec3 r = {0.0f, 0.0f, 0.0f};
    if (sdfb200_get_info(sdf, &i) == SDFB200_OK) { r.x = i.box_min[0]; r.y = i.box_min[1]; r.z = i.box_min[2]; }
    return r;
}

sdfb200_vec3 getBBSize(SdfFunctionHandle* sdf) {
    sdfb200_info i;
    sdfb200_vec3 r = {0.0f, 0.0f, 0.0f};
    if (sdfb200_get_info(sdf, &i) == SDFB200_OK) { r.x = i.box_max[0] - i.box_min[0]; r.y = i.box_max[1] - i.box_min[1]; r.z = i.box_max[2] - i.box_min[2]; }
    return r;
}

uint32_t getStartGridSize(SdfFunctionHandle* sdf) {
    sdfb200_info i;
    return (sdfb200_get_info(sdf, &i) == SDFB200_OK && i.format == SDFB200_FORMAT_OCTREE) ? uint32_t(i.start_grid_size) : 0u;
}

uint32_t getOctreeDataSize(SdfFunctionHandle* sdf) {
    sdfb200_info i;
    return (sdfb200_get_info(sdf, &i) == SDFB200_OK && i.format == SDFB200_FORMAT_OCTREE) ? uint32_t(i.octree_words) : 0u;
}

void getOctreeData(SdfFunctionHandle* sdf, uint32_t* data) {
    sdfb200_info i;
    if (data && sdfb200_get_info(sdf, &i) == SDFB200_OK && i.format == SDFB200_FORMAT_OCTREE) sdfb200_get_octree_data(sdf, data, i.octree_words);
}

void deleteSdf(SdfFunctionHandle* sdf) { sdfb200_free(sdf); }

}  // extern "C"
