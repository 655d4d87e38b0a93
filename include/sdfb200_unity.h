/* The 12 C exports of the reference's Unity plugin (src/tools/SdfLibUnity/SdfExportFunc.h:16-58), same names,
 * argument order and meaning, implemented on top of include/sdfb200.h by sdflib_b200/csrc/unity_abi.cpp and built
 * as sdflib_b200/libSdfLibUnity.so (the reference builds a library of that name from SdfExportFunc.cpp).
 * `SdfFunction*` of the reference is an opaque handle here; glm::vec3 is three packed floats. */
#ifndef SDFB200_UNITY_H
#define SDFB200_UNITY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdfb200_sdf SdfFunctionHandle;
typedef struct sdfb200_vec3 { float x, y, z; } sdfb200_vec3;

void saveSdf(SdfFunctionHandle* sdfPointer, char* path);                                   /* SdfExportFunc.cpp:43-46 */
SdfFunctionHandle* loadSdf(char* path);                                                    /* :48-51 (NULL on failure) */
SdfFunctionHandle* createExactOctreeSdf(sdfb200_vec3* vertices, uint32_t numVertices, uint32_t* indices, uint32_t numIndices,
                                        float bbMinX, float bbMinY, float bbMinZ, float bbMaxX, float bbMaxY, float bbMaxZ,
                                        uint32_t startOctreeDepth, uint32_t maxOctreeDepth, uint32_t minTrianglesPerNode,
                                        uint32_t numThreads);                              /* :53-80 */
SdfFunctionHandle* createOctreeSdf(sdfb200_vec3* vertices, uint32_t numVertices, uint32_t* indices, uint32_t numIndices,
                                   float bbMinX, float bbMinY, float bbMinZ, float bbMaxX, float bbMaxY, float bbMaxZ,
                                   uint32_t startOctreeDepth, uint32_t maxOctreeDepth, float maxError,
                                   uint32_t numThreads);                                   /* :82-110 */
float getDistance(SdfFunctionHandle* sdfPointer, float pointX, float pointY, float pointZ);                 /* :112-115 */
float getDistanceAndGradient(SdfFunctionHandle* sdfPointer, float pointX, float pointY, float pointZ,
                             sdfb200_vec3* outGradient);                                                    /* :117-120 */
sdfb200_vec3 getBBMinPoint(SdfFunctionHandle* sdfPointer);                                 /* :122-125 */
sdfb200_vec3 getBBSize(SdfFunctionHandle* sdfPointer);                                     /* :127-130 */
uint32_t getStartGridSize(SdfFunctionHandle* sdfPointer);                                  /* :132-137 (0 unless OCTREE) */
uint32_t getOctreeDataSize(SdfFunctionHandle* sdfPointer);                                 /* :139-147 (#uint32 words, 0 unless OCTREE) */
void getOctreeData(SdfFunctionHandle* sdfPointer, uint32_t* data);                         /* :149-159 (OCTREE only) */
void deleteSdf(SdfFunctionHandle* sdfPointer);                                             /* :161-173 (here: frees every format) */

#ifdef __cplusplus
}
#endif
#endif
