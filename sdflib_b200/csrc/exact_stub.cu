// placeholder until exact_build.cu / exact_query.cu land
#include "sdf_internal.h"
namespace sdfb200 {
void buildExactOnDevice(sdfb200_sdf&, const HostMesh&, const float*, uint32_t, uint32_t, uint32_t, uint32_t) {
    throw Error(SDFB200_ERR_UNSUPPORTED, "ExactOctreeSdf construction is not built yet");
}
void launchExactQuery(const sdfb200_sdf&, const float*, uint64_t, float*, float*, cudaStream_t) {
    throw Error(SDFB200_ERR_UNSUPPORTED, "ExactOctreeSdf query is not built yet");
}
}
