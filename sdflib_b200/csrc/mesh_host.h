// Host-side mesh preprocessing: the float64 bounding-sphere BVH whose traversal order defines the reference's
// nearest-triangle choice (it stays on the host: its shape depends on std::sort's unstable tie order), the host
// restatement of TriangleData (sdfb200_triangle_data, and the A/B switch SDFB200_HOST_TRIANGLE_DATA; the builders use the
// device path of mesh_device.cu) and the non-manifold repair both TriangleData paths share.
#pragma once
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>
#include "tri_math.cuh"

namespace sdfb200 {

struct HostMesh {
    const f3* verts; uint32_t nVerts;
    const uint32_t* idx; uint32_t nIdx;
    uint32_t numTriangles() const { return nIdx / 3; }
};

// std::vector without value-initialisation: the big per-triangle arrays are written in full by parallel loops, and
// zero-filling tens of MB from one thread first (plus its page faults) cost more than the computation itself.
template <class T> struct DefaultInitAllocator : std::allocator<T> {
    template <class U> struct rebind { using other = DefaultInitAllocator<U>; };
    template <class U> void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
template <class T> using RawVec = std::vector<T, DefaultInitAllocator<T>>;
using TriVec = RawVec<TriData>;

int hostThreads();   // SDFB200_HOST_THREADS, else the machine's cores divided by the ranks of this node (LOCAL_WORLD_SIZE)
void setHostThreadsForThisThread(int n);   // > 0: overrides hostThreads() for calls made by this thread (0 = back to the default)

// reference: TriangleUtils::calculateMeshTriangleData, src/utils/TriangleUtils.cpp:7-428
TriVec computeTriangleData(const HostMesh& mesh);

// Non-manifold repair over the edge uses the pairing left open (shared by the host and the device path; mesh_host.cpp)
struct OpenEdgeUse { uint32_t lo, hi, corner; };   // vertex ids (lo <= hi), corner = 3 * triangle + edge
struct OpenEdgeRepair {
    std::vector<uint32_t> edgeCorner; std::vector<f3> edgeNormal;   // edgesNormal[corner % 3] of triangle corner / 3, in its frame, in application order
    std::vector<uint32_t> vertex; std::vector<f3> vertexNormal;     // final (merged) vertex normals, world frame
};
OpenEdgeRepair repairOpenEdges(const HostMesh& mesh, const std::vector<OpenEdgeUse>& openInKeyOrder, const f3* vNormal);

// Device-friendly BVH node: both child spheres + child links, 80 bytes, 16-byte aligned
// (reference: tmd::TriangleMeshDistance::Node, TriangleMeshDistance.h:96-109).
// pad[0] != 0 marks a leaf node (`right` = triangle id). In inner nodes a link >= 0 is the index of an inner
// child and a link < 0 is ~triangleId of a leaf child, so the traversal never has to load a leaf node.
struct alignas(16) BvhNode {
    double lc[3], lr;
    double rc[3], rr;
    int32_t left, right;
    int32_t pad[2];
};
static_assert(sizeof(BvhNode) == 80, "BvhNode layout");

// reference: tmd::TriangleMeshDistance::_build_tree, TriangleMeshDistance.h:421-490.
// Node ids equal the reference's push order (pre-order: a subtree over n triangles owns 2n-1
// consecutive ids), which lets independent subtrees be built by parallel host tasks.
RawVec<BvhNode> buildBvh(const HostMesh& mesh);

}  // namespace sdfb200
