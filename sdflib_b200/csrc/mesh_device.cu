// Mesh ingestion on the device: TriangleData for all triangles, the ExactOctreeSdf side arrays
// and the upload of the host-built BVH, packaged as a PreparedMesh that every builder reads and that can be replicated
// on other devices / ranks without repeating the work.
//
// Reference: TriangleUtils::calculateMeshTriangleData, src/utils/TriangleUtils.cpp:7-428. The reference walks the
// triangles serially through a std::map of edges; the result only depends on WHICH corners pair up and on the order
// of the float additions, so it is restated as data-parallel passes that keep both:
//   1. trianglePassKernel   one thread per triangle: the TriangleData constructor (tri_data_build.cuh, shared with the
//                           host path), angle * normal of the three corners (acosfLibm = the host libm's acosf bit
//                           for bit) and the (min, max) vertex key of the three edges
//   2. radix sort of the 3T edge uses by key (stable: equal keys stay in corner order) + edgePairKernel: uses of one
//      key pair up (1st, 2nd), (3rd, 4th), ... exactly like the map's insert / erase cycle (:60-83); both triangles of
//      a pair get n_t + n_t' in their own frame; an odd use is OPEN (appended to a list)
//   3. radix sort of the 3T corners by vertex id (stable: ascending corner order inside a vertex) + vertexNormalKernel:
//      one thread per vertex adds its corner contributions in that order (:85-86: float addition order matters)
//   4. only when open edges exist (non-manifold input): the reference's repair by vertex merging on two staggered
//      2048^3 hash grids (:292-420) runs on the HOST over the open uses alone (mesh_host.cpp, repairOpenEdges — a
//      serial std::map algorithm over what is normally an empty list) and its patches are scattered back
//   5. toFrameKernel        vertex normals into every triangle's frame
// The sorts are cub::DeviceRadixSort (library code for a plain sort, like cuBLAS for a plain GEMM).
// Bit parity with the host path / the oracle: tests/test_gpu_mesh.py (manifold and random non-manifold meshes).
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <thread>
#include <type_traits>

#include "device_utils.cuh"
#include "sdf_internal.h"
#include "tri_data_build.cuh"

namespace sdfb200 {

namespace {

__global__ void __launch_bounds__(128)
trianglePassKernel(const f3* __restrict__ verts, const uint32_t* __restrict__ idx, uint32_t nT, int keyShift, TriData* __restrict__ tris,
                   f3* __restrict__ cornerN, uint64_t* __restrict__ edgeKey, uint32_t* __restrict__ cornerId) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nT) return;
    const uint32_t ix[3] = {idx[3 * size_t(t)], idx[3 * size_t(t) + 1], idx[3 * size_t(t) + 2]};
    const f3 p[3] = {verts[ix[0]], verts[ix[1]], verts[ix[2]]};
    const TriData d = makeTriData(p[0], p[1], p[2]);
    tris[t] = d;
    const f3 n = triNormal(d);
#pragma unroll
    for (uint32_t k = 0; k < 3; k++) {
        const uint32_t a = ix[k], b = ix[(k + 1) % 3];
        const f3 pa = p[k], pb = p[(k + 1) % 3], pc = p[(k + 2) % 3];
        const float cosA = dot3(normalize3(pb - pa), normalize3(pc - pa));
        const float angle = acosfLibm(gmin(gmax(cosA, -1.0f), 1.0f));
        cornerN[3 * size_t(t) + k] = angle * n;
        edgeKey[3 * size_t(t) + k] = (uint64_t(min(a, b)) << keyShift) | max(a, b);
        cornerId[3 * size_t(t) + k] = 3 * t + k;
    }
}

// sorted edge uses -> pairs; `open` receives the sorted positions of the unpaired uses
__global__ void edgePairKernel(const uint64_t* __restrict__ key, const uint32_t* __restrict__ corner, uint64_t n, TriData* __restrict__ tris,
                               uint32_t* __restrict__ openCount, uint32_t* __restrict__ open) {
    const uint64_t p = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint64_t k = key[p];
    uint64_t g = p;
    while (g > 0 && key[g - 1] == k) g--;
    if ((p - g) & 1) return;   // second of a pair: written by its partner
    if (p + 1 < n && key[p + 1] == k) {
        const uint32_t first = corner[p], second = corner[p + 1];
        const uint32_t t2 = first / 3, t = second / 3;
        const f3 nn = triNormal(tris[t]) + triNormal(tris[t2]);
        const f3 a = matMul(tris[t].T, nn), b = matMul(tris[t2].T, nn);
        float* ea = tris[t].edgesNormal[second % 3];
        float* eb = tris[t2].edgesNormal[first % 3];
        // same store order as the reference (:78-81): when both corners are the same slot the second store wins
        ea[0] = a.x; ea[1] = a.y; ea[2] = a.z;
        eb[0] = b.x; eb[1] = b.y; eb[2] = b.z;
    } else {
        open[atomicAdd(openCount, 1u)] = uint32_t(p);
    }
}

__global__ void vertexNormalKernel(const uint32_t* __restrict__ vertexOf, const uint32_t* __restrict__ corner, uint64_t n,
                                   const f3* __restrict__ cornerN, f3* __restrict__ vNormal) {
    const uint64_t p = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t v = vertexOf[p];
    if (p > 0 && vertexOf[p - 1] == v) return;   // one thread per vertex: the one at the head of its run
    f3 acc = mk3(0.0f, 0.0f, 0.0f);
    for (uint64_t q = p; q < n && vertexOf[q] == v; q++) acc = acc + cornerN[corner[q]];
    vNormal[v] = acc;
}

__global__ void applyEdgePatches(const uint32_t* corner, const f3* value, uint32_t n, TriData* tris) {   // serial: patch order is part of the result
    if (blockIdx.x || threadIdx.x) return;
    for (uint32_t i = 0; i < n; i++) {
        float* e = tris[corner[i] / 3].edgesNormal[corner[i] % 3];
        e[0] = value[i].x; e[1] = value[i].y; e[2] = value[i].z;
    }
}
__global__ void applyVertexPatches(const uint32_t* vertex, const f3* value, uint32_t n, f3* vNormal) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vNormal[vertex[i]] = value[i];
}

__global__ void toFrameKernel(const uint32_t* __restrict__ idx, const f3* __restrict__ vNormal, uint64_t n, TriData* __restrict__ tris) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f3 r = matMul(tris[i / 3].T, vNormal[idx[i]]);
    float* dst = tris[i / 3].verticesNormal[i % 3];
    dst[0] = r.x; dst[1] = r.y; dst[2] = r.z;
}

// ---- ExactOctreeSdf side arrays --------------------------------------------------------------------------------
__global__ void framesKernel(const TriData* tris, float4* frames, uint32_t n) {   // first 19 floats of TriangleData, padded to 5 x float4
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 5u) return;
    const uint32_t t = i / 5u, q = i % 5u;
    const float* src = reinterpret_cast<const float*>(tris + t) + 4 * q;
    frames[i] = make_float4(src[0], src[1], src[2], q == 4 ? 0.0f : src[3]);
}
__global__ void validFlagKernel(const TriData* tris, uint32_t n, uint8_t* flag) {   // ExactOctreeSdfDepthFirst.h:106 (false for NaN)
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const f3 nrm = triNormal(tris[t]);
    flag[t] = dot3(nrm, nrm) > 1e-3f ? 1 : 0;
}
__global__ void validScatterKernel(const uint8_t* flag, const uint32_t* pos, uint32_t n, uint32_t* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && flag[t]) out[pos[t]] = t;
}

__global__ void gatherTriVertsKernel(const f3* verts, const uint32_t* idx, uint32_t nTris, float4* triVerts) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTris * 3) return;
    const f3 v = verts[idx[t]];
    triVerts[t] = make_float4(v.x, v.y, v.z, 0.f);
}

template <class K> void sortPairs(const K* keyIn, K* keyOut, const uint32_t* valIn, uint32_t* valOut, uint64_t n, int endBit, DevBuf<uint8_t>& temp) {
    if (n >= (uint64_t(1) << 31)) throw Error(SDFB200_ERR_INVALID, "mesh too large for the device edge sort (more than 2^31 corners)");
    size_t bytes = 0;
    SDFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keyIn, keyOut, valIn, valOut, int(n), 0, endBit));
    if (temp.n < bytes) temp.alloc(bytes);
    SDFB_CUDA(cub::DeviceRadixSort::SortPairs(temp.p, bytes, keyIn, keyOut, valIn, valOut, int(n), 0, endBit));
}

int bitsFor(uint32_t n) {   // bits needed for values 0 .. n - 1 (at least 1)
    int b = 1;
    while (b < 32 && (uint64_t(1) << b) < n) b++;
    return b;
}

}  // namespace

// TriangleData of `mesh` into pm.dev.tris (verts / idx already on the device). Default stream, synchronises.
static void triangleDataOnDevice(PreparedMesh& pm, const HostMesh& mesh) {
    const uint32_t nT = pm.nTris;
    const uint64_t nC = uint64_t(nT) * 3;
    pm.dev.tris.alloc(nT);
    DevBuf<f3> cornerN(nC), vNormal(pm.nVerts);
    DevBuf<uint64_t> edgeKey(nC), edgeKeySorted(nC);
    DevBuf<uint32_t> cornerId(nC), cornerSorted(nC), vertSorted(nC), cornerByVert(nC), openCount(1), open(nC);
    DevBuf<uint8_t> temp;
    const int vb = bitsFor(pm.nVerts);
    trianglePassKernel<<<divUp(nT, 128), 128>>>(pm.dev.verts.p, pm.dev.idx.p, nT, vb, pm.dev.tris.p, cornerN.p, edgeKey.p, cornerId.p);
    SDFB_CUDA(cudaGetLastError());
    sortPairs<uint64_t>(edgeKey.p, edgeKeySorted.p, cornerId.p, cornerSorted.p, nC, 2 * vb, temp);
    SDFB_CUDA(cudaMemsetAsync(openCount.p, 0, 4));
    edgePairKernel<<<divUp(nC, 256), 256>>>(edgeKeySorted.p, cornerSorted.p, nC, pm.dev.tris.p, openCount.p, open.p);
    sortPairs<uint32_t>(pm.dev.idx.p, vertSorted.p, cornerId.p, cornerByVert.p, nC, vb, temp);
    SDFB_CUDA(cudaMemsetAsync(vNormal.p, 0, size_t(pm.nVerts) * sizeof(f3)));
    vertexNormalKernel<<<divUp(nC, 256), 256>>>(vertSorted.p, cornerByVert.p, nC, cornerN.p, vNormal.p);
    SDFB_CUDA(cudaGetLastError());
    uint32_t nOpen = 0;
    SDFB_CUDA(cudaMemcpy(&nOpen, openCount.p, 4, cudaMemcpyDeviceToHost));
    if (nOpen) {
        // non-manifold input: the reference's repair over the open uses, on the host (serial std::map algorithm)
        std::vector<uint32_t> pos(nOpen);
        open.download(pos.data(), nOpen);
        SDFB_CUDA(cudaDeviceSynchronize());
        std::sort(pos.begin(), pos.end());   // sorted positions = key order (open keys are unique)
        std::vector<uint64_t> keys(nOpen);
        std::vector<uint32_t> corners(nOpen);
        {   // gather the few keys / corners (contiguous ranges are rare: one small copy per use would be slow, so batch by download of both arrays when many)
            if (uint64_t(nOpen) * 16 > nC) {
                std::vector<uint64_t> allK(nC);
                std::vector<uint32_t> allC(nC);
                edgeKeySorted.download(allK.data(), nC);
                cornerSorted.download(allC.data(), nC);
                SDFB_CUDA(cudaDeviceSynchronize());
                for (uint32_t i = 0; i < nOpen; i++) { keys[i] = allK[pos[i]]; corners[i] = allC[pos[i]]; }
            } else {
                for (uint32_t i = 0; i < nOpen; i++) {
                    SDFB_CUDA(cudaMemcpyAsync(&keys[i], edgeKeySorted.p + pos[i], 8, cudaMemcpyDeviceToHost));
                    SDFB_CUDA(cudaMemcpyAsync(&corners[i], cornerSorted.p + pos[i], 4, cudaMemcpyDeviceToHost));
                }
                SDFB_CUDA(cudaDeviceSynchronize());
            }
        }
        std::vector<OpenEdgeUse> uses(nOpen);
        const uint64_t lowMask = (uint64_t(1) << vb) - 1;
        for (uint32_t i = 0; i < nOpen; i++) uses[i] = OpenEdgeUse{uint32_t(keys[i] >> vb), uint32_t(keys[i] & lowMask), corners[i]};
        std::vector<f3> hostVNormal(pm.nVerts);
        vNormal.download(hostVNormal.data(), pm.nVerts);
        SDFB_CUDA(cudaDeviceSynchronize());
        OpenEdgeRepair rep = repairOpenEdges(mesh, uses, hostVNormal.data());
        if (!rep.edgeCorner.empty()) {
            DevBuf<uint32_t> dc(rep.edgeCorner.size());
            DevBuf<f3> dv(rep.edgeNormal.size());
            dc.upload(rep.edgeCorner.data(), rep.edgeCorner.size());
            dv.upload(rep.edgeNormal.data(), rep.edgeNormal.size());
            applyEdgePatches<<<1, 32>>>(dc.p, dv.p, uint32_t(rep.edgeCorner.size()), pm.dev.tris.p);
            SDFB_CUDA(cudaDeviceSynchronize());
        }
        if (!rep.vertex.empty()) {
            DevBuf<uint32_t> dc(rep.vertex.size());
            DevBuf<f3> dv(rep.vertexNormal.size());
            dc.upload(rep.vertex.data(), rep.vertex.size());
            dv.upload(rep.vertexNormal.data(), rep.vertexNormal.size());
            applyVertexPatches<<<divUp(rep.vertex.size(), 256), 256>>>(dc.p, dv.p, uint32_t(rep.vertex.size()), vNormal.p);
            SDFB_CUDA(cudaDeviceSynchronize());
        }
    }
    toFrameKernel<<<divUp(nC, 256), 256>>>(pm.dev.idx.p, vNormal.p, nC, pm.dev.tris.p);
    SDFB_CUDA(cudaGetLastError());
    SDFB_CUDA(cudaDeviceSynchronize());
}

static void exactPartsOnDevice(PreparedMesh& pm) {
    const uint32_t nT = pm.nTris;
    pm.frames.alloc(size_t(nT) * 5);
    framesKernel<<<divUp(uint64_t(nT) * 5, 256), 256>>>(pm.dev.tris.p, pm.frames.p, nT);
    DevBuf<uint8_t> flag(nT);
    DevBuf<uint32_t> pos(size_t(nT) + 1);
    validFlagKernel<<<divUp(nT, 256), 256>>>(pm.dev.tris.p, nT, flag.p);
    FlagScanner scan;
    pm.numValid = scan.run(flag.p, pos.p, nT);
    pm.valid.alloc(size_t(pm.numValid) + 8);   // + 8: the TMA window of the sample kernel may read past the end
    SDFB_CUDA(cudaMemsetAsync(pm.valid.p, 0, (size_t(pm.numValid) + 8) * 4));
    validScatterKernel<<<divUp(nT, 256), 256>>>(flag.p, pos.p, nT, pm.valid.p);
    SDFB_CUDA(cudaGetLastError());
    SDFB_CUDA(cudaDeviceSynchronize());
    pm.hasExactParts = true;
}

static void gatherTriVerts(MeshOnDevice& m) {   // one 48-byte record per triangle: what the BVH build and the traversal's leaf test read
    m.triVerts.alloc(size_t(m.numTriangles) * 3);
    gatherTriVertsKernel<<<divUp(uint64_t(m.numTriangles) * 3, 256), 256>>>(m.verts.p, m.idx.p, m.numTriangles, m.triVerts.p);
    SDFB_CUDA(cudaGetLastError());
}

static void uploadBvh(PreparedMesh& pm, const RawVec<BvhNode>& bvh) {
    MeshOnDevice& m = pm.dev;
    m.bvh.alloc(bvh.size());
    m.bvh.upload(bvh.data(), bvh.size());
    m.rootLink = bvh[0].pad[0] ? ~bvh[0].right : 0;   // single-triangle mesh: the root is a leaf
    // height of the median-split tree (mesh_host.cpp: halves of floor / ceil size) = deepest possible stack, + 1 spare
    uint32_t n = m.numTriangles, h = 0;
    while (n > 1) { n = n - n / 2; h++; }
    m.stackDepth = int(h) + 1;
    gatherTriVerts(m);
    SDFB_CUDA(cudaDeviceSynchronize());
    pm.hasBvh = true;
}

bool hostBvhRequested() {   // A/B switch: the host builder of mesh_host.cpp (the only one until the end of round 2)
    static const bool host = std::getenv("SDFB200_HOST_BVH") != nullptr;
    return host;
}

std::shared_ptr<PreparedMesh> prepareMesh(const HostMesh& mesh, bool withBvh, bool withExactParts) {
    std::shared_ptr<PreparedMesh> pm(new PreparedMesh());
    SDFB_CUDA(cudaGetDevice(&pm->device));
    pm->nVerts = mesh.nVerts; pm->nIdx = mesh.nIdx; pm->nTris = mesh.numTriangles();
    pm->dev.numTriangles = pm->nTris;
    auto t0 = std::chrono::steady_clock::now();
    pm->dev.verts.alloc(mesh.nVerts); pm->dev.verts.upload(mesh.verts, mesh.nVerts);
    pm->dev.idx.alloc(mesh.nIdx); pm->dev.idx.upload(mesh.idx, mesh.nIdx);
    SDFB_CUDA(cudaDeviceSynchronize());
    pm->uploadMs = msSince(t0);
    // TriangleData, then the BVH, both on the device (bvh_device.cu). SDFB200_HOST_BVH: the host builder instead, on the host
    // threads while the device computes TriangleData.
    RawVec<BvhNode> bvh;
    double bvhMs = 0.0;
    const bool hostBvh = withBvh && hostBvhRequested();
    t0 = std::chrono::steady_clock::now();
    NvtxRange nvtx("sdfb200:mesh:triangle_data");
    static const bool hostTriangleData = std::getenv("SDFB200_HOST_TRIANGLE_DATA") != nullptr;   // A/B switch: the round-1 host path
    if (hostTriangleData) {
        pm->hostTris = computeTriangleData(mesh);
        pm->dev.tris.alloc(pm->nTris);
        pm->dev.tris.upload(pm->hostTris.data(), pm->nTris);
        SDFB_CUDA(cudaDeviceSynchronize());
        pm->triangleDataMs = msSince(t0);
        if (hostBvh) { t0 = std::chrono::steady_clock::now(); bvh = buildBvh(mesh); bvhMs = msSince(t0); }
    } else if (hostBvh) {
        // device work is enqueued by this thread; the BVH build forks its own host threads meanwhile
        std::exception_ptr bvhError;
        std::thread worker([&] {
            try {
                const auto tb = std::chrono::steady_clock::now();
                bvh = buildBvh(mesh);
                bvhMs = msSince(tb);
            } catch (...) { bvhError = std::current_exception(); }
        });
        try { triangleDataOnDevice(*pm, mesh); } catch (...) { worker.join(); throw; }
        pm->triangleDataMs = msSince(t0);
        worker.join();
        if (bvhError) std::rethrow_exception(bvhError);
    } else {
        triangleDataOnDevice(*pm, mesh);
        pm->triangleDataMs = msSince(t0);
    }
    nvtx.next("sdfb200:mesh:bvh");
    if (withBvh && !hostBvh) {
        t0 = std::chrono::steady_clock::now();
        gatherTriVerts(pm->dev);
        buildBvhOnDevice(pm->dev, &mesh);
        pm->hasBvh = true;
        bvhMs = msSince(t0);
    }
    pm->bvhMs = bvhMs;
    nvtx.next("sdfb200:mesh:exact_parts");
    t0 = std::chrono::steady_clock::now();
    if (hostBvh) uploadBvh(*pm, bvh);
    if (withExactParts) exactPartsOnDevice(*pm);
    pm->uploadMs += msSince(t0);
    return pm;
}

// BVH built elsewhere (one host build shared by the devices of a multi-device build): upload + pre-gathered triangle vertices
void attachBvh(PreparedMesh& pm, const RawVec<BvhNode>& bvh, double bvhMs) {
    const auto t0 = std::chrono::steady_clock::now();
    uploadBvh(pm, bvh);
    pm.bvhMs = bvhMs;
    pm.uploadMs += msSince(t0);
}

const TriVec& PreparedMesh::hostTriangleData() {
    std::lock_guard<std::mutex> lock(lazy);
    if (hostTris.size() != nTris) {
        int current = 0;
        SDFB_CUDA(cudaGetDevice(&current));
        SDFB_CUDA(cudaSetDevice(device));
        hostTris.resize(nTris);
        if (nTris) SDFB_CUDA(cudaMemcpy(hostTris.data(), dev.tris.p, size_t(nTris) * sizeof(TriData), cudaMemcpyDeviceToHost));
        SDFB_CUDA(cudaSetDevice(current));
    }
    return hostTris;
}

// ---- replication: a prepared mesh as one flat device blob (header + arrays, 256-byte aligned) ----------------------------
namespace {
struct BlobHeader {
    uint64_t magic, bytes;
    uint32_t nVerts, nIdx, nTris, numValid;
    uint32_t hasBvh, hasExactParts;
    int32_t rootLink, stackDepth;
    uint64_t bvhNodes;
    uint64_t off[8];   // verts, idx, tris, bvh, triVerts, frames, valid
};
constexpr uint64_t kBlobMagic = 0x5344464232303042ull;   // "SDFB200B"
uint64_t align256(uint64_t v) { return (v + 255) & ~uint64_t(255); }

struct BlobLayout { BlobHeader h; uint64_t size[7]; };
BlobLayout layoutOf(const PreparedMesh& pm) {
    BlobLayout L;
    std::memset(&L, 0, sizeof(L));
    L.h.magic = kBlobMagic;
    L.h.nVerts = pm.nVerts; L.h.nIdx = pm.nIdx; L.h.nTris = pm.nTris; L.h.numValid = pm.numValid;
    L.h.hasBvh = pm.hasBvh; L.h.hasExactParts = pm.hasExactParts;
    L.h.rootLink = pm.dev.rootLink; L.h.stackDepth = pm.dev.stackDepth;
    L.h.bvhNodes = pm.hasBvh ? pm.dev.bvh.n : 0;
    L.size[0] = uint64_t(pm.nVerts) * sizeof(f3);
    L.size[1] = uint64_t(pm.nIdx) * 4;
    L.size[2] = uint64_t(pm.nTris) * sizeof(TriData);
    L.size[3] = L.h.bvhNodes * sizeof(BvhNode);
    L.size[4] = pm.hasBvh ? uint64_t(pm.nTris) * 3 * sizeof(float4) : 0;
    L.size[5] = pm.hasExactParts ? uint64_t(pm.nTris) * 5 * sizeof(float4) : 0;
    L.size[6] = pm.hasExactParts ? (uint64_t(pm.numValid) + 8) * 4 : 0;
    uint64_t at = align256(sizeof(BlobHeader));
    for (int k = 0; k < 7; k++) { L.h.off[k] = at; at = align256(at + L.size[k]); }
    L.h.bytes = at;
    return L;
}
}  // namespace

uint64_t meshBlobBytes(const PreparedMesh& pm) { return layoutOf(pm).h.bytes; }

void meshBlobExport(const PreparedMesh& pm, void* dDst, uint64_t capacity, cudaStream_t st) {
    const BlobLayout L = layoutOf(pm);
    if (capacity < L.h.bytes) throw Error(SDFB200_ERR_INVALID, "mesh blob buffer too small");
    uint8_t* d = static_cast<uint8_t*>(dDst);
    SDFB_CUDA(cudaMemcpyAsync(d, &L.h, sizeof(BlobHeader), cudaMemcpyHostToDevice, st));
    const void* src[7] = {pm.dev.verts.p, pm.dev.idx.p, pm.dev.tris.p, pm.dev.bvh.p, pm.dev.triVerts.p, pm.frames.p, pm.valid.p};
    for (int k = 0; k < 7; k++)
        if (L.size[k]) SDFB_CUDA(cudaMemcpyAsync(d + L.h.off[k], src[k], L.size[k], cudaMemcpyDeviceToDevice, st));
    SDFB_CUDA(cudaStreamSynchronize(st));   // the header is read from this frame
}

std::shared_ptr<PreparedMesh> meshBlobImport(const void* dSrc, uint64_t bytes) {
    BlobHeader h;
    if (bytes < sizeof(BlobHeader)) throw Error(SDFB200_ERR_INVALID, "mesh blob too small");
    SDFB_CUDA(cudaMemcpy(&h, dSrc, sizeof(BlobHeader), cudaMemcpyDeviceToHost));
    if (h.magic != kBlobMagic || h.bytes > bytes) throw Error(SDFB200_ERR_INVALID, "not a prepared-mesh blob (or truncated)");
    std::shared_ptr<PreparedMesh> pm(new PreparedMesh());
    SDFB_CUDA(cudaGetDevice(&pm->device));
    pm->nVerts = h.nVerts; pm->nIdx = h.nIdx; pm->nTris = h.nTris; pm->numValid = h.numValid;
    pm->hasBvh = h.hasBvh != 0; pm->hasExactParts = h.hasExactParts != 0;
    pm->dev.numTriangles = h.nTris; pm->dev.rootLink = h.rootLink; pm->dev.stackDepth = h.stackDepth;
    pm->dev.verts.alloc(h.nVerts); pm->dev.idx.alloc(h.nIdx); pm->dev.tris.alloc(h.nTris);
    if (pm->hasBvh) { pm->dev.bvh.alloc(h.bvhNodes); pm->dev.triVerts.alloc(size_t(h.nTris) * 3); }
    if (pm->hasExactParts) { pm->frames.alloc(size_t(h.nTris) * 5); pm->valid.alloc(size_t(h.numValid) + 8); }
    const BlobLayout L = layoutOf(*pm);
    if (L.h.bytes != h.bytes) throw Error(SDFB200_ERR_INVALID, "prepared-mesh blob has an inconsistent layout");
    const uint8_t* s = static_cast<const uint8_t*>(dSrc);
    void* dst[7] = {pm->dev.verts.p, pm->dev.idx.p, pm->dev.tris.p, pm->dev.bvh.p, pm->dev.triVerts.p, pm->frames.p, pm->valid.p};
    for (int k = 0; k < 7; k++)
        if (L.size[k]) SDFB_CUDA(cudaMemcpyAsync(dst[k], s + h.off[k], L.size[k], cudaMemcpyDeviceToDevice));
    SDFB_CUDA(cudaDeviceSynchronize());
    return pm;
}

// Copy of `src` on the CURRENT device (peer copies over NVLink when the devices can reach each other, staged otherwise).
std::shared_ptr<PreparedMesh> cloneMeshToCurrentDevice(const PreparedMesh& src) {
    std::shared_ptr<PreparedMesh> pm(new PreparedMesh());
    SDFB_CUDA(cudaGetDevice(&pm->device));
    pm->nVerts = src.nVerts; pm->nIdx = src.nIdx; pm->nTris = src.nTris; pm->numValid = src.numValid;
    pm->hasBvh = src.hasBvh; pm->hasExactParts = src.hasExactParts;
    pm->dev.numTriangles = src.nTris; pm->dev.rootLink = src.dev.rootLink; pm->dev.stackDepth = src.dev.stackDepth;
    pm->triangleDataMs = 0; pm->bvhMs = 0;
    const auto t0 = std::chrono::steady_clock::now();
    auto copy = [&](auto& dst, const auto& from, size_t count) {
        using T = std::remove_pointer_t<decltype(dst.p)>;
        if (!count) return;
        dst.alloc(count);
        SDFB_CUDA(cudaMemcpyPeerAsync(dst.p, pm->device, from.p, src.device, count * sizeof(T)));
    };
    copy(pm->dev.verts, src.dev.verts, src.nVerts);
    copy(pm->dev.idx, src.dev.idx, src.nIdx);
    copy(pm->dev.tris, src.dev.tris, src.nTris);
    if (src.hasBvh) { copy(pm->dev.bvh, src.dev.bvh, src.dev.bvh.n); copy(pm->dev.triVerts, src.dev.triVerts, size_t(src.nTris) * 3); }
    if (src.hasExactParts) { copy(pm->frames, src.frames, size_t(src.nTris) * 5); copy(pm->valid, src.valid, size_t(src.numValid) + 8); }
    SDFB_CUDA(cudaDeviceSynchronize());
    pm->uploadMs = msSince(t0);
    return pm;
}

}  // namespace sdfb200
