// .bin persistence: byte-compatible with the reference's cereal PortableBinary archives
// (src/sdf/SdfFunction.cpp:9-79). Layout, little endian:
//   u8 1 | u32 format | 6 f32 box(min,max) | i32 startGridSize | ...
//   OCTREE       (OctreeSdf.h:225):      u32 maxDepth | f32 valueRange | f32 minBorderValue | u64 n | n x u32
//   EXACT_OCTREE (ExactOctreeSdf.h:141): u32 startDepth | u32 minTrianglesInLeafs | u32 maxTrianglesInLeafs |
//       u32 maxTrianglesEncodedInLeafs | u32 bitEncodingStartDepth | u32 bitsPerIndex | u32 maxDepth |
//       u64 n | n x (u32 childrenIndex, u32 trianglesArrayIndex) | u64 n | n x u32 sets | u64 n | n x u8 masks |
//       u64 T | T x 37 f32 TriangleData (TriangleUtils.h:53)
// Derived members (cell size, scratch sizes) are recomputed on load exactly as the reference does
// (OctreeSdf.h:233-234, ExactOctreeSdf.h:149-153).
#include <cstring>
#include <algorithm>
#include <fstream>

#include "sdf_internal.h"

namespace sdfb200 {

namespace {
template <class T> void put(std::ostream& os, const T& v) { os.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <class T> void get(std::istream& is, T& v) {
    is.read(reinterpret_cast<char*>(&v), sizeof(T));
    if (!is) throw Error(SDFB200_ERR_IO, "truncated .bin file");
}
template <class T> void putArray(std::ostream& os, const T* p, uint64_t n) {
    put(os, n);
    os.write(reinterpret_cast<const char*>(p), std::streamsize(n * sizeof(T)));
}
template <class V> void getArray(std::istream& is, V& v, uint64_t elemsPerCount = 1) {
    using T = typename std::remove_reference<decltype(v[0])>::type;
    uint64_t n = 0;
    get(is, n);
    if (n > (uint64_t(1) << 36)) throw Error(SDFB200_ERR_IO, "implausible array size in .bin file");
    v.resize(size_t(n * elemsPerCount));
    is.read(reinterpret_cast<char*>(v.data()), std::streamsize(v.size() * sizeof(T)));
    if (!is) throw Error(SDFB200_ERR_IO, "truncated .bin file");
}
}  // namespace

void saveBin(const sdfb200_sdf& s, const char* path) {
    std::ofstream os(path, std::ios::out | std::ios::binary);
    if (!os.is_open()) throw Error(SDFB200_ERR_IO, std::string("cannot open file ") + path);
    put(os, uint8_t(1));
    put(os, uint32_t(s.format));
    for (int i = 0; i < 3; i++) put(os, s.boxMin[i]);
    for (int i = 0; i < 3; i++) put(os, s.boxMax[i]);
    put(os, int32_t(s.startGridSize));
    if (s.format == SDFB200_FORMAT_OCTREE) {
        put(os, s.maxDepth); put(os, s.valueRange); put(os, s.minBorderValue);
        putArray(os, s.octree.data(), uint64_t(s.octree.size()));
    } else {
        put(os, s.startDepth); put(os, s.minTrisInLeafs); put(os, s.maxTrisInLeafs); put(os, s.maxTrisEncoded);
        put(os, s.bitEncodingStartDepth); put(os, s.bitsPerIndex); put(os, s.maxDepth);
        put(os, uint64_t(s.octree.size() / 2));
        os.write(reinterpret_cast<const char*>(s.octree.data()), std::streamsize(s.octree.size() * 4));
        putArray(os, s.sets.data(), uint64_t(s.sets.size()));
        putArray(os, s.masks.data(), uint64_t(s.masks.size()));
        { const TriVec& t = const_cast<sdfb200_sdf&>(s).hostTris(); putArray(os, t.data(), uint64_t(t.size())); }
    }
    if (!os) throw Error(SDFB200_ERR_IO, std::string("write failed on ") + path);
}

void loadBin(sdfb200_sdf& s, const char* path) {
    std::ifstream is(path, std::ios::binary);
    if (!is.is_open()) throw Error(SDFB200_ERR_IO, std::string("cannot open file ") + path);
    uint8_t littleEndian = 0;
    uint32_t format = 0;
    int32_t startGrid = 0;
    get(is, littleEndian);
    get(is, format);
    if (littleEndian != 1) throw Error(SDFB200_ERR_IO, "big-endian archives are not supported");
    if (format != SDFB200_FORMAT_OCTREE && format != SDFB200_FORMAT_EXACT_OCTREE)
        throw Error(SDFB200_ERR_IO, "unknown or unsupported SdfFormat in file (only OCTREE and EXACT_OCTREE)");
    s.format = int(format);
    for (int i = 0; i < 3; i++) get(is, s.boxMin[i]);
    for (int i = 0; i < 3; i++) get(is, s.boxMax[i]);
    get(is, startGrid);
    if (startGrid <= 0 || startGrid > 1024) throw Error(SDFB200_ERR_IO, "implausible start grid size");
    s.startGridSize = startGrid;
    if (format == SDFB200_FORMAT_OCTREE) {
        get(is, s.maxDepth); get(is, s.valueRange); get(is, s.minBorderValue);
        getArray(is, s.octree);
    } else {
        get(is, s.startDepth); get(is, s.minTrisInLeafs); get(is, s.maxTrisInLeafs); get(is, s.maxTrisEncoded);
        get(is, s.bitEncodingStartDepth); get(is, s.bitsPerIndex); get(is, s.maxDepth);
        getArray(is, s.octree, 2);
        getArray(is, s.sets);
        getArray(is, s.masks);
        getArray(is, s.tris);
    }
    s.cellSize = (s.boxMax[0] - s.boxMin[0]) / float(s.startGridSize);
    s.nOctree = s.octree.size(); s.nSets = s.sets.size(); s.nMasks = s.masks.size();
    s.hostMirror = true;
    validateStructure(s);
}

// A .bin file is untrusted input for the GPU kernels: every index the query kernels will follow is
// checked here once (children blocks and coefficient / triangle-set blocks inside their arrays), and
// the 16-byte alignment of OctreeSdf leaf blocks is recorded for the vectorised loads.
void validateStructure(sdfb200_sdf& s) {
    const uint64_t G3 = uint64_t(s.startGridSize) * s.startGridSize * s.startGridSize;
    if (s.format == SDFB200_FORMAT_OCTREE) {
        const uint64_t n = s.octree.size();
        if (n < G3) throw Error(SDFB200_ERR_IO, "octree array smaller than its start grid");
        s.leafBlocksAligned = true;
        std::vector<uint32_t> stack, depthOf;   // word index and its depth below the start grid
        stack.reserve(1024); depthOf.reserve(1024);
        uint64_t visited = 0;
        uint32_t deepest = 0;
        for (uint64_t r = 0; r < G3; r++) {
            stack.push_back(uint32_t(r)); depthOf.push_back(0u);
            while (!stack.empty()) {
                const uint32_t w = s.octree[stack.back()];
                const uint32_t dRel = depthOf.back();
                stack.pop_back(); depthOf.pop_back();
                deepest = std::max(deepest, dRel);
                const uint64_t at = w & kOctIndexMask;
                if (++visited > n) throw Error(SDFB200_ERR_IO, "octree array contains a cycle");
                if (w & kLeafBit) {
                    if (at + 64 > n) throw Error(SDFB200_ERR_IO, "leaf coefficient block out of range");
                    if (at % 4) s.leafBlocksAligned = false;
                } else {
                    if (at < G3 || at + 8 > n) throw Error(SDFB200_ERR_IO, "children block out of range");
                    for (uint32_t c = 0; c < 8; c++) { stack.push_back(uint32_t(at + c)); depthOf.push_back(dRel + 1); }
                }
            }
        }
        // The query kernels take the child choices from 16 path bits of the start-cell fraction: a file whose tree goes
        // deeper than that below its start grid would silently pick wrong children. The header's maxDepth is
        // untrusted and counts absolute depth, so the walk above measures it.
        if (deepest > 16) throw Error(SDFB200_ERR_IO, "octree deeper than 16 levels below its start grid is not supported by the query kernels");
        s.relativeDepth = deepest;
    } else {
        const uint64_t n = s.octree.size() / 2;
        if (n < G3) throw Error(SDFB200_ERR_IO, "node array smaller than its start grid");
        if (s.bitsPerIndex == 0 || s.bitsPerIndex > 31) throw Error(SDFB200_ERR_IO, "implausible bitsPerIndex");
        if (s.bitEncodingStartDepth < s.startDepth || s.bitEncodingStartDepth > s.maxDepth)
            throw Error(SDFB200_ERR_IO, "bitEncodingStartDepth outside [startDepth, maxDepth]");
        if (s.startGridSize != (1 << s.startDepth)) throw Error(SDFB200_ERR_IO, "start grid size does not match startDepth");
        for (uint64_t i = 0; i < n; i++) {
            const uint32_t w = s.octree[2 * i];
            if (!(w & kLeafBit) && (uint64_t(w) + 8 > n || w < G3)) throw Error(SDFB200_ERR_IO, "children block out of range");
        }
        // triangle indices inside the sets are checked lazily by the kernels against num_triangles
    }
}

void uploadStructure(sdfb200_sdf& s) {
    SDFB_CUDA(cudaGetDevice(&s.device));
    s.dOctree.alloc(s.octree.size());
    s.dOctree.upload(s.octree.data(), s.octree.size());
    if (s.format == SDFB200_FORMAT_EXACT_OCTREE) {
        // one zero pad word after the sets: the bit decoder always reads word w+1 (ExactOctreeSdf.cpp:73-77)
        s.dSets.alloc(s.sets.size() + 1);
        SDFB_CUDA(cudaMemsetAsync(s.dSets.p, 0, (s.sets.size() + 1) * 4));
        s.dSets.upload(s.sets.data(), s.sets.size());
        s.dMasks.alloc(s.masks.size() + 8);
        SDFB_CUDA(cudaMemsetAsync(s.dMasks.p, 0, s.masks.size() + 8));
        s.dMasks.upload(s.masks.data(), s.masks.size());
        s.numTris = uint32_t(s.tris.size());
        s.dTris.alloc(s.tris.size());
        s.dTris.upload(s.tris.data(), s.tris.size());
        prepareExactQuery(s);
    } else prepareOctreeQuery(s);
    SDFB_CUDA(cudaDeviceSynchronize());
}

}  // namespace sdfb200
