// Sharded construction, exchange side: pack the blocks of the own start-depth roots into one flat
// device buffer for the caller's all-gather, and rebuild the complete arrays from the gathered buffers.
//
// The reference merges per-start-voxel sub-octrees on the host after its OpenMP tasks finish
// (src/sdf/OctreeSdfDepthFirst.h:471-503, include/SdfLib/ExactOctreeSdfDepthFirst.h:576-622): concatenate the
// blocks in layout order and rebase their indices. Here every rank already emitted its blocks with FINAL indices
// (the per-root sizes were exchanged first), so assembly is pure data movement: block copies, no fix-up pass.
//
// Payload of a rank (uint32 words):
//   [scalar0][scalar1]   OCTREE: valueRange bits, order-preserving minBorderValue   EXACT: max leaf / max encoded list
//   for every own root, in layout order:
//     [start-slot record: 1 word (OCTREE) or 2 words (EXACT)]
//     for every stream (OCTREE: words; EXACT: node records, set words, mask bytes): the root's block, padded to a word
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sdf_internal.h"

namespace sdfb200 {

namespace {
inline uint64_t segWords(const ShardStream& s, uint32_t r) { return (uint64_t(s.rootSize[r]) * s.elemBytes + 3) / 4; }

void requireFinished(const sdfb200_sdf& s) {
    if (s.streams.empty()) throw Error(SDFB200_ERR_INVALID, "shard has no emitted blocks yet: call sdfb200_shard_finish first");
}
}  // namespace

uint64_t shardPayloadWords(const sdfb200_sdf& s) {
    requireFinished(s);
    uint64_t words = 2;
    for (uint32_t r = 0; r < s.plan.G3; r++) {
        if (!s.plan.owned[r]) continue;
        words += s.slotWords;
        for (const ShardStream& st : s.streams) words += segWords(st, r);
    }
    return words;
}

// One launch moves all blocks: segment k copies `bytes[k]` bytes (multiples of 4, 4-byte aligned both sides) from
// src[k] to dst[k]; a CTA per segment, strided 16-byte moves where the alignment allows.
struct Segment { const uint8_t* src; uint8_t* dst; uint64_t bytes; };

__global__ void __launch_bounds__(256) copySegmentsKernel(const Segment* seg) {
    const Segment g = seg[blockIdx.x];
    const uintptr_t both = reinterpret_cast<uintptr_t>(g.src) | reinterpret_cast<uintptr_t>(g.dst);
    if (both & 3u) {   // mask blocks start at arbitrary bytes
        for (uint64_t i = threadIdx.x; i < g.bytes; i += blockDim.x) g.dst[i] = g.src[i];
    } else if ((both & 15u) == 0) {
        const uint64_t n16 = g.bytes / 16;
        const uint4* s4 = reinterpret_cast<const uint4*>(g.src);
        uint4* d4 = reinterpret_cast<uint4*>(g.dst);
        for (uint64_t i = threadIdx.x; i < n16; i += blockDim.x) d4[i] = s4[i];
        for (uint64_t i = n16 * 16 + threadIdx.x; i < g.bytes; i += blockDim.x) g.dst[i] = g.src[i];
    } else {
        const uint64_t n4 = g.bytes / 4;
        const uint32_t* s1 = reinterpret_cast<const uint32_t*>(g.src);
        uint32_t* d1 = reinterpret_cast<uint32_t*>(g.dst);
        for (uint64_t i = threadIdx.x; i < n4; i += blockDim.x) d1[i] = s1[i];
        for (uint64_t i = n4 * 4 + threadIdx.x; i < g.bytes; i += blockDim.x) g.dst[i] = g.src[i];
    }
}

static void runSegments(const std::vector<Segment>& segs) {
    if (segs.empty()) return;
    // large blocks are cut so that one CTA moves at most 4 MiB
    std::vector<Segment> cut;
    constexpr uint64_t kPiece = uint64_t(4) << 20;
    for (const Segment& g : segs)
        for (uint64_t at = 0; at < g.bytes; at += kPiece) cut.push_back(Segment{g.src + at, g.dst + at, std::min(kPiece, g.bytes - at)});
    DevBuf<Segment> d(cut.size());
    d.upload(cut.data(), cut.size());
    copySegmentsKernel<<<uint32_t(cut.size()), 256>>>(d.p);
    SDFB_CUDA(cudaGetLastError());
    SDFB_CUDA(cudaDeviceSynchronize());   // `cut` and `d` go out of scope
}

void shardExport(const sdfb200_sdf& s, uint32_t* dDst, uint64_t capacityWords) {
    requireFinished(s);
    if (capacityWords < shardPayloadWords(s)) throw Error(SDFB200_ERR_INVALID, "shard export buffer too small");
    SDFB_CUDA(cudaMemcpyAsync(dDst, s.shardScalars, 8, cudaMemcpyHostToDevice));
    std::vector<Segment> segs;
    uint64_t at = 2;
    for (uint32_t i = 0; i < s.plan.G3; i++) {
        const uint32_t r = s.plan.order[i];
        if (!s.plan.owned[r]) continue;
        segs.push_back(Segment{reinterpret_cast<const uint8_t*>(s.dOctree.p + size_t(s.slotWords) * s.plan.rootSlot[r]),
                               reinterpret_cast<uint8_t*>(dDst + at), 4ull * s.slotWords});
        at += s.slotWords;
        for (const ShardStream& st : s.streams) {
            const uint64_t bytes = uint64_t(st.rootSize[r]) * st.elemBytes;
            if (bytes) segs.push_back(Segment{st.dBase + st.rootBase[r] * st.elemBytes, reinterpret_cast<uint8_t*>(dDst + at), bytes});
            at += segWords(st, r);
        }
    }
    runSegments(segs);
}

void shardAssemble(sdfb200_sdf& s, const uint32_t* dGathered, const uint64_t* wordsPerRank, uint64_t strideWords, uint32_t world) {
    requireFinished(s);
    if (world != s.plan.world) throw Error(SDFB200_ERR_INVALID, "world size differs from the one the shard was built with");
    // validate the exchanged sizes against the plan BEFORE anything is copied: the handle stays untouched on failure
    std::vector<uint64_t> expect(world, 2);
    for (uint32_t r = 0; r < s.plan.G3; r++) {
        uint64_t w = s.slotWords;
        for (const ShardStream& st : s.streams) w += segWords(st, r);
        expect[s.plan.ownerOf[r]] += w;
    }
    for (uint32_t q = 0; q < world; q++) {
        if (expect[q] != wordsPerRank[q]) throw Error(SDFB200_ERR_INVALID, "gathered payload sizes do not match the exchanged root sizes");
        if (wordsPerRank[q] > strideWords) throw Error(SDFB200_ERR_INVALID, "a rank's payload is larger than the gather stride");
    }
    std::vector<uint64_t> at(world, 2);
    uint32_t sc0 = s.shardScalars[0], sc1 = s.shardScalars[1];
    std::vector<uint32_t> scalars(size_t(world) * 2);
    for (uint32_t q = 0; q < world; q++) SDFB_CUDA(cudaMemcpyAsync(&scalars[2 * size_t(q)], dGathered + q * strideWords, 8, cudaMemcpyDeviceToHost));
    SDFB_CUDA(cudaDeviceSynchronize());
    for (uint32_t q = 0; q < world; q++) {
        sc0 = std::max(sc0, scalars[2 * size_t(q)]);
        sc1 = s.format == SDFB200_FORMAT_OCTREE ? std::min(sc1, scalars[2 * size_t(q) + 1]) : std::max(sc1, scalars[2 * size_t(q) + 1]);
    }
    std::vector<Segment> segs;
    for (uint32_t i = 0; i < s.plan.G3; i++) {
        const uint32_t r = s.plan.order[i], q = s.plan.ownerOf[r];
        const uint32_t* src = dGathered + q * strideWords;
        if (q != s.plan.rank)   // own blocks are already in place
            segs.push_back(Segment{reinterpret_cast<const uint8_t*>(src + at[q]),
                                   reinterpret_cast<uint8_t*>(s.dOctree.p + size_t(s.slotWords) * s.plan.rootSlot[r]), 4ull * s.slotWords});
        at[q] += s.slotWords;
        for (const ShardStream& st : s.streams) {
            const uint64_t bytes = uint64_t(st.rootSize[r]) * st.elemBytes;
            if (bytes && q != s.plan.rank) segs.push_back(Segment{reinterpret_cast<const uint8_t*>(src + at[q]), st.dBase + st.rootBase[r] * st.elemBytes, bytes});
            at[q] += segWords(st, r);
        }
    }
    runSegments(segs);
    SDFB_CUDA(cudaDeviceSynchronize());
    s.shardScalars[0] = sc0;
    s.shardScalars[1] = sc1;
    // the structure is complete on this device; its host mirrors are fetched when a getter or the .bin writer asks
    s.hostMirror = false;
    static const bool timing = std::getenv("SDFB200_TIMING") != nullptr;
    const auto tq = std::chrono::steady_clock::now();
    if (s.format == SDFB200_FORMAT_OCTREE) { finalizeOctreeScalars(s); prepareOctreeQuery(s); }
    else {
        s.maxTrisInLeafs = sc0;
        s.maxTrisEncoded = sc1;
        prepareExactQuery(s);
    }
    if (timing) { SDFB_CUDA(cudaDeviceSynchronize()); std::fprintf(stderr, "[sdfb200] assemble: query-side structures %8.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tq).count()); }
    s.isShard = false;
    s.build.reset();
}

void ensureHostMirror(sdfb200_sdf& s) {
    if (s.hostMirror || !s.dOctree.p) return;
    int current = 0;
    SDFB_CUDA(cudaGetDevice(&current));
    SDFB_CUDA(cudaSetDevice(s.device));
    s.octree.resize(s.nOctree);
    s.dOctree.download(s.octree.data(), s.nOctree);
    if (s.format == SDFB200_FORMAT_EXACT_OCTREE) {
        s.sets.resize(s.nSets);
        s.masks.resize(s.nMasks);
        s.dSets.download(s.sets.data(), s.nSets);
        s.dMasks.download(s.masks.data(), s.nMasks);
    }
    SDFB_CUDA(cudaDeviceSynchronize());
    SDFB_CUDA(cudaSetDevice(current));
    s.hostMirror = true;
}

}  // namespace sdfb200
