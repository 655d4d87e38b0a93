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
#include <cstring>

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

void shardExport(const sdfb200_sdf& s, uint32_t* dDst, uint64_t capacityWords) {
    requireFinished(s);
    if (capacityWords < shardPayloadWords(s)) throw Error(SDFB200_ERR_INVALID, "shard export buffer too small");
    SDFB_CUDA(cudaMemcpyAsync(dDst, s.shardScalars, 8, cudaMemcpyHostToDevice));
    uint64_t at = 2;
    for (uint32_t i = 0; i < s.plan.G3; i++) {
        const uint32_t r = s.plan.order[i];
        if (!s.plan.owned[r]) continue;
        SDFB_CUDA(cudaMemcpyAsync(dDst + at, s.dOctree.p + size_t(s.slotWords) * s.plan.rootSlot[r], 4 * s.slotWords, cudaMemcpyDeviceToDevice));
        at += s.slotWords;
        for (const ShardStream& st : s.streams) {
            const uint64_t bytes = uint64_t(st.rootSize[r]) * st.elemBytes;
            if (bytes) SDFB_CUDA(cudaMemcpyAsync(dDst + at, st.dBase + st.rootBase[r] * st.elemBytes, bytes, cudaMemcpyDeviceToDevice));
            at += segWords(st, r);
        }
    }
    SDFB_CUDA(cudaDeviceSynchronize());
}

void shardAssemble(sdfb200_sdf& s, const uint32_t* dGathered, const uint64_t* wordsPerRank, uint64_t strideWords, uint32_t world) {
    requireFinished(s);
    if (world != s.plan.world) throw Error(SDFB200_ERR_INVALID, "world size differs from the one the shard was built with");
    std::vector<uint64_t> at(world, 2);
    uint32_t sc0 = s.shardScalars[0], sc1 = s.shardScalars[1];
    for (uint32_t q = 0; q < world; q++) {
        uint32_t sc[2];
        SDFB_CUDA(cudaMemcpy(sc, dGathered + q * strideWords, 8, cudaMemcpyDeviceToHost));
        sc0 = std::max(sc0, sc[0]);
        sc1 = s.format == SDFB200_FORMAT_OCTREE ? std::min(sc1, sc[1]) : std::max(sc1, sc[1]);
    }
    for (uint32_t i = 0; i < s.plan.G3; i++) {
        const uint32_t r = s.plan.order[i], q = s.plan.ownerOf[r];
        const uint32_t* src = dGathered + q * strideWords;
        SDFB_CUDA(cudaMemcpyAsync(s.dOctree.p + size_t(s.slotWords) * s.plan.rootSlot[r], src + at[q], 4 * s.slotWords, cudaMemcpyDeviceToDevice));
        at[q] += s.slotWords;
        for (const ShardStream& st : s.streams) {
            const uint64_t bytes = uint64_t(st.rootSize[r]) * st.elemBytes;
            if (bytes) SDFB_CUDA(cudaMemcpyAsync(st.dBase + st.rootBase[r] * st.elemBytes, src + at[q], bytes, cudaMemcpyDeviceToDevice));
            at[q] += segWords(st, r);
        }
    }
    for (uint32_t q = 0; q < world; q++)
        if (at[q] != wordsPerRank[q]) throw Error(SDFB200_ERR_INVALID, "gathered payload sizes do not match the exchanged root sizes");
    SDFB_CUDA(cudaDeviceSynchronize());
    s.shardScalars[0] = sc0;
    s.shardScalars[1] = sc1;
    // host mirrors of the now complete structure
    s.octree.resize(s.dOctree.n);
    s.dOctree.download(s.octree.data(), s.octree.size());
    if (s.format == SDFB200_FORMAT_OCTREE) { finalizeOctreeScalars(s); prepareOctreeQuery(s); }
    else {
        s.maxTrisInLeafs = sc0;
        s.maxTrisEncoded = sc1;
        s.dSets.download(s.sets.data(), s.sets.size());
        s.dMasks.download(s.masks.data(), s.masks.size());
    }
    SDFB_CUDA(cudaDeviceSynchronize());
    if (s.format == SDFB200_FORMAT_EXACT_OCTREE) prepareExactQuery(s);
    s.isShard = false;
    s.build.reset();
}

}  // namespace sdfb200
