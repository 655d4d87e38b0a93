// Kernels of the device-side build of the float64 sphere BVH (bvh_device.cu is the translation unit and holds the design notes).
// Reference: tmd::TriangleMeshDistance::_build_tree, libs/InteractiveComputerGraphics/.../TriangleMeshDistance.h:421-490.
// Kept in a header so that tests/cpp/simt_bvh_main.cpp can run this very source on the CPU under a CTA emulation (threads of
// a CTA = host threads, __syncthreads / warp collectives = barriers) and compare every node with the host builder of
// mesh_host.cpp bit for bit, without a GPU. Include inside namespace sdfb200, inside an anonymous namespace.
//
// What has to be reproduced exactly, because the traversal order of this tree decides which of two equidistant triangles
// the OctreeSdf builders pick:
//   * the split axis: widest extent of the node's vertices (float64 differences of float32 coordinates, first maximum);
//   * the ORDER std::sort leaves the node's triangles in. The key is the first vertex along the split axis, so all triangles
//     that start at the same mesh vertex tie, and std::sort is not stable: libstdc++'s introsort (median-of-three Hoare
//     partitions to pieces of <= 16, heap sort when the depth limit 2 lg n runs out, one final insertion pass) is restated
//     here operation for operation (bvhPartitionStep, bvhHeapSort, bvhFinishPieces);
//   * the sphere centre: a SEQUENTIAL float64 sum over the node's vertices in the order the parent's sort left them
//     (bvhCentreKernel), and the radius: a maximum (any order).
#pragma once

#ifndef BVH_SMALL_MAX   // the CPU emulation overrides these three so that small meshes reach every path
#define BVH_SMALL_MAX 2048
#define BVH_BIG_THREADS 1024
#define BVH_SMALL_THREADS 128
#endif
constexpr int kBvhSmallMax = BVH_SMALL_MAX;          // sort tasks up to this many elements are finished inside one CTA's shared memory
constexpr int kBvhInsertion = 16;                    // libstdc++'s _S_threshold: pieces of at most 16 elements are left to the final insertion pass
constexpr int kBvhBigThreads = BVH_BIG_THREADS;      // CTA of the global-memory partition step
constexpr int kBvhSmallThreads = BVH_SMALL_THREADS;  // CTA of the shared-memory sort
constexpr int kBvhWarpCentre = 64;     // segments of at least this many triangles get a warp for their centre sum

// ---- the static shape of the tree ---------------------------------------------------------------------------------------------
// mid = int(0.5 * (begin + end)) whatever the data, so ranges and node ids (pre-order: a subtree over m triangles owns 2m - 1
// ids) follow from n alone; only the order inside the ranges and the spheres depend on the mesh.
struct BvhSeg {
    int32_t b, e;       // triangle range [b, e) of the node
    int32_t node;       // node id
    int32_t parent;     // parent's node id (-1 for the root)
    int32_t side;       // 0 = left child, 1 = right child
    bool valid;         // false: an ancestor above this level is already a leaf
};

__host__ __device__ __forceinline__ void bvhDescend(BvhSeg& s, bool right) {
    const int32_t mid = int32_t((uint32_t(s.b) + uint32_t(s.e)) >> 1);
    s.parent = s.node;
    if (right) { s.node = s.node + 2 * (mid - s.b); s.b = mid; s.side = 1; }
    else { s.node = s.node + 1; s.e = mid; s.side = 0; }
}
// the level-`level` node that holds position p of the triangle order; slot = its index among the 2^level nodes of the level
__host__ __device__ __forceinline__ BvhSeg bvhSegOfPosition(int32_t n, int level, int32_t p, uint32_t& slot) {
    BvhSeg s{0, n, 0, -1, 0, true};
    slot = 0;
    for (int k = 0; k < level; k++) {
        if (s.e - s.b <= 1) { s.valid = false; return s; }
        const int32_t mid = int32_t((uint32_t(s.b) + uint32_t(s.e)) >> 1);
        const bool right = p >= mid;
        bvhDescend(s, right);
        slot = slot * 2u + (right ? 1u : 0u);
    }
    return s;
}
__host__ __device__ __forceinline__ BvhSeg bvhSegOfSlot(int32_t n, int level, uint32_t slot) {
    BvhSeg s{0, n, 0, -1, 0, true};
    for (int bit = level - 1; bit >= 0; bit--) {
        if (s.e - s.b <= 1) { s.valid = false; return s; }
        bvhDescend(s, ((slot >> bit) & 1u) != 0u);
    }
    return s;
}

// floats as integers that order the same way (atomicMin / atomicMax on coordinates)
__host__ __device__ __forceinline__ int32_t bvhOrderedInt(float f) {
    int32_t i;
#ifdef __CUDA_ARCH__
    i = __float_as_int(f);
#else
    memcpy(&i, &f, 4);
#endif
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__host__ __device__ __forceinline__ float bvhOrderedFloat(int32_t i) {
    i = i >= 0 ? i : i ^ 0x7FFFFFFF;
    float f;
#ifdef __CUDA_ARCH__
    f = __int_as_float(i);
#else
    memcpy(&f, &i, 4);
#endif
    return f;
}

__device__ __forceinline__ float bvhComponent(float4 v, int dim) { return dim == 0 ? v.x : (dim == 1 ? v.y : v.z); }

// ---- per level: bounding box of every node (decides the split axis) ---------------------------------------------------------
__global__ void __launch_bounds__(256)
bvhBoundsKernel(int32_t n, int level, const int32_t* __restrict__ ids, const float4* __restrict__ triVerts, int32_t* boxMin, int32_t* boxMax) {
    constexpr unsigned kFull = 0xffffffffu;
    const int64_t p64 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t slot = 0;
    bool active = p64 < int64_t(n);
    if (active) {
        const BvhSeg s = bvhSegOfPosition(n, level, int32_t(p64), slot);
        active = s.valid && s.e - s.b > 1;
    }
    int32_t lo[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, hi[3] = {int32_t(0x80000000), int32_t(0x80000000), int32_t(0x80000000)};
    if (active) {
        const int32_t id = ids[p64];
        for (int k = 0; k < 3; k++) {
            const float4 v = triVerts[size_t(id) * 3 + k];
            const int32_t c[3] = {bvhOrderedInt(v.x), bvhOrderedInt(v.y), bvhOrderedInt(v.z)};
            for (int a = 0; a < 3; a++) { lo[a] = c[a] < lo[a] ? c[a] : lo[a]; hi[a] = c[a] > hi[a] ? c[a] : hi[a]; }
        }
    }
    const unsigned act = __ballot_sync(kFull, active);
    if (act == 0) return;                                              // warp-uniform
    const int leader = __ffs(int(act)) - 1;
    const uint32_t slot0 = __shfl_sync(kFull, slot, leader);
    const bool uniform = __ballot_sync(kFull, active && slot != slot0) == 0;   // every active lane sits in the leader's node
    if (uniform) {
        for (int m = 16; m >= 1; m >>= 1)
            for (int a = 0; a < 3; a++) {
                const int32_t l2 = __shfl_xor_sync(kFull, lo[a], m), h2 = __shfl_xor_sync(kFull, hi[a], m);
                lo[a] = l2 < lo[a] ? l2 : lo[a]; hi[a] = h2 > hi[a] ? h2 : hi[a];
            }
        if (int(lane) == leader)
            for (int a = 0; a < 3; a++) { atomicMin(boxMin + size_t(slot0) * 3 + a, lo[a]); atomicMax(boxMax + size_t(slot0) * 3 + a, hi[a]); }
    } else if (active) {
        for (int a = 0; a < 3; a++) { atomicMin(boxMin + size_t(slot) * 3 + a, lo[a]); atomicMax(boxMax + size_t(slot) * 3 + a, hi[a]); }
    }
}

// ---- sort tasks ---------------------------------------------------------------------------------------------------------------
struct BvhSortTask { int32_t first, last, depth; };   // a range std::sort's introsort loop still has to work on, and its remaining depth

// per level: split axis -> keys; the node's first thread queues its std::sort call (ranges of <= 16 go to bvhTinySortKernel)
__global__ void __launch_bounds__(256)
bvhKeysKernel(int32_t n, int level, const int32_t* __restrict__ ids, const float4* __restrict__ triVerts, const int32_t* __restrict__ boxMin,
              const int32_t* __restrict__ boxMax, float* __restrict__ keys, BvhSortTask* bigTasks, uint32_t* bigCount, BvhSortTask* smallTasks,
              uint32_t* smallCount) {
    const int64_t p64 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p64 >= int64_t(n)) return;
    const int32_t p = int32_t(p64);
    uint32_t slot;
    const BvhSeg s = bvhSegOfPosition(n, level, p, slot);
    const int32_t m = s.e - s.b;
    if (!s.valid || m <= 1) return;
    // split_dim = first maximum of (top - bottom), float64 differences (TriangleMeshDistance.h:455-456)
    double ext[3];
    for (int a = 0; a < 3; a++) ext[a] = double(bvhOrderedFloat(boxMax[size_t(slot) * 3 + a])) - double(bvhOrderedFloat(boxMin[size_t(slot) * 3 + a]));
    int dim = 0;
    for (int a = 1; a < 3; a++)
        if (ext[a] > ext[dim]) dim = a;
    keys[p] = bvhComponent(triVerts[size_t(ids[p]) * 3], dim);
    if (p == s.b && m > kBvhInsertion) {
        const BvhSortTask t{s.b, s.e, 2 * (31 - __clz(m))};                 // std::sort: depth limit 2 * floor(lg n)
        if (m > kBvhSmallMax) bigTasks[atomicAdd(bigCount, 1u)] = t;
        else smallTasks[atomicAdd(smallCount, 1u)] = t;
    }
}

// std::__insertion_sort on [first, last): a stable sort (guarded form; the library's two variants give the same permutation)
template <class KeyPtr, class IdPtr> __device__ __forceinline__ void bvhInsertionSort(KeyPtr keys, IdPtr ids, int32_t first, int32_t last) {
    for (int32_t i = first + 1; i < last; i++) {
        const float k = keys[i];
        const int32_t id = ids[i];
        int32_t j = i;
        while (j > first && k < keys[j - 1]) { keys[j] = keys[j - 1]; ids[j] = ids[j - 1]; j--; }
        keys[j] = k; ids[j] = id;
    }
}

// levels whose nodes hold at most 16 triangles: std::sort is the insertion pass alone, one thread per node
__global__ void __launch_bounds__(256) bvhTinySortKernel(int32_t n, int level, float* keys, int32_t* ids) {
    const int64_t p64 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p64 >= int64_t(n)) return;
    uint32_t slot;
    const BvhSeg s = bvhSegOfPosition(n, level, int32_t(p64), slot);
    const int32_t m = s.e - s.b;
    if (!s.valid || int32_t(p64) != s.b || m <= 1 || m > kBvhInsertion) return;
    bvhInsertionSort(keys, ids, s.b, s.e);
}

// std::__partial_sort(first, last, last) = make_heap + sort_heap (bits/stl_heap.h), on records (key, id)
template <class KeyPtr, class IdPtr>
__device__ void bvhAdjustHeap(KeyPtr keys, IdPtr ids, int32_t first, int32_t hole, int32_t len, float vk, int32_t vid) {
    const int32_t top = hole;
    int32_t child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (keys[first + child] < keys[first + child - 1]) child--;
        keys[first + hole] = keys[first + child]; ids[first + hole] = ids[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        keys[first + hole] = keys[first + child - 1]; ids[first + hole] = ids[first + child - 1];
        hole = child - 1;
    }
    int32_t parent = (hole - 1) / 2;                                       // __push_heap
    while (hole > top && keys[first + parent] < vk) {
        keys[first + hole] = keys[first + parent]; ids[first + hole] = ids[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    keys[first + hole] = vk; ids[first + hole] = vid;
}
template <class KeyPtr, class IdPtr> __device__ void bvhHeapSort(KeyPtr keys, IdPtr ids, int32_t first, int32_t last) {
    const int32_t len = last - first;
    if (len < 2) return;
    for (int32_t parent = (len - 2) / 2;; parent--) {                      // __make_heap
        bvhAdjustHeap(keys, ids, first, parent, len, float(keys[first + parent]), int32_t(ids[first + parent]));
        if (parent == 0) break;
    }
    for (int32_t end = last; end - first > 1;) {                           // __sort_heap: __pop_heap(first, end - 1, end - 1)
        --end;
        const float vk = keys[end];
        const int32_t vid = ids[end];
        keys[end] = keys[first]; ids[end] = ids[first];
        bvhAdjustHeap(keys, ids, first, 0, end - first, vk, vid);
    }
}

// One std::__unguarded_partition_pivot by a whole CTA. The serial loop swaps the k-th element from the left that is not
// below the pivot with the k-th element from the right that is not above it, while the former lies left of the latter;
// swapped elements are never looked at again, so the pairs are known from two ordered compactions:
//   L_k = k-th position (ascending) with !(a < pivot),   R_k = k-th position (descending) with !(pivot < a),
//   K = #{k : L_k < R_k} swaps,   cut = min(L_K, R_{K-1})   (tests/cpp/simt_bvh_main.cpp checks this against the library).
// lpos / rpos: scratch indexed like the range itself (Pos = uint32_t in global memory, uint16_t in shared memory).
constexpr int kBvhItems = 4;   // elements per thread and pass of the compaction loop
struct BvhPartitionShared {
    uint32_t count[2][kBvhItems * 32];     // per (item, warp): candidates on the left in the low half, on the right in the high half
    uint32_t before[2][kBvhItems * 32 + 1];
    float pivot;
    uint32_t swaps;
};
__device__ __forceinline__ int32_t bvhMedianOfThree(float ka, float kb, float kc, int32_t a, int32_t b, int32_t c) {   // __move_median_to_first
    if (ka < kb) return kb < kc ? b : (ka < kc ? c : a);
    return ka < kc ? a : (kb < kc ? c : b);
}
template <int kThreads, class Pos, class KeyPtr, class IdPtr>
__device__ int32_t bvhPartitionStep(KeyPtr keys, IdPtr ids, int32_t first, int32_t last, Pos* lpos, Pos* rpos, BvhPartitionShared& sh) {
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int kWarps = kThreads / 32, kEntries = kBvhItems * kWarps, kPerLane = (kEntries + 31) / 32;
    const int tid = int(threadIdx.x), lane = tid & 31, warp = tid >> 5;
    __syncthreads();                                                       // the range as the previous step left it
    if (tid == 0) {                                                        // __move_median_to_first(first, first + 1, mid, last - 1)
        const int32_t a = first + 1, b = first + (last - first) / 2, c = last - 1;
        const int32_t med = bvhMedianOfThree(keys[a], keys[b], keys[c], a, b, c);
        const float kf = keys[first], km = keys[med];
        const int32_t idf = ids[first], idm = ids[med];
        keys[first] = km; keys[med] = kf; ids[first] = idm; ids[med] = idf;
        sh.pivot = km;
        sh.swaps = 0;
    }
    __syncthreads();
    const float pivot = sh.pivot;
    const int32_t base0 = first + 1;
    uint32_t nL = 0, nR = 0;
    int buf = 0;
    for (int32_t base = base0; base < last; base += kThreads * kBvhItems, buf ^= 1) {
        bool fl[kBvhItems], fr[kBvhItems];
        unsigned bl[kBvhItems], br[kBvhItems];
        float k[kBvhItems];
#pragma unroll
        for (int j = 0; j < kBvhItems; j++) {                              // the loads of a pass are independent: all in flight together
            const int32_t i = base + j * kThreads + tid;
            k[j] = i < last ? float(keys[i]) : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < kBvhItems; j++) {
            const bool in = base + j * kThreads + tid < last;
            fl[j] = in && !(k[j] < pivot); fr[j] = in && !(pivot < k[j]);
            bl[j] = __ballot_sync(kFull, fl[j]); br[j] = __ballot_sync(kFull, fr[j]);
            if (lane == 0) sh.count[buf][j * kWarps + warp] = uint32_t(__popc(bl[j])) | (uint32_t(__popc(br[j])) << 16);
        }
        __syncthreads();
        if (warp == 0) {                                                   // exclusive prefix over the (item, warp) counts: both halves at once
            uint32_t c[kPerLane], sum = 0;
#pragma unroll
            for (int q = 0; q < kPerLane; q++) {
                const int e = lane * kPerLane + q;
                c[q] = e < kEntries ? sh.count[buf][e] : 0u;
                sum += c[q];
            }
            uint32_t incl = sum;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_sync(kFull, incl, lane >= o ? lane - o : lane);
                if (lane >= o) incl += t;
            }
            uint32_t run = incl - sum;
#pragma unroll
            for (int q = 0; q < kPerLane; q++) {
                const int e = lane * kPerLane + q;
                if (e < kEntries) sh.before[buf][e] = run;
                run += c[q];
            }
            if (lane == 31) sh.before[buf][kEntries] = incl;
        }
        __syncthreads();
        const unsigned below = (1u << lane) - 1u;
#pragma unroll
        for (int j = 0; j < kBvhItems; j++) {
            const uint32_t bef = sh.before[buf][j * kWarps + warp];
            const int32_t i = base + j * kThreads + tid;
            if (fl[j]) lpos[base0 + nL + (bef & 0xFFFFu) + uint32_t(__popc(bl[j] & below))] = Pos(i);
            if (fr[j]) rpos[base0 + nR + (bef >> 16) + uint32_t(__popc(br[j] & below))] = Pos(i);
        }
        const uint32_t total = sh.before[buf][kEntries];
        nL += total & 0xFFFFu; nR += total >> 16;
    }
    __syncthreads();                                                       // both lists complete
    const uint32_t nMin = nL < nR ? nL : nR;
    uint32_t mine = 0;
    for (uint32_t k0 = uint32_t(tid); k0 < nMin; k0 += kThreads * kBvhItems) {
        int32_t pl[kBvhItems], pr[kBvhItems];
        bool ok[kBvhItems];
#pragma unroll
        for (int u = 0; u < kBvhItems; u++) {
            const uint32_t k = k0 + uint32_t(u) * kThreads;
            ok[u] = k < nMin;
            pl[u] = ok[u] ? int32_t(lpos[base0 + k]) : 0;
            pr[u] = ok[u] ? int32_t(rpos[base0 + nR - 1 - k]) : 0;
        }
        bool done = false;
#pragma unroll
        for (int u = 0; u < kBvhItems; u++) {
            ok[u] = ok[u] && pl[u] < pr[u];                                // monotone in k: once a pair does not swap, no later pair does
            if (!ok[u]) done = true;
        }
        float kl[kBvhItems], kr[kBvhItems];
        int32_t il[kBvhItems], ir[kBvhItems];
#pragma unroll
        for (int u = 0; u < kBvhItems; u++)
            if (ok[u]) { kl[u] = keys[pl[u]]; kr[u] = keys[pr[u]]; il[u] = ids[pl[u]]; ir[u] = ids[pr[u]]; }
#pragma unroll
        for (int u = 0; u < kBvhItems; u++)
            if (ok[u]) { keys[pl[u]] = kr[u]; keys[pr[u]] = kl[u]; ids[pl[u]] = ir[u]; ids[pr[u]] = il[u]; mine++; }
        if (done) break;
    }
    if (mine) atomicAdd(&sh.swaps, mine);
    __syncthreads();
    const uint32_t K = sh.swaps;
    int32_t cut = 0x7FFFFFFF;
    if (K < nL) cut = int32_t(lpos[base0 + K]);
    if (K > 0) { const int32_t r = int32_t(rpos[base0 + nR - K]); cut = r < cut ? r : cut; }
    return cut;
}

// one round of the introsort loop for the ranges that do not fit a CTA's shared memory: partition, queue the two parts
__global__ void __launch_bounds__(kBvhBigThreads)
bvhBigPartitionKernel(float* keys, int32_t* ids, uint32_t* lpos, uint32_t* rpos, const BvhSortTask* __restrict__ in, const uint32_t* __restrict__ inCount,
                      BvhSortTask* out, uint32_t* outCount, BvhSortTask* smallTasks, uint32_t* smallCount) {
    __shared__ BvhPartitionShared sh;
    const uint32_t count = *inCount;
    for (uint32_t t = blockIdx.x; t < count; t += gridDim.x) {
        const BvhSortTask task = in[t];
        if (task.depth == 0) {                                             // depth limit spent: std::__partial_sort, then nothing is left to do
            if (threadIdx.x == 0) bvhHeapSort(keys, ids, task.first, task.last);
            continue;
        }
        const int32_t cut = bvhPartitionStep<kBvhBigThreads, uint32_t>(keys, ids, task.first, task.last, lpos, rpos, sh);
        if (threadIdx.x == 0) {
            const BvhSortTask part[2] = {{task.first, cut, task.depth - 1}, {cut, task.last, task.depth - 1}};
            for (int c = 0; c < 2; c++) {
                const int32_t len = part[c].last - part[c].first;
                if (len > kBvhSmallMax) out[atomicAdd(outCount, 1u)] = part[c];
                else if (len > 1) smallTasks[atomicAdd(smallCount, 1u)] = part[c];
            }
        }
    }
}

// std::__unguarded_partition_pivot by ONE thread, literally (ranges of <= kBvhSequential elements in shared memory). The scans
// are bounded by the range as well: with the median in front they never get there, but NaN keys must not walk out of the array.
template <class KeyPtr, class IdPtr> __device__ int32_t bvhSequentialPartition(KeyPtr keys, IdPtr ids, int32_t first, int32_t last) {
    {
        const int32_t a = first + 1, b = first + (last - first) / 2, c = last - 1;
        const int32_t med = bvhMedianOfThree(keys[a], keys[b], keys[c], a, b, c);
        const float kf = keys[first]; const int32_t idf = ids[first];
        keys[first] = keys[med]; ids[first] = ids[med]; keys[med] = kf; ids[med] = idf;
    }
    const float pivot = keys[first];
    int32_t lo = first + 1, hi = last;
    for (;;) {
        while (lo < last && keys[lo] < pivot) lo++;
        hi--;
        while (hi > first && pivot < keys[hi]) hi--;
        if (!(lo < hi)) return lo;
        const float k = keys[lo]; const int32_t id = ids[lo];
        keys[lo] = keys[hi]; ids[lo] = ids[hi]; keys[hi] = k; ids[hi] = id;
        lo++;
    }
}

constexpr int kBvhSequential = 64;   // ranges of at most this many elements are finished by one thread each

// the rest of std::sort for a range that fits shared memory: introsort loop, then the final insertion pass. The CTA partitions
// together (bvhPartitionStep, explicit stack) while ranges are longer than kBvhSequential; the shorter ranges are then taken
// one per thread and partitioned literally. After the loop the range is a sequence of pieces (<= 16 elements each, or
// heap-sorted), every piece <= the next one, so the insertion pass — a stable sort — is a stable sort of each piece: every
// element counts the piece members that precede it.
constexpr int kBvhShortMax = kBvhSmallMax / (kBvhInsertion + 1) + 2;       // disjoint ranges of more than 16 elements

__global__ void __launch_bounds__(kBvhSmallThreads)
bvhSmallSortKernel(float* keys, int32_t* ids, const BvhSortTask* __restrict__ tasks, const uint32_t* __restrict__ taskCount) {
    __shared__ float sKey[kBvhSmallMax];
    __shared__ int32_t sId[kBvhSmallMax];
    __shared__ uint16_t sL[kBvhSmallMax], sR[kBvhSmallMax];                // scratch of the CTA-wide partition steps
    __shared__ uint16_t sShortFirst[kBvhShortMax], sShortLast[kBvhShortMax];
    __shared__ uint8_t sShortDepth[kBvhShortMax];
    __shared__ uint32_t sStart[kBvhSmallMax / 32 + 1];                     // bit p set: a piece starts at p
    __shared__ BvhSortTask sStack[64];
    __shared__ BvhPartitionShared sh;
    __shared__ int sTop, sShort;
    const int tid = int(threadIdx.x);
    const uint32_t count = *taskCount;
    for (uint32_t t = blockIdx.x; t < count; t += gridDim.x) {
        const BvhSortTask task = tasks[t];
        const int32_t m = task.last - task.first;
        __syncthreads();                                                   // previous task's shared arrays are free
        for (int32_t i = tid; i < m; i += kBvhSmallThreads) { sKey[i] = keys[task.first + i]; sId[i] = ids[task.first + i]; }
        for (int32_t i = tid; i < kBvhSmallMax / 32 + 1; i += kBvhSmallThreads) sStart[i] = i == 0 ? 1u : 0u;
        if (tid == 0) { sStack[0] = BvhSortTask{0, m, task.depth}; sTop = 1; sShort = 0; }
        __syncthreads();
        while (sTop > 0) {                                                 // CTA-uniform: sTop changes only between barriers
            const BvhSortTask cur = sStack[sTop - 1];
            __syncthreads();
            if (tid == 0) sTop--;
            int32_t first = cur.first, last = cur.last, depth = cur.depth;
            while (last - first > kBvhSequential) {
                if (depth == 0) {                                          // depth limit spent: std::__partial_sort, nothing left of this range
                    __syncthreads();
                    if (tid == 0) bvhHeapSort(sKey, sId, first, last);
                    for (int32_t i = first + tid; i < last; i += kBvhSmallThreads) atomicOr(&sStart[i >> 5], 1u << (i & 31));   // sorted: every element its own piece
                    first = last;
                    break;
                }
                depth--;
                const int32_t cut = bvhPartitionStep<kBvhSmallThreads, uint16_t>(sKey, sId, first, last, sL, sR, sh);
                if (tid == 0) {
                    atomicOr(&sStart[cut >> 5], 1u << (cut & 31));
                    if (last - cut > kBvhSequential) sStack[sTop++] = BvhSortTask{cut, last, depth};   // __introsort_loop(cut, last, depth): later
                    else if (last - cut > kBvhInsertion) { sShortFirst[sShort] = uint16_t(cut); sShortLast[sShort] = uint16_t(last); sShortDepth[sShort++] = uint8_t(depth); }
                }
                last = cut;                                                // ... and loop on [first, cut)
            }
            if (tid == 0 && last - first > kBvhInsertion) { sShortFirst[sShort] = uint16_t(first); sShortLast[sShort] = uint16_t(last); sShortDepth[sShort++] = uint8_t(depth); }
            __syncthreads();
        }
        // the short ranges: the same loop, one thread per range, partitions done literally
        for (int q = tid; q < sShort; q += kBvhSmallThreads) {
            int32_t stackFirst[4], stackLast[4], stackDepth[4];            // pending right parts: longer than 16, disjoint, inside 64 elements
            int top = 0;
            int32_t first = sShortFirst[q], last = sShortLast[q], depth = sShortDepth[q];
            for (;;) {
                while (last - first > kBvhInsertion) {
                    if (depth == 0) {
                        bvhHeapSort(sKey, sId, first, last);
                        for (int32_t i = first; i < last; i++) atomicOr(&sStart[i >> 5], 1u << (i & 31));
                        break;
                    }
                    depth--;
                    const int32_t cut = bvhSequentialPartition(sKey, sId, first, last);
                    atomicOr(&sStart[cut >> 5], 1u << (cut & 31));
                    if (last - cut > kBvhInsertion && top < 4) { stackFirst[top] = cut; stackLast[top] = last; stackDepth[top++] = depth; }
                    last = cut;
                }
                if (top == 0) break;
                top--;
                first = stackFirst[top]; last = stackLast[top]; depth = stackDepth[top];
            }
        }
        __syncthreads();
        // final insertion pass, piece by piece, as ranks
        float rk[kBvhSmallMax / kBvhSmallThreads];
        int32_t rid[kBvhSmallMax / kBvhSmallThreads], rdst[kBvhSmallMax / kBvhSmallThreads];
        int cnt = 0;
        for (int32_t i = tid; i < m; i += kBvhSmallThreads, cnt++) {
            int32_t ps = i;
            while (!((sStart[ps >> 5] >> (ps & 31)) & 1u)) ps--;           // bit 0 is always set
            int32_t pe = i + 1;
            while (pe < m && !((sStart[pe >> 5] >> (pe & 31)) & 1u)) pe++;
            const float k = sKey[i];
            int32_t rank = 0;
            for (int32_t q = ps; q < pe; q++) {
                const float kq = sKey[q];
                rank += (kq < k || (!(k < kq) && q < i)) ? 1 : 0;
            }
            rk[cnt] = k; rid[cnt] = sId[i]; rdst[cnt] = ps + rank;
        }
        __syncthreads();
        for (int c = 0; c < cnt; c++) { keys[task.first + rdst[c]] = rk[c]; ids[task.first + rdst[c]] = rid[c]; }
    }
}

// ---- spheres ------------------------------------------------------------------------------------------------------------------
// where the sphere of a node lives: the parent's child slot (BvhNode: lc, lr | rc, rr); the root's sphere is never stored
__device__ __forceinline__ double* bvhSphereSlot(BvhNode* nodes, const BvhSeg& s) {
    return s.side == 0 ? nodes[s.parent].lc : nodes[s.parent].rc;         // [0..2] centre, [3] radius
}

// centre = (sequential float64 sum over the node's vertices, triangle by triangle in the order the parent's sort left) / 3m.
// The sum is a dependent chain of 3m additions per axis, whatever the hardware: a warp streams the node's vertices through
// shared memory (coalesced index loads, one 48-byte record per lane, the next batch in flight) and lanes 0-2 add, one axis each.
__global__ void __launch_bounds__(128)
bvhCentreKernel(int32_t n, int level, int wide, const int32_t* __restrict__ order, const float4* __restrict__ triVerts, BvhNode* nodes) {
    __shared__ double2 buf[4][3][48];                                     // per warp and axis: the 96 float64 operands of a batch of 32 triangles, in order
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint64_t unit = wide ? (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5 : uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (unit >= (uint64_t(1) << level)) return;                            // warp-uniform when wide
    const BvhSeg s = bvhSegOfSlot(n, level, uint32_t(unit));
    const int32_t m = s.e - s.b;
    if (!s.valid || m <= 1 || s.parent < 0) return;
    if (!wide) {
        double c[3] = {0.0, 0.0, 0.0};
        for (int32_t i = s.b; i < s.e; i++) {
            const int32_t id = order[i];
            for (int k = 0; k < 3; k++) {
                const float4 v = triVerts[size_t(id) * 3 + k];
                c[0] += double(v.x); c[1] += double(v.y); c[2] += double(v.z);
            }
        }
        double* out = bvhSphereSlot(nodes, s);
        const double count = double(3 * m);
        for (int a = 0; a < 3; a++) out[a] = c[a] / count;
        return;
    }
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 v0 = zero, v1 = zero, v2 = zero;
    {
        const int32_t i = s.b + int32_t(lane);
        if (i < s.e) { const int32_t id = order[i]; v0 = triVerts[size_t(id) * 3]; v1 = triVerts[size_t(id) * 3 + 1]; v2 = triVerts[size_t(id) * 3 + 2]; }
    }
    double acc = 0.0;                                                      // lanes 0-2: the axis `lane`
    for (int32_t base = s.b; base < s.e; base += 32) {
        // lanes past the end of the range stage +0.0: the running sum never is -0.0 (it starts at +0.0, and x + (-x) = +0.0),
        // so adding +0.0 leaves it unchanged and every batch runs the same 96 additions
        double* bx = reinterpret_cast<double*>(buf[warp][0]) + 3 * lane;
        double* by = reinterpret_cast<double*>(buf[warp][1]) + 3 * lane;
        double* bz = reinterpret_cast<double*>(buf[warp][2]) + 3 * lane;
        bx[0] = double(v0.x); bx[1] = double(v1.x); bx[2] = double(v2.x);
        by[0] = double(v0.y); by[1] = double(v1.y); by[2] = double(v2.y);
        bz[0] = double(v0.z); bz[1] = double(v1.z); bz[2] = double(v2.z);
        __syncwarp();
        {
            const int32_t i = base + 32 + int32_t(lane);                   // next batch: in flight while lanes 0-2 add
            v0 = zero; v1 = zero; v2 = zero;
            if (i < s.e) { const int32_t id = order[i]; v0 = triVerts[size_t(id) * 3]; v1 = triVerts[size_t(id) * 3 + 1]; v2 = triVerts[size_t(id) * 3 + 2]; }
        }
        if (lane < 3) {
            const double2* col = buf[warp][lane];                          // 128-bit shared loads: two operands each, issued ahead of the chain
#pragma unroll
            for (int q0 = 0; q0 < 48; q0 += 8) {
                double2 x[8];
#pragma unroll
                for (int q = 0; q < 8; q++) x[q] = col[q0 + q];
#pragma unroll
                for (int q = 0; q < 8; q++) { acc += x[q].x; acc += x[q].y; }
            }
        }
        __syncwarp();
    }
    if (lane < 3) bvhSphereSlot(nodes, s)[lane] = acc / double(3 * m);
}

// radius^2 = max over the node's vertices of |centre - v|^2 (float64, (x^2 + y^2) + z^2): a maximum, so any order; kept as the
// bit pattern of a non-negative double under atomicMax in the radius slot until bvhLinkKernel takes the root
__global__ void __launch_bounds__(256)
bvhRadiusKernel(int32_t n, int level, const int32_t* __restrict__ order, const float4* __restrict__ triVerts, BvhNode* nodes) {
    constexpr unsigned kFull = 0xffffffffu;
    const int64_t p64 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    uint32_t slot = 0;
    bool active = p64 < int64_t(n);
    BvhSeg s{0, 0, 0, -1, 0, false};
    if (active) {
        s = bvhSegOfPosition(n, level, int32_t(p64), slot);
        active = s.valid && s.e - s.b > 1 && s.parent >= 0;
    }
    double r2 = 0.0;
    if (active) {
        const double* c = bvhSphereSlot(nodes, s);
        const double cx = c[0], cy = c[1], cz = c[2];
        const int32_t id = order[p64];
        for (int k = 0; k < 3; k++) {
            const float4 v = triVerts[size_t(id) * 3 + k];
            const double x = cx - double(v.x), y = cy - double(v.y), z = cz - double(v.z);
            const double d = x * x + y * y + z * z;
            r2 = d > r2 ? d : r2;
        }
    }
    const unsigned act = __ballot_sync(kFull, active);
    if (act == 0) return;
    const int leader = __ffs(int(act)) - 1;
    const uint32_t slot0 = __shfl_sync(kFull, slot, leader);
    const bool uniform = __ballot_sync(kFull, active && slot != slot0) == 0;
    unsigned long long bits = (unsigned long long)__double_as_longlong(r2);
    if (uniform) {
        for (int m = 16; m >= 1; m >>= 1) {
            const unsigned long long o = __shfl_xor_sync(kFull, bits, m);
            bits = o > bits ? o : bits;
        }
        if (int(lane) != leader) return;
    } else if (!active) return;
    atomicMax(reinterpret_cast<unsigned long long*>(bvhSphereSlot(nodes, s) + 3), bits);
}

// links, leaves, and the square roots: one thread per node of the level (run for every level once all radii are in)
__global__ void __launch_bounds__(256)
bvhLinkKernel(int32_t n, int level, const int32_t* __restrict__ finalOrder, const float4* __restrict__ triVerts, BvhNode* nodes) {
    const uint64_t unit = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (unit >= (uint64_t(1) << level)) return;
    const BvhSeg s = bvhSegOfSlot(n, level, uint32_t(unit));
    if (!s.valid) return;
    const int32_t m = s.e - s.b;
    BvhNode& node = nodes[s.node];
    if (m == 1) {                                                          // leaf: TriangleMeshDistance.h:429-441
        const int32_t id = finalOrder[s.b];
        node.left = -1; node.right = id; node.pad[0] = 1; node.pad[1] = 0;
        if (s.parent < 0) return;
        double v[3][3];
        for (int k = 0; k < 3; k++) { const float4 q = triVerts[size_t(id) * 3 + k]; v[k][0] = double(q.x); v[k][1] = double(q.y); v[k][2] = double(q.z); }
        double c[3], r = 0.0;
        for (int a = 0; a < 3; a++) c[a] = (v[0][a] + v[1][a] + v[2][a]) / 3.0;
        for (int k = 0; k < 3; k++) {
            const double x = v[k][0] - c[0], y = v[k][1] - c[1], z = v[k][2] - c[2];
            const double d = sqrt(x * x + y * y + z * z);
            r = d > r ? d : r;
        }
        double* out = bvhSphereSlot(nodes, s);
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = r;
        return;
    }
    const int32_t mid = int32_t((uint32_t(s.b) + uint32_t(s.e)) >> 1);
    // a link >= 0 is an inner child, a link < 0 is ~triangleId of a leaf child (mesh_host.h): leaves are never loaded
    node.left = mid - s.b == 1 ? ~finalOrder[s.b] : s.node + 1;
    node.right = s.e - mid == 1 ? ~finalOrder[mid] : s.node + 2 * (mid - s.b);
    node.pad[0] = 0; node.pad[1] = 0;
    if (s.parent < 0) return;
    double* out = bvhSphereSlot(nodes, s);
    out[3] = sqrt(__longlong_as_double((long long)(reinterpret_cast<unsigned long long*>(out)[3])));
}

__global__ void __launch_bounds__(256) bvhIotaKernel(int32_t n, int32_t* ids) {
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < int64_t(n)) ids[i] = int32_t(i);
}

// The longest centre chains on the HOST. A dependent float64 addition takes ~30 cycles on the device and ~1 ns on a host
// core, and the chains of the top levels are the critical path of the whole build (level 1: half of the triangles each), so a
// runtime may take them: it gets the order snapshot of the level and fills centres[slot * 3 + axis] (bvhHostCentres: the same
// serial sum over the same order), which bvhPutCentresKernel stores where bvhCentreKernel would have. The radius pass follows as
// usual. The sorting — the parallel part — stays on the device, and nothing waits for the host except the radii of those levels.
template <class Vertex>   // Vertex(triangle, k) -> const float* (x, y, z)
inline void bvhHostCentres(int32_t n, int level, const int32_t* order, uint32_t slotBegin, uint32_t slotEnd, Vertex vertex, double* centres) {
    for (uint32_t slot = slotBegin; slot < slotEnd; slot++) {
        const BvhSeg s = bvhSegOfSlot(n, level, slot);
        double c[3] = {0.0, 0.0, 0.0};
        if (s.valid && s.e - s.b > 1) {
            for (int32_t i = s.b; i < s.e; i++)
                for (int k = 0; k < 3; k++) {
                    const float* v = vertex(order[i], k);
                    c[0] += double(v[0]); c[1] += double(v[1]); c[2] += double(v[2]);
                }
            const double count = double(3 * (s.e - s.b));
            for (int a = 0; a < 3; a++) c[a] /= count;
        }
        for (int a = 0; a < 3; a++) centres[size_t(slot) * 3 + a] = c[a];
    }
}
__global__ void bvhPutCentresKernel(int32_t n, int level, const double* __restrict__ centres, BvhNode* nodes) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= (1u << level)) return;
    const BvhSeg s = bvhSegOfSlot(n, level, slot);
    if (!s.valid || s.e - s.b <= 1 || s.parent < 0) return;
    double* out = bvhSphereSlot(nodes, s);
    for (int a = 0; a < 3; a++) out[a] = centres[size_t(slot) * 3 + a];
}
#ifndef BVH_HOST_CHAIN
#define BVH_HOST_CHAIN 32768
#endif
constexpr int32_t kBvhHostChain = BVH_HOST_CHAIN;   // levels whose nodes hold at least this many triangles: centre sums offered to the host

// ---- the build, as a sequence of launches -------------------------------------------------------------------------------------
// Written against a small runtime interface so that the CPU emulation of tests/cpp/simt_bvh_main.cpp drives the very same
// sequence: BVH_LAUNCH(kernel, grid, block, stream, args...) and an Rt with
//   fill(ptr, byte, bytes, stream), copy(dst, src, bytes, stream), mainStream(), sideStream(i), sideWaitsForMain(i), mainWaitsForSides(),
//   hostCentres(level, side, order, n, nodes): true = the runtime queued the centre sums of this level itself (on sideStream(side)).
struct BvhShape {
    int sortLevels = 0;           // levels 0 .. sortLevels - 1 hold nodes with more than one triangle
    int nodeLevels = 0;           // levels 0 .. nodeLevels - 1 hold nodes at all
    int32_t maxSize[34], minSize[34];
};
inline BvhShape bvhShape(int32_t n) {
    BvhShape sh;
    int32_t hi = n, lo = n;
    int l = 0;
    for (;; l++) {
        sh.maxSize[l] = hi; sh.minSize[l] = lo;
        if (hi <= 1) break;
        hi = hi - hi / 2;          // the right half: ceil
        lo = lo > 1 ? lo / 2 : 1;  // the left half: floor (a single triangle stays a leaf)
    }
    sh.sortLevels = l;
    sh.nodeLevels = l + 1;
    return sh;
}

struct BvhBuffers {
    float* keys; int32_t* ids;                 // n each: the triangle order being sorted, and its keys
    uint32_t* lpos; uint32_t* rpos;            // n each (only touched when n > kBvhSmallMax)
    int32_t* orders;                           // sortLevels x n: the order every level starts from (what its centre sums walk)
    int32_t* boxMin; int32_t* boxMax;          // 3 x 2^(sortLevels - 1) each
    BvhSortTask* big[2]; BvhSortTask* small;   // n / kBvhSmallMax + 2 each; n / 2 + 2
    uint32_t* counters;                        // 80: [0] small, [1 ..] big tasks per round
    BvhNode* nodes;                            // 2n - 1
};
constexpr int kBvhCounters = 80;

template <class Rt>
void bvhBuildLevels(Rt& rt, int32_t n, const float4* triVerts, const BvhBuffers& B, uint32_t smCount) {
    const BvhShape shape = bvhShape(n);
    auto main = rt.mainStream();
    const uint32_t perElement = uint32_t((int64_t(n) + 255) / 256);
    rt.fill(B.nodes, 0, size_t(2 * int64_t(n) - 1) * sizeof(BvhNode), main);   // radius slots start at +0.0 (atomicMax of bit patterns)
    BVH_LAUNCH(bvhIotaKernel, perElement, 256, main, n, B.ids);
    for (int l = 0; l < shape.sortLevels; l++) {
        int32_t* order = B.orders + size_t(l) * size_t(n);
        rt.copy(order, B.ids, size_t(n) * 4, main);
        if (l > 0) {   // spheres of this level's nodes: off the critical path (the next level only needs the sorted order)
            const int side = l - 1;
            rt.sideWaitsForMain(side);
            auto ss = rt.sideStream(side);
            const int wide = shape.maxSize[l] >= kBvhWarpCentre ? 1 : 0;
            const uint64_t threads = (uint64_t(1) << l) * (wide ? 32u : 1u);
            if (!(shape.minSize[l] >= kBvhHostChain && rt.hostCentres(l, side, order, n, B.nodes)))
                BVH_LAUNCH(bvhCentreKernel, uint32_t((threads + 127) / 128), 128, ss, n, l, wide, order, triVerts, B.nodes);
            BVH_LAUNCH(bvhRadiusKernel, perElement, 256, ss, n, l, order, triVerts, B.nodes);
        }
        const size_t boxBytes = (size_t(3) << l) * 4;
        rt.fill(B.boxMin, 0x7F, boxBytes, main);
        rt.fill(B.boxMax, 0x80, boxBytes, main);
        rt.fill(B.counters, 0, kBvhCounters * 4, main);
        BVH_LAUNCH(bvhBoundsKernel, perElement, 256, main, n, l, B.ids, triVerts, B.boxMin, B.boxMax);
        BVH_LAUNCH(bvhKeysKernel, perElement, 256, main, n, l, B.ids, triVerts, B.boxMin, B.boxMax, B.keys, B.big[0], B.counters + 1, B.small, B.counters);
        if (shape.maxSize[l] > kBvhSmallMax) {
            int lg = 0;
            while ((int64_t(2) << lg) <= int64_t(shape.maxSize[l])) lg++;
            const int rounds = 2 * lg + 1;   // depth limit 2 lg m partitions, then the heap-sort round
            for (int r = 0; r < rounds; r++)
                BVH_LAUNCH(bvhBigPartitionKernel, smCount, kBvhBigThreads, main, B.keys, B.ids, B.lpos, B.rpos, B.big[r & 1], B.counters + 1 + r,
                           B.big[(r + 1) & 1], B.counters + 2 + r, B.small, B.counters);
        }
        if (shape.maxSize[l] > kBvhInsertion) BVH_LAUNCH(bvhSmallSortKernel, smCount * 8u, kBvhSmallThreads, main, B.keys, B.ids, B.small, B.counters);
        if (shape.minSize[l] <= kBvhInsertion) BVH_LAUNCH(bvhTinySortKernel, perElement, 256, main, n, l, B.keys, B.ids);
    }
    rt.mainWaitsForSides();
    for (int l = 0; l < shape.nodeLevels; l++)
        BVH_LAUNCH(bvhLinkKernel, uint32_t(((uint64_t(1) << l) + 255) / 256), 256, main, n, l, B.ids, triVerts, B.nodes);
}
