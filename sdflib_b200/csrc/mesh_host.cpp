#include "mesh_host.h"
#include "host_sort.h"
#include "tri_data_build.cuh"

#include <algorithm>
#include <atomic>
#include <thread>
#include <cmath>
#include <limits>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <omp.h>
#include <parallel/algorithm>
#include <utility>

namespace sdfb200 {

// Threads of the host-side set-up steps: SDFB200_HOST_THREADS when set; otherwise the machine's cores divided by
// the ranks of this node (LOCAL_WORLD_SIZE). OMP_NUM_THREADS is deliberately not consulted: a launcher like
// torchrun pins it to 1 for every rank, which serialised TriangleData and the BVH build (9x slower at 2 ranks).
static thread_local int tHostThreadsOverride = 0;
void setHostThreadsForThisThread(int n) { tHostThreadsOverride = n; }

int hostThreads() {
    if (tHostThreadsOverride > 0) return tHostThreadsOverride;
    static const int n = [] {
        const char* e = std::getenv("SDFB200_HOST_THREADS");
        const int v = e ? std::atoi(e) : 0;
        if (v > 0) return v;
        const char* w = std::getenv("LOCAL_WORLD_SIZE");
        const int ranks = w ? std::max(1, std::atoi(w)) : 1;
        return std::max(1, omp_get_num_procs() / ranks);
    }();
    return n;
}

namespace {

inline void st3(float* dst, f3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

struct EdgeUse { uint64_t key; uint32_t corner; };   // key = min<<32 | max, corner = 3*t + k

}  // namespace

// Non-manifold repair (src/utils/TriangleUtils.cpp:292-420) over the OPEN edge uses (those left unpaired by the edge
// pairing), given in key order: vertices of open edges that lie within 1e-5/extent of each other (found through two
// staggered 2048^3 hash grids) are merged with a union-find, open edges are re-paired under the merged ids, and merged
// vertices share the sum of their normals. Returns the patches instead of applying them, so that the host path
// (computeTriangleData) and the device path (mesh_device.cu) share it; the triangles involved are rebuilt from their
// vertices with the same constructor both paths use.
OpenEdgeRepair repairOpenEdges(const HostMesh& mesh, const std::vector<OpenEdgeUse>& open, const f3* vNormalIn) {
    OpenEdgeRepair out;
    auto triOf = [&](uint32_t t) { return makeTriData(mesh.verts[mesh.idx[3 * size_t(t)]], mesh.verts[mesh.idx[3 * size_t(t) + 1]], mesh.verts[mesh.idx[3 * size_t(t) + 2]]); };
    std::map<uint32_t, uint32_t> parentOf;
    auto root = [&](uint32_t v) {
        auto it = parentOf.find(v);
        while (it != parentOf.end() && it->second != v) { v = it->second; it = parentOf.find(v); }
        return v;
    };
    std::vector<uint32_t> nm;
    for (const OpenEdgeUse& e : open) { nm.push_back(e.lo); nm.push_back(e.hi); }
    std::sort(nm.begin(), nm.end());
    nm.erase(std::unique(nm.begin(), nm.end()), nm.end());
    std::map<uint32_t, f3> vNormal;   // working copy of the normals the repair touches
    for (uint32_t v : nm) vNormal[v] = vNormalIn[v];
    f3 lo = mk3(INFINITY, INFINITY, INFINITY), hi = mk3(-INFINITY, -INFINITY, -INFINITY);
    for (uint32_t i = 0; i < mesh.nVerts; i++) {
        const f3 v = mesh.verts[i];
        lo = mk3(gmin(lo.x, v.x), gmin(lo.y, v.y), gmin(lo.z, v.z));
        hi = mk3(gmax(hi.x, v.x), gmax(hi.y, v.y), gmax(hi.z, v.z));
    }
    const f3 ext = hi - lo;
    const float maxExt = gmax(ext.x, gmax(ext.y, ext.z));
    const uint32_t res = 2048;
    const float scale = float(res) / maxExt;
    const float thr = float(1e-5 / double(maxExt));
    const float sqThr = thr * thr;
    auto cell = [&](f3 p, float off) {
        const f3 q = (p - lo) * scale;
        const int x = int(q.x + off), y = int(q.y + off), z = int(q.z + off);
        return uint64_t(uint32_t(x + y * res + z * res * res));
    };
    std::map<uint64_t, std::vector<uint32_t>> grid[2];
    for (uint32_t v : nm) { grid[0][cell(mesh.verts[v], 0.0f)].push_back(v); grid[1][cell(mesh.verts[v], 0.5f)].push_back(v); }
    for (uint32_t v : nm)
        for (int gsel = 0; gsel < 2; gsel++) {
            auto it = grid[gsel].find(cell(mesh.verts[v], gsel ? 0.5f : 0.0f));
            if (it == grid[gsel].end()) continue;
            for (uint32_t u : it->second) {
                const f3 diff = mesh.verts[v] - mesh.verts[u];
                if (dot3(diff, diff) < sqThr) {
                    const uint32_t p1 = root(v), p2 = root(u);
                    if (v == p1) parentOf[p1] = p1;
                    parentOf[p2] = p1;
                    break;
                }
            }
        }
    // open edges in key order (the reference iterates its ordered map)
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> merged;
    for (const OpenEdgeUse& e : open) {
        const uint32_t a = root(e.lo), b = root(e.hi);
        auto ins = merged.insert(std::make_pair(std::make_pair(std::min(a, b), std::max(a, b)), e.corner));
        if (!ins.second) {
            const uint32_t t = e.corner / 3, t2 = ins.first->second / 3;
            const TriData dt = triOf(t), dt2 = triOf(t2);
            const f3 n = triNormal(dt) + triNormal(dt2);
            out.edgeCorner.push_back(e.corner); out.edgeNormal.push_back(matMul(dt.T, n));
            out.edgeCorner.push_back(ins.first->second); out.edgeNormal.push_back(matMul(dt2.T, n));
            merged.erase(ins.first);
        }
    }
    for (uint32_t v : nm) { const uint32_t p = root(v); if (p != v) vNormal[p] = vNormal[p] + vNormal[v]; }
    for (uint32_t v : nm) vNormal[v] = vNormal[root(v)];
    for (uint32_t v : nm) { out.vertex.push_back(v); out.vertexNormal.push_back(vNormal[v]); }
    return out;
}

TriVec computeTriangleData(const HostMesh& mesh) {
    const uint32_t nT = mesh.numTriangles();
    const bool timing = std::getenv("SDFB200_TIMING") != nullptr;
    auto tick = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[sdfb200] triangle data: %-18s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tick).count());
        tick = now;
    };
    TriVec tris(nT);
    RawVec<f3> cornerContribution(size_t(nT) * 3);
    RawVec<EdgeUse> uses(size_t(nT) * 3);

    // Frames and per-corner angle * normal are pure per-triangle functions: parallel.
#pragma omp parallel for schedule(static) num_threads(hostThreads()) if (nT > 8192)
    for (int64_t t = 0; t < int64_t(nT); t++) {
        const uint32_t* ix = mesh.idx + 3 * t;
        tris[size_t(t)] = makeTriData(mesh.verts[ix[0]], mesh.verts[ix[1]], mesh.verts[ix[2]]);
        const f3 n = triNormal(tris[size_t(t)]);
        for (uint32_t k = 0; k < 3; k++) {
            const uint32_t a = ix[k], b = ix[(k + 1) % 3], c = ix[(k + 2) % 3];
            const float cosA = dot3(normalize3(mesh.verts[b] - mesh.verts[a]), normalize3(mesh.verts[c] - mesh.verts[a]));
            const float angle = std::acos(gmin(gmax(cosA, -1.0f), 1.0f));   // == acosfLibm (tri_data_build.cuh), checked in tests/cpp
            cornerContribution[size_t(3 * t + k)] = angle * n;
            uses[size_t(3 * t + k)] = EdgeUse{(uint64_t(std::min(a, b)) << 32) | std::max(a, b), uint32_t(3 * t + k)};
        }
    }

    lap("frames");
    // Edge pairing. The reference walks the corners in order through an ordered map: an edge seen
    // while its key is stored pairs with the stored corner and both get n_t + n_t'; the key is then
    // erased, so occurrences pair up (1st,2nd), (3rd,4th), ... and an odd one stays open. Sorting the
    // uses by (edge, corner) reproduces exactly those pairs without the serial map.
    // (key, corner) pairs are unique, so any correct sort gives the same sequence: use the multi-threaded one
    // libstdc++'s parallel mode goes sequential when omp_get_max_threads() == 1, whatever the tag asks for (torchrun
    // exports OMP_NUM_THREADS=1 to every rank): raise this thread's nthreads-var for the call and put it back.
    const int callerThreads = omp_get_max_threads();
    omp_set_num_threads(hostThreads());
    __gnu_parallel::sort(uses.begin(), uses.end(), [](const EdgeUse& x, const EdgeUse& y) {
        return x.key != y.key ? x.key < y.key : x.corner < y.corner;
    }, __gnu_parallel::default_parallel_tag(hostThreads()));
    omp_set_num_threads(callerThreads);
    lap("edge sort");
    // Groups of equal keys are independent: cut the sorted array into chunks at group boundaries, one per thread;
    // the open (unpaired) uses are concatenated in chunk order, i.e. still in key order.
    const int nChunks = std::max(1, std::min(hostThreads(), int(uses.size() / 4096) + 1));
    std::vector<size_t> cut;
    cut.resize(size_t(nChunks) + 1);
    for (int c = 0; c <= nChunks; c++) {
        size_t at = uses.size() * size_t(c) / size_t(nChunks);
        while (at > 0 && at < uses.size() && uses[at].key == uses[at - 1].key) at++;
        cut[size_t(c)] = at;
    }
    std::vector<std::vector<EdgeUse>> openOf;
    openOf.resize(size_t(nChunks));
#pragma omp parallel for schedule(static, 1) num_threads(nChunks) if (nChunks > 1)
    for (int c = 0; c < nChunks; c++) {
        for (size_t i = cut[size_t(c)]; i < cut[size_t(c) + 1];) {
            size_t j = i;
            while (j < uses.size() && uses[j].key == uses[i].key) j++;
            size_t p = i;
            for (; p + 1 < j; p += 2) {
                const uint32_t first = uses[p].corner, second = uses[p + 1].corner;
                const uint32_t t2 = first / 3, t = second / 3;
                const f3 n = triNormal(tris[t]) + triNormal(tris[t2]);
                st3(tris[t].edgesNormal[second % 3], matMul(tris[t].T, n));
                st3(tris[t2].edgesNormal[first % 3], matMul(tris[t2].T, n));
            }
            if (p < j) openOf[size_t(c)].push_back(uses[p]);
            i = j;
        }
    }
    std::vector<EdgeUse> open;
    for (const auto& o : openOf) open.insert(open.end(), o.begin(), o.end());
    lap("edge pairing");
    // Angle-weighted vertex normals, accumulated in corner order (float addition order matters).
    std::vector<f3> vNormal(mesh.nVerts, mk3(0.f, 0.f, 0.f));
    for (size_t cix = 0; cix < size_t(nT) * 3; cix++) {
        const uint32_t a = mesh.idx[cix];
        vNormal[a] = vNormal[a] + cornerContribution[cix];
    }

    lap("vertex normals");
    if (!open.empty()) {
        std::vector<OpenEdgeUse> uses(open.size());
        for (size_t i = 0; i < open.size(); i++) uses[i] = OpenEdgeUse{uint32_t(open[i].key >> 32), uint32_t(open[i].key), open[i].corner};
        const OpenEdgeRepair rep = repairOpenEdges(mesh, uses, vNormal.data());
        for (size_t i = 0; i < rep.edgeCorner.size(); i++) st3(tris[rep.edgeCorner[i] / 3].edgesNormal[rep.edgeCorner[i] % 3], rep.edgeNormal[i]);
        for (size_t i = 0; i < rep.vertex.size(); i++) vNormal[rep.vertex[i]] = rep.vertexNormal[i];
    }

#pragma omp parallel for schedule(static) num_threads(hostThreads()) if (nT > 8192)
    for (int64_t i = 0; i < int64_t(mesh.nIdx); i++)
        st3(tris[size_t(i / 3)].verticesNormal[i % 3], matMul(tris[size_t(i / 3)].T, vNormal[mesh.idx[i]]));
    lap("to triangle frame");
    return tris;
}

// ---------------------------------------------------------------------------------------------
// BVH (float64 bounding spheres, median split on the first vertex along the widest axis)
// ---------------------------------------------------------------------------------------------
namespace {

struct BuildTri { double v[3][3]; };          // by triangle id, never moved
struct SortKey { double key; int32_t id; };     // what std::sort actually moves (16 bytes instead of 80)

inline double sq3(const double* a, const double* b) {
    const double x = a[0] - b[0], y = a[1] - b[1], z = a[2] - b[2];
    return x * x + y * y + z * z;
}

struct BvhBuilder {
    const RawVec<BuildTri>& tri;   // vertices in float64, by triangle id
    RawVec<int32_t>& order;        // current triangle order; a node owns the range [begin, end)
    RawVec<SortKey>& scratch;      // same ranges, disjoint between threads
    RawVec<BvhNode>& nodes;
    int forkLevels;                // levels below this node that still fork a thread for the left half

    // sphere = where the bounding sphere of this subtree is stored (a child slot of the parent)
    void build(int32_t nodeId, double* sphereCenter, double* sphereRadius, int32_t begin, int32_t end) {
        const int32_t n = end - begin;
        BvhNode& node = nodes[size_t(nodeId)];
        if (n == 1) {
            const BuildTri& t = tri[size_t(order[size_t(begin)])];
            double c[3];
            for (int a = 0; a < 3; a++) c[a] = (t.v[0][a] + t.v[1][a] + t.v[2][a]) / 3.0;
            const double r0 = std::sqrt(sq3(t.v[0], c)), r1 = std::sqrt(sq3(t.v[1], c)), r2 = std::sqrt(sq3(t.v[2], c));
            for (int a = 0; a < 3; a++) sphereCenter[a] = c[a];
            *sphereRadius = std::max(std::max(r0, r1), r2);
            node.left = -1;
            node.right = order[size_t(begin)];
            node.pad[0] = 1;   // leaf
            node.pad[1] = 0;
            return;
        }
        double top[3], bottom[3], c[3] = {0, 0, 0};
        for (int a = 0; a < 3; a++) { top[a] = std::numeric_limits<double>::lowest(); bottom[a] = std::numeric_limits<double>::max(); }
        for (int32_t i = begin; i < end; i++) {   // the centre is a sequential float64 sum: its order is part of the result
            const BuildTri& t = tri[size_t(order[size_t(i)])];
            for (int k = 0; k < 3; k++)
                for (int a = 0; a < 3; a++) {
                    const double p = t.v[k][a];
                    c[a] += p;
                    top[a] = std::max(top[a], p);
                    bottom[a] = std::min(bottom[a], p);
                }
        }
        const double count = double(3 * n);
        for (int a = 0; a < 3; a++) c[a] /= count;
        int dim = 0;
        for (int a = 1; a < 3; a++)
            if (top[a] - bottom[a] > top[dim] - bottom[dim]) dim = a;
        // Threads this node may use: the whole team at the root, half of it per child, ... (the halves run side by side).
        const int team = std::max(1, (1 << forkLevels) / 2);
        SortKey* keys = scratch.data() + begin;
        double r2 = 0.0;   // a maximum: any evaluation order gives the same value
        {
            std::vector<double> part(size_t(team), 0.0);
            std::atomic<int> slot{0};
            forChunks(n, team, [&](int32_t lo, int32_t hi) {
                double m = 0.0;
                for (int32_t i = lo; i < hi; i++) {
                    const int32_t id = order[size_t(begin + i)];
                    const BuildTri& t = tri[size_t(id)];
                    for (int k = 0; k < 3; k++) m = std::max(m, sq3(c, t.v[k]));
                    keys[i] = SortKey{t.v[0][dim], id};
                }
                part[size_t(slot.fetch_add(1))] = m;
            });
            for (double m : part) r2 = std::max(r2, m);
        }
        for (int a = 0; a < 3; a++) sphereCenter[a] = c[a];
        *sphereRadius = std::sqrt(r2);
        // Same call as the reference: std::sort is not stable, and triangles sharing their first vertex tie on this
        // key, so the library's comparison sequence is part of the result. That sequence depends only on the
        // comparison outcomes, not on the element type, so 16-byte (key, id) records give the same permutation as
        // sorting the reference's 80-byte triangle records.
        sortLikeStd(keys, keys + n, [](const SortKey& x, const SortKey& y) { return x.key < y.key; }, team);
        for (int32_t i = 0; i < n; i++) order[size_t(begin + i)] = keys[i].id;
        const int32_t mid = int32_t(0.5 * (begin + end));
        node.left = nodeId + 1;
        node.right = nodeId + 2 * (mid - begin);
        node.pad[0] = node.pad[1] = 0;
        const int32_t l = node.left, r = node.right;
        double* lc = node.lc;
        double* lr = &node.lr;
        // The two halves are independent once sorted and all arrays are pre-sized, so they only touch disjoint ranges.
        // Plain threads down to `forkLevels` levels (about 2 x hostThreads leaves of the fork tree), not an OpenMP
        // task team: a team's idle threads SPIN while the master runs the sequential top-level sorts, which starves
        // whatever else the host is doing (TriangleData of the same constructor, the other ranks of a multi-GPU job).
        std::vector<std::thread> fork;
        if (forkLevels > 0 && n > 8192 &&
            tryFork(fork, [=] { BvhBuilder sub{tri, order, scratch, nodes, forkLevels - 1}; sub.build(l, lc, lr, begin, mid); })) {
            BvhBuilder sub{tri, order, scratch, nodes, forkLevels - 1};
            sub.build(r, node.rc, &node.rr, mid, end);
            fork[0].join();
        } else {
            build(l, lc, lr, begin, mid);
            build(r, node.rc, &node.rr, mid, end);
        }
    }
};

}  // namespace

RawVec<BvhNode> buildBvh(const HostMesh& mesh) {
    const uint32_t nT = mesh.numTriangles();
    const bool timing = std::getenv("SDFB200_TIMING") != nullptr;
    auto tick = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!timing) return;
        const auto now = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[sdfb200] bvh: %-27s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tick).count());
        tick = now;
    };
    RawVec<BuildTri> bt(nT);
    RawVec<int32_t> order(nT);
    RawVec<SortKey> scratch(nT);
    forChunks(int32_t(nT), hostThreads(), [&](int32_t lo, int32_t hi) {
        for (int32_t t = lo; t < hi; t++) {
            order[size_t(t)] = t;
            for (int k = 0; k < 3; k++) {
                const f3 p = mesh.verts[mesh.idx[3 * size_t(t) + size_t(k)]];
                bt[size_t(t)].v[k][0] = double(p.x); bt[size_t(t)].v[k][1] = double(p.y); bt[size_t(t)].v[k][2] = double(p.z);
            }
        }
    });
    lap("float64 records");
    RawVec<BvhNode> nodes(size_t(2) * nT - 1);
    double rootCenter[3], rootRadius;
    int forkLevels = 1;
    while ((1 << forkLevels) < 2 * hostThreads()) forkLevels++;
    BvhBuilder b{bt, order, scratch, nodes, forkLevels};
    b.build(0, rootCenter, &rootRadius, 0, int32_t(nT));
    lap("tree");
    // Device traversal never loads a leaf node: links to leaves are replaced by ~triangleId (mesh_host.h).
    forChunks(int32_t(nodes.size()), hostThreads(), [&](int32_t lo, int32_t hi) {
        for (int32_t i = lo; i < hi; i++) {
            BvhNode& nd = nodes[size_t(i)];
            if (nd.pad[0]) continue;
            const BvhNode& l = nodes[size_t(nd.left)];
            const BvhNode& r = nodes[size_t(nd.right)];
            if (l.pad[0]) nd.left = ~l.right;
            if (r.pad[0]) nd.right = ~r.right;
        }
    });
    lap("leaf links");
    return nodes;
}

}  // namespace sdfb200
