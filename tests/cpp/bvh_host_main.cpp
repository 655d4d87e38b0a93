// Host-side check of sdfb200::buildBvh (sdflib_b200/csrc/mesh_host.cpp) against a plain serial restatement of
// tmd::TriangleMeshDistance::_build_tree (libs/InteractiveComputerGraphics/TriangleMeshDistance/.../TriangleMeshDistance.h:421-490):
// 80-byte triangle records moved by ONE std::sort call per node, exactly like the reference. The product builder
// sorts 16-byte proxies and splits the sort's recursion across threads; both must leave every node identical,
// including the order std::sort gives to triangles that tie on their first vertex (isospheres are full of those).
//
//   bvh_host_main <isosphere subdivisions> <displace 0|1> [threads]     prints "ok <nodes> <serial ms> <product ms>"
//   bvh_host_main sort                                                  the threaded sort alone, see sortCheck()
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "host_sort.h"
#include "mesh_host.h"
#include "sdfb200.h"

using namespace sdfb200;

namespace {

struct Triangle { std::array<std::array<double, 3>, 3> vertices; int id = -1; };
struct RefNode { double lc[3], lr, rc[3], rr; int left = -1, right = -1; };

struct RefBuilder {
    std::vector<RefNode> nodes;
    std::vector<Triangle> triangles;

    void build(int nodeId, double* center, double* radius, int begin, int end) {
        const int n = end - begin;
        if (n == 1) {
            const Triangle& t = triangles[size_t(begin)];
            double c[3], r[3];
            for (int a = 0; a < 3; a++) c[a] = (t.vertices[0][a] + t.vertices[1][a] + t.vertices[2][a]) / 3.0;
            for (int k = 0; k < 3; k++) {
                double s = 0.0;
                for (int a = 0; a < 3; a++) { const double d = t.vertices[k][a] - c[a]; s += d * d; }
                r[k] = std::sqrt(s);
            }
            for (int a = 0; a < 3; a++) center[a] = c[a];
            *radius = std::max(std::max(r[0], r[1]), r[2]);
            nodes[size_t(nodeId)].left = -1;
            nodes[size_t(nodeId)].right = t.id;
            return;
        }
        double top[3], bottom[3], c[3] = {0, 0, 0};
        for (int a = 0; a < 3; a++) { top[a] = std::numeric_limits<double>::lowest(); bottom[a] = std::numeric_limits<double>::max(); }
        for (int i = begin; i < end; i++)
            for (int k = 0; k < 3; k++)
                for (int a = 0; a < 3; a++) {
                    const double p = triangles[size_t(i)].vertices[k][a];
                    c[a] += p;
                    top[a] = std::max(top[a], p);
                    bottom[a] = std::min(bottom[a], p);
                }
        for (int a = 0; a < 3; a++) c[a] /= double(3 * n);
        double diag[3] = {top[0] - bottom[0], top[1] - bottom[1], top[2] - bottom[2]};
        const int dim = int(std::max_element(diag, diag + 3) - diag);
        double r2 = 0.0;
        for (int i = begin; i < end; i++)
            for (int k = 0; k < 3; k++) {
                double s = 0.0;
                for (int a = 0; a < 3; a++) { const double d = c[a] - triangles[size_t(i)].vertices[k][a]; s += d * d; }
                r2 = std::max(r2, s);
            }
        for (int a = 0; a < 3; a++) center[a] = c[a];
        *radius = std::sqrt(r2);
        std::sort(triangles.begin() + begin, triangles.begin() + end,
                  [dim](const Triangle& x, const Triangle& y) { return x.vertices[0][size_t(dim)] < y.vertices[0][size_t(dim)]; });
        const int mid = int(0.5 * (begin + end));
        const int l = int(nodes.size());
        nodes.push_back(RefNode());
        nodes[size_t(nodeId)].left = l;
        {
            double lc[3], lr;
            build(l, lc, &lr, begin, mid);
            std::memcpy(nodes[size_t(nodeId)].lc, lc, sizeof lc);
            nodes[size_t(nodeId)].lr = lr;
        }
        const int r = int(nodes.size());
        nodes.push_back(RefNode());
        nodes[size_t(nodeId)].right = r;
        {
            double rc[3], rr;
            build(r, rc, &rr, mid, end);
            std::memcpy(nodes[size_t(nodeId)].rc, rc, sizeof rc);
            nodes[size_t(nodeId)].rr = rr;
        }
    }
};

double nowMs() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// sortLikeStd / sortPiece (host_sort.h) against the library: unstable-sort permutations must be IDENTICAL, on inputs
// with few distinct keys (ties everywhere), presorted / reversed / saw-tooth inputs, and with the depth limit forced
// low so that the heap-sort fallback runs inside forked pieces.
struct Rec { double key; int id; };
int sortCheck() {
    auto less = [](const Rec& a, const Rec& b) { return a.key < b.key; };
    uint32_t rng = 12345u;
    auto next = [&] { rng = rng * 1664525u + 1013904223u; return rng >> 8; };
    int cases = 0;
    for (int n : {0, 1, 17, 1000, 70001, 300000})
        for (int pattern = 0; pattern < 6; pattern++) {
            std::vector<Rec> in(static_cast<size_t>(n));
            for (int i = 0; i < n; i++) {
                double k = 0;
                switch (pattern) {
                    case 0: k = double(next()); break;                       // distinct
                    case 1: k = double(next() % 7); break;                   // 7 values
                    case 2: k = double(next() % 1000); break;                // many ties
                    case 3: k = double(i / 3); break;                        // presorted triples
                    case 4: k = double((n - i) / 5); break;                  // reversed
                    default: k = double(i % 4096 < 2048 ? i % 4096 : 4096 - i % 4096); break;   // organ pipes
                }
                in[size_t(i)] = Rec{k, i};
            }
            std::vector<Rec> want = in;
            std::sort(want.begin(), want.end(), less);
            for (int threads : {1, 2, 5, 16}) {
                std::vector<Rec> got = in;
                sortLikeStd(got.begin(), got.end(), less, threads);
                for (int i = 0; i < n; i++)
                    if (got[size_t(i)].id != want[size_t(i)].id) { std::fprintf(stderr, "sort n=%d pattern=%d threads=%d differs at %d\n", n, pattern, threads, i); return 1; }
                cases++;
            }
#if defined(__GLIBCXX__)
            if (n > 16)
                for (long limit : {0L, 1L, 3L, 8L}) {
                    auto comp = __gnu_cxx::__ops::__iter_comp_iter(less);
                    std::vector<Rec> lib = in, got = in;
                    std::__introsort_loop(lib.begin(), lib.end(), limit, comp);
                    std::__final_insertion_sort(lib.begin(), lib.end(), comp);
                    sortPiece(got.begin(), got.end(), limit, comp, 2048L);
                    for (int i = 0; i < n; i++)
                        if (got[size_t(i)].id != lib[size_t(i)].id) { std::fprintf(stderr, "depth limit %ld n=%d pattern=%d differs at %d\n", limit, n, pattern, i); return 1; }
                    cases++;
                }
#endif
        }
    std::printf("ok %d sort cases\n", cases);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc > 1 && std::strcmp(argv[1], "sort") == 0) return sortCheck();
    const uint32_t subdivisions = argc > 1 ? uint32_t(std::atoi(argv[1])) : 3;
    const bool displace = argc > 2 && std::atoi(argv[2]) != 0;
    if (argc > 3) setenv("SDFB200_HOST_THREADS", argv[3], 1);
    uint32_t nv = 0, ni = 0;
    sdfb200_make_isosphere(subdivisions, nullptr, nullptr, &nv, &ni);
    std::vector<float> verts(size_t(nv) * 3);
    std::vector<uint32_t> idx(ni);
    sdfb200_make_isosphere(subdivisions, verts.data(), idx.data(), &nv, &ni);
    if (displace)   // generic position: no two first vertices share a coordinate by construction of the lattice
        for (uint32_t v = 0; v < nv; v++) {
            float* p = &verts[size_t(v) * 3];
            const float d = 0.15f * std::sin(3.f * p[0]) * std::sin(3.f * p[1] + 2.1f) * std::sin(3.f * p[2] + 0.7f);
            for (int a = 0; a < 3; a++) p[a] = p[a] * (1.f + d);
        }

    RefBuilder ref;
    ref.triangles.resize(ni / 3);
    for (uint32_t t = 0; t < ni / 3; t++) {
        ref.triangles[t].id = int(t);
        for (int k = 0; k < 3; k++)
            for (int a = 0; a < 3; a++) ref.triangles[t].vertices[size_t(k)][size_t(a)] = double(verts[size_t(idx[3 * t + uint32_t(k)]) * 3 + size_t(a)]);
    }
    ref.nodes.reserve(size_t(2) * (ni / 3));
    ref.nodes.push_back(RefNode());
    double rootC[3], rootR;
    const double t0 = nowMs();
    ref.build(0, rootC, &rootR, 0, int(ni / 3));
    const double t1 = nowMs();

    HostMesh mesh{reinterpret_cast<const f3*>(verts.data()), nv, idx.data(), ni};
    RawVec<BvhNode> got;
    double best = 1e30;   // best of five: the first call pays page faults and thread start-up
    for (int rep = 0; rep < 5; rep++) {
        const double t2 = nowMs();
        got = buildBvh(mesh);
        best = std::min(best, nowMs() - t2);
    }

    if (got.size() != ref.nodes.size()) { std::fprintf(stderr, "node count %zu != %zu\n", got.size(), ref.nodes.size()); return 1; }
    for (size_t i = 0; i < got.size(); i++) {
        const RefNode& r = ref.nodes[i];
        const BvhNode& g = got[i];
        if (r.left == -1) {   // leaf node: never loaded by the device traversal, but its id must still be right
            if (!g.pad[0] || g.right != r.right) { std::fprintf(stderr, "leaf %zu differs\n", i); return 1; }
            continue;
        }
        // product links: >= 0 inner child index, < 0 ~triangleId of a leaf child
        const RefNode& rl = ref.nodes[size_t(r.left)];
        const RefNode& rr = ref.nodes[size_t(r.right)];
        const int wantL = rl.left == -1 ? ~rl.right : r.left;
        const int wantR = rr.left == -1 ? ~rr.right : r.right;
        if (g.pad[0] || g.left != wantL || g.right != wantR || std::memcmp(g.lc, r.lc, 24) || std::memcmp(&g.lr, &r.lr, 8) ||
            std::memcmp(g.rc, r.rc, 24) || std::memcmp(&g.rr, &r.rr, 8)) {
            std::fprintf(stderr, "inner node %zu differs\n", i);
            return 1;
        }
    }
    std::printf("ok %zu %.2f %.2f\n", got.size(), t1 - t0, best);
    return 0;
}
