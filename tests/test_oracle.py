"""CPU tests of the oracle: the restatement (oracle/oracle.cpp, "port") against the committed golden
vectors generated from the reference, and — where oracle/_ref exists — against the unmodified
reference itself, bit for bit. No GPU involved."""
import os

import numpy as np
import pytest

from conftest import assert_bit_equal, golden, displaced_sphere, octree_topology, edge_case_meshes


# ---- port vs golden vectors (works without the reference) -----------------------------------------
def test_isosphere_matches_reference_fixture(port):
    g = golden("mesh_small.npz")
    v, i = port.isosphere(2)
    assert_bit_equal(v, g["sphere_vertices"], "isosphere vertices")
    assert np.array_equal(i, g["indices"])
    assert i.size == 960 and len(v) == 162   # config 1: 320 triangles


def test_triangle_data_golden(port):
    g = golden("mesh_small.npz")
    assert_bit_equal(port.triangle_data(g["vertices"], g["indices"]), g["triangle_data"], "TriangleData")
    k = golden("kernels.npz")
    assert_bit_equal(port.triangle_data(k["tet_vertices"], k["tet_indices"]), k["tet_triangle_data"], "tet TriangleData")


def test_point_triangle_golden_all_regions(port):
    k = golden("kernels.npz")
    td, pts = k["tet_triangle_data"][0], k["points"]
    w = k["tet_vertices"][k["tet_indices"][:3]].reshape(-1)
    assert_bit_equal(port.sq_dist(td, pts), k["sq_dist"], "getSqDistPointAndTriangle")
    d0, _ = port.signed_dist(td, w, pts, 0)
    d1, g1 = port.signed_dist(td, w, pts, 1)
    d2, g2 = port.signed_dist(td, w, pts, 2)
    assert_bit_equal(d0, k["signed0"]); assert_bit_equal(d1, k["signed1"]); assert_bit_equal(g1, k["grad1"])
    assert_bit_equal(d2, k["signed2"]); assert_bit_equal(g2, k["grad2"])
    # the reference's own TriangleDistanceTest tolerance (src/tools/TriangleDistanceTest/main.cpp:59-63)
    assert np.all(np.abs(d0 * d0 - k["sq_dist"]) < 1e-3)
    # the sample covers all 7 Voronoi regions: distinct gradients for vertex / edge / face cases
    assert len(np.unique(np.round(g2, 3), axis=0)) > 100


def test_tricubic_golden(port):
    k = golden("kernels.npz")
    c = port.tricubic_coefficients(k["corner_values"], float(k["node_size"]))
    assert_bit_equal(c, k["coefficients"], "calculateCoefficients")
    v, g, vv = port.tricubic_eval(c, k["frac"], float(k["node_size"]))
    assert_bit_equal(v, k["value"]); assert_bit_equal(g, k["gradient"]); assert_bit_equal(vv, k["vertex_values"])
    e = np.float32([port.error_estimate(c, k["mid_values"], 1), port.error_estimate(c, k["mid_values"], 3, 0.1)])
    assert_bit_equal(e, k["error"], "error integrals")


def test_hermite_fit_interpolates_corners(port):
    """Property of the fit: the polynomial reproduces the corner values and (scaled) first derivatives."""
    rng = np.random.default_rng(1)
    vals = rng.standard_normal((8, 8)).astype(np.float32)
    vals[:, 4:] = 0
    s = 0.5
    c = port.tricubic_coefficients(vals, s)
    corners = np.float32([[c_ & 1, (c_ >> 1) & 1, c_ >> 2] for c_ in range(8)])
    v, g, vv = port.tricubic_eval(c, corners, s)
    assert np.allclose(v, vals[:, 0], atol=2e-5)
    assert np.allclose(vv[:, 1:4], vals[:, 1:4], atol=1e-4)


def test_nearest_triangle_and_filter_golden(port):
    g = golden("mesh_small.npz")
    near = port.nearest_triangle(g["vertices"], g["indices"], g["query_points"])
    assert np.array_equal(near, g["nearest"])
    kept = port.filter_triangles(g["vertices"], g["indices"], g["filter_centre"], float(g["filter_half"]),
                                 np.arange(g["indices"].size // 3, dtype=np.uint32), g["filter_corner_tris"])
    assert np.array_equal(kept, g["filter_kept"])
    # brute force over all triangles (RealSdf definition, src/sdf/RealSdf.cpp:10-25) agrees on the distance
    td = g["triangle_data"]
    p = g["query_points"][:32]
    brute = np.stack([port.sq_dist(td[t], p) for t in range(len(td))]).min(0)
    chosen = np.array([port.sq_dist(td[t], p[k:k + 1])[0] for k, t in enumerate(near[:32])])
    assert np.allclose(brute, chosen, rtol=1e-5, atol=1e-7)


def test_config1_octree_golden(port):
    """Config 1 verbatim: 1 144 552 words, hash, header scalars and queries of the reference's single-thread build."""
    import hashlib
    g = golden("config1_octree.npz")
    v, i = port.isosphere(2)
    s = port.build_octree(v, i, g["box"], 5, 3, 1e-3, 1, 1, use_cache=True)
    d = s.octree_data()
    assert d.size == int(g["words"]) == 1144552
    assert hashlib.sha256(d.tobytes()).hexdigest() == str(g["sha256"])
    assert np.array_equal(d[:512], g["start_slots"])
    h = s.header()
    assert_bit_equal(np.float32([h["value_range"], h["min_border_value"]]), np.float32([g["value_range"], g["min_border_value"]]))
    topo, leaves, inner = octree_topology(d, 8)
    assert (leaves, inner) == (17571, 2437)           # SURVEY.md §6 probe numbers
    assert topo.sum() + 64 * leaves == d.size         # every word is reachable exactly once
    dist, grad = s.query(g["query_points"], True)
    assert_bit_equal(dist, g["distances"]); assert_bit_equal(grad, g["gradients"])


def test_bin_files_golden(port, tmp_path):
    """Byte-exact .bin files for both formats, and load -> save round trip."""
    g = golden("small_structures.npz")
    gm = golden("mesh_small.npz")
    v, i = gm["vertices"], gm["indices"]
    o = port.build_octree(v, i, g["box"], 4, 2, 1e-3, 1, 1, use_cache=True)
    e = port.build_exact(v, i, g["box"], 4, 1, 16, 1, use_cache=True)
    for s, key in ((o, "octree_bin"), (e, "exact_bin")):
        p = str(tmp_path / (key + ".bin"))
        assert s.save(p)
        assert open(p, "rb").read() == g[key].tobytes(), key
        s2 = port.load(p)
        p2 = str(tmp_path / (key + "2.bin"))
        s2.save(p2)
        assert open(p2, "rb").read() == g[key].tobytes(), key + " round trip"
    q = g["query_points"]
    od, og = o.query(q, True); ed, eg = e.query(q, True)
    assert_bit_equal(od, g["octree_distances"]); assert_bit_equal(og, g["octree_gradients"])
    assert_bit_equal(ed, g["exact_distances"]); assert_bit_equal(eg, g["exact_gradients"])
    # exact distances never exceed |tri-cubic distance| + threshold, and agree closely inside the box
    inside = np.all((q > o.sample_area()[:3]) & (q < o.sample_area()[3:]), axis=1)
    assert np.abs(od[inside] - ed[inside]).max() < 0.05


def test_history_free_variant_keeps_topology(port):
    """use_cache=False (what the GPU build is compared with) has the reference's topology; only tie-broken
    gradients move, so queries stay within the parity gate of SURVEY.md §8(d)."""
    v, i = displaced_sphere(3)
    from sdflib_b200 import meshes
    box = meshes.bounding_box_with_margin(v)
    a = port.build_octree(v, i, box, 5, 2, 1e-3, 1, 1, use_cache=True)
    b = port.build_octree(v, i, box, 5, 2, 1e-3, 1, 1, use_cache=False)
    da, db = a.octree_data(), b.octree_data()
    assert da.size == db.size
    ta, _, _ = octree_topology(da, 4)
    tb, _, _ = octree_topology(db, 4)
    assert np.array_equal(ta, tb) and np.array_equal(da[ta], db[tb])
    area = a.sample_area()
    q = (area[:3] + np.random.default_rng(3).uniform(0, 1, (50000, 3)) * (area[3:] - area[:3])).astype(np.float32)
    qa, qb = a.query(q), b.query(q)
    tol = 1e-5 * np.maximum(np.abs(qa), 1e-3 * (area[3] - area[0]))
    assert (np.abs(qa - qb) > tol).mean() < 1e-4


def test_mt_layout_is_a_relayout(port):
    """numThreads >= 2 selects the per-start-voxel layout: same tree, blocks grouped per voxel."""
    v, i = displaced_sphere(2)
    from sdflib_b200 import meshes
    box = meshes.bounding_box_with_margin(v)
    a = port.build_octree(v, i, box, 5, 3, 1e-3, 1, 1).octree_data()
    b = port.build_octree(v, i, box, 5, 3, 1e-3, 1, 2).octree_data()
    assert a.size == b.size and not np.array_equal(a, b)

    def canon(d):
        out, stack = [], list(range(511, -1, -1))
        while stack:
            w = int(d[stack.pop()])
            c = w & 0x3FFFFFFF
            if w & 0x80000000:
                out.append(d[c:c + 64])
            else:
                stack.extend(range(c + 7, c - 1, -1))
        return np.concatenate(out)
    assert np.array_equal(canon(a), canon(b))


def test_continuity_golden(port):
    """InitAlgorithm::CONTINUITY and the three termination rules against reference-generated fixtures."""
    g = golden("continuity_small.npz")
    import hashlib
    v, i = displaced_sphere(2)
    c = port.tricubic_coefficients(g["corner_values"], float(g["node_size"]))
    assert_bit_equal(c, g["coefficients"], "calculateCoefficients with mixed derivatives")
    e = np.float32([port.error_estimate(c, g["mid_values"], r, 0.1) for r in (1, 2, 3)])
    assert_bit_equal(e, g["error"], "trapezoid / Simpson / by-distance integrals")
    for name, rule, p1 in (("trapezoid", 1, 0.0), ("simpson", 2, 0.0), ("by_distance", 3, 0.1)):
        a = port.build_octree(v, i, g["box"], 5, 3, 1e-3, 2, 1, termination_rule=rule, param1=p1, use_cache=True)
        d = a.octree_data()
        assert d.size == int(g[name + "_words"])
        assert hashlib.sha256(d.tobytes()).hexdigest() == str(g[name + "_sha256"]), name
        assert_bit_equal(np.float32(a.header()["min_border_value"]), g[name + "_min_border_value"])
        if rule == 1:
            assert np.array_equal(d[:512], g["start_slots"])
            dist, grad = a.query(g["query_points"], True)
            assert_bit_equal(dist, g["distances"]); assert_bit_equal(grad, g["gradients"])


def test_continuity_field_is_continuous_across_t_junctions(port):
    """The property the algorithm exists for: across every face of the start grid the tri-cubic field is
    continuous (value jump far below the termination threshold), which NO_CONTINUITY does not give."""
    v, i = displaced_sphere(2)
    from sdflib_b200 import meshes
    box = meshes.bounding_box_with_margin(v)
    jumps = {}
    for alg in (1, 2):
        a = port.build_octree(v, i, box, 5, 3, 1e-3, alg, 1, use_cache=False)
        area = a.sample_area()
        rng = np.random.default_rng(9)
        n = 20000
        uv = rng.uniform(0.02, 0.98, (n, 3))
        plane = rng.integers(1, 32, n) / 32.0   # faces of depth-5 cells
        axis = rng.integers(0, 3, n)
        eps = 2e-6
        lo, hi = uv.copy(), uv.copy()
        lo[np.arange(n), axis] = plane - eps
        hi[np.arange(n), axis] = plane + eps
        size = area[3:] - area[:3]
        dl = a.query((area[:3] + lo * size).astype(np.float32))
        dh = a.query((area[:3] + hi * size).astype(np.float32))
        jumps[alg] = float(np.abs(dl - dh).max())
    assert jumps[2] < 2e-4, jumps
    assert jumps[2] < 0.25 * jumps[1], jumps


# ---- port vs the unmodified reference (only where oracle/_ref exists) -------------------------------
@pytest.mark.parametrize("subdiv,depth,start,rule", [(2, 5, 3, 1), (2, 5, 2, 1), (3, 6, 3, 1), (2, 4, 0, 1), (3, 5, 3, 2), (3, 5, 3, 3)])
def test_port_continuity_equals_reference(port, ref, subdiv, depth, start, rule, tmp_path):
    v, i = displaced_sphere(subdiv)
    from sdflib_b200 import meshes
    box = meshes.bounding_box_with_margin(v)
    a = ref.build_octree(v, i, box, depth, start, 1e-3, 2, 1, termination_rule=rule, param1=0.1)
    b = port.build_octree(v, i, box, depth, start, 1e-3, 2, 1, termination_rule=rule, param1=0.1, use_cache=True)
    assert np.array_equal(a.octree_data(), b.octree_data())
    pa, pb = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    a.save(pa); b.save(pb)
    ba, bb = bytearray(open(pa, "rb").read()), bytearray(open(pb, "rb").read())
    # mValueRange is never initialised on the reference's CONTINUITY path (OctreeSdf.h:251 — max() accumulates
    # onto heap garbage, OctreeSdfBreadthFirstNoDelay.h:728): bytes 37..40 of the file are excluded.
    ba[37:41] = bb[37:41] = b"\0\0\0\0"
    assert ba == bb
    area = a.sample_area()
    q = (area[:3] + np.random.default_rng(5).uniform(-0.1, 1.1, (20000, 3)) * (area[3:] - area[:3])).astype(np.float32)
    (da, ga), (db, gb) = a.query(q, True), b.query(q, True)
    assert_bit_equal(da, db); assert_bit_equal(ga, gb)


@pytest.mark.parametrize("subdiv,depth,start,threads", [(2, 5, 3, 1), (3, 5, 2, 1), (3, 6, 3, 1)])
def test_port_octree_equals_reference(port, ref, subdiv, depth, start, threads, tmp_path):
    v, i = displaced_sphere(subdiv)
    from sdflib_b200 import meshes
    box = meshes.bounding_box_with_margin(v)
    a = ref.build_octree(v, i, box, depth, start, 1e-3, 1, threads)
    b = port.build_octree(v, i, box, depth, start, 1e-3, 1, threads, use_cache=True)
    assert np.array_equal(a.octree_data(), b.octree_data())
    pa, pb = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    a.save(pa); b.save(pb)
    assert open(pa, "rb").read() == open(pb, "rb").read()
    area = a.sample_area()
    q = (area[:3] + np.random.default_rng(5).uniform(-0.1, 1.1, (20000, 3)) * (area[3:] - area[:3])).astype(np.float32)
    (da, ga), (db, gb) = a.query(q, True), b.query(q, True)
    assert_bit_equal(da, db); assert_bit_equal(ga, gb)


@pytest.mark.parametrize("subdiv,depth,start,min_tris", [(2, 5, 3, 8), (3, 5, 1, 16), (3, 5, 2, 32)])
def test_port_exact_equals_reference(port, ref, subdiv, depth, start, min_tris, tmp_path):
    v, i = displaced_sphere(subdiv)
    from sdflib_b200 import meshes
    box = meshes.bounding_box_with_margin(v)
    a = ref.build_exact(v, i, box, depth, start, min_tris, 1)
    b = port.build_exact(v, i, box, depth, start, min_tris, 1, use_cache=True)
    pa, pb = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    a.save(pa); b.save(pb)
    assert open(pa, "rb").read() == open(pb, "rb").read()
    area = a.sample_area()
    q = (area[:3] + np.random.default_rng(6).uniform(-0.1, 1.1, (5000, 3)) * (area[3:] - area[:3])).astype(np.float32)
    (da, ga), (db, gb) = a.query(q, True), b.query(q, True)
    assert_bit_equal(da, db); assert_bit_equal(ga, gb)


def test_port_kernels_equal_reference(port, ref):
    rng = np.random.default_rng(11)
    v, i = displaced_sphere(3)
    ta, tb = ref.triangle_data(v, i), port.triangle_data(v, i)
    assert_bit_equal(ta, tb, "TriangleData")
    pts = rng.uniform(-2, 2, (20000, 3)).astype(np.float32)
    assert np.array_equal(ref.nearest_triangle(v, i, pts[:4000]), port.nearest_triangle(v, i, pts[:4000]))
    for t in (0, 99, 1000):
        w = v[i[3 * t:3 * t + 3]].reshape(-1)
        assert_bit_equal(ref.sq_dist(ta[t], pts), port.sq_dist(ta[t], pts))
        for mode in (0, 1, 2):
            (da, ga), (db, gb) = ref.signed_dist(ta[t], w, pts, mode), port.signed_dist(ta[t], w, pts, mode)
            assert_bit_equal(da, db); assert_bit_equal(ga, gb)
    for _ in range(200):
        half = float(rng.uniform(0.05, 0.5))
        r = rng.uniform(0, 0.3, 8).astype(np.float32); r[rng.integers(8)] = 0
        tri = rng.uniform(-1, 1, 9).astype(np.float32)
        thr = float(rng.uniform(0, 0.5))
        assert ref.is_near_minimize(half, r, tri, thr) == port.is_near_minimize(half, r, tri, thr)


@pytest.mark.parametrize("name", ["tetrahedron", "single_triangle", "two_spheres"])
def test_port_equals_reference_on_edge_case_meshes(port, ref, name):
    """Open / tiny / disconnected inputs, start depth equal to the depth, start depth 0 — all three builders."""
    from sdflib_b200 import meshes
    v, i = edge_case_meshes()[name]
    box = meshes.bounding_box_with_margin(v)
    for alg in (1, 2):
        for depth, start in ((4, 2), (3, 3), (4, 0)):
            a = ref.build_octree(v, i, box, depth, start, 1e-3, alg, 1).octree_data()
            b = port.build_octree(v, i, box, depth, start, 1e-3, alg, 1, use_cache=True).octree_data()
            assert np.array_equal(a, b), (name, alg, depth, start)
    if name == "single_triangle":
        return   # ExactOctreeSdf packs indices in ceil(log2(#triangles)) = 0 bits for one triangle: undefined in the reference
    a, b = ref.build_exact(v, i, box, 4, 1, 4, 1), port.build_exact(v, i, box, 4, 1, 4, 1, use_cache=True)
    assert np.array_equal(a.octree_data(), b.octree_data())
    q = (box[:3] + np.random.default_rng(3).uniform(0, 1, (2000, 3)) * (box[3:] - box[:3])).astype(np.float32)
    assert_bit_equal(a.query(q), b.query(q))


def test_port_builds_equal_reference_on_random_non_manifold_meshes(port, ref):
    """Random triangle soups, spheres with duplicated vertices or holes: every builder, random depths / start depths /
    rules / thresholds (a longer run of the same loop — 80 meshes — found no difference either)."""
    from test_capi_host import _random_meshes
    rng = np.random.default_rng(21)
    for n, (v, i) in enumerate(_random_meshes(rng, 16)):
        lo, hi = v.min(0), v.max(0)
        m = 0.2 * float((hi - lo).max())
        box = np.concatenate([lo - m, hi + m]).astype(np.float32)
        depth = int(rng.integers(3, 5)); start = int(rng.integers(0, 4)); thr = float(rng.choice([1e-2, 1e-1])); rule = int(rng.integers(1, 4))
        for alg in (1, 2):
            a = ref.build_octree(v, i, box, depth, start, thr, alg, 1, termination_rule=rule, param1=0.1).octree_data()
            b = port.build_octree(v, i, box, depth, start, thr, alg, 1, termination_rule=rule, param1=0.1, use_cache=True).octree_data()
            assert np.array_equal(a, b), (n, alg, depth, start, rule, thr)
        if i.size // 3 >= 2:
            start = int(rng.integers(0, 3)); depth = start + int(rng.integers(2, 4)); min_tris = int(rng.choice([1, 4, 16, 64]))
            ea, eb = ref.build_exact(v, i, box, depth, start, min_tris, 1), port.build_exact(v, i, box, depth, start, min_tris, 1, use_cache=True)
            assert np.array_equal(ea.octree_data(), eb.octree_data()), (n, depth, start, min_tris)
            q = (box[:3] + rng.random((300, 3)) * (box[3:] - box[:3])).astype(np.float32)
            assert_bit_equal(ea.query(q), eb.query(q))
