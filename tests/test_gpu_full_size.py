"""Parity AT THE BASELINE SIZES: BASELINE configs 2 and 3 built on the GPU at their stated
sizes and compared with tests/golden/full_size.npz, which tests/golden/make_golden_full.py wrote from the compiled,
unmodified reference and from the history-free oracle (both on the CPU, single thread).

  * arrays, header scalars and .bin bytes: BIT-EXACT against the history-free oracle (sha256 of every array);
  * OctreeSdf NO_CONTINUITY: node words (topology + indices) and c0 of every leaf additionally bit-exact against the
    REFERENCE's own single-thread build; CONTINUITY and ExactOctreeSdf differ from it in 28 of 703 669 leaves / 152 of
    1 063 720 nodes, because the reference's 32^3 vertex cache makes tie-broken samples depend on traversal history
     — there the reference comparison is on distances;
  * getDistance on a strided sample of the 256^3 grid (every 61st point, 275 037 points): exact-order kernel bit-exact
    against the oracle's query of the same array; against the REFERENCE's distances (its own build) the north-star gate
    |d - ref| <= 1e-5 max(|ref|, 1e-3 box) with the fraction of tie-affected points below 1e-4; ExactOctreeSdf distances
    and gradients bit-exact against the reference;
  * the GPU-built .bin loaded by the compiled reference (when oracle/_ref travelled to this box) answers the sample with
    the GPU exact-order kernel's bits;
  * kernel-level known answers (tests/golden/kernels.npz, all seven regions of the reference's TriangleDistanceTest
    triangle) through sdfb200_point_triangle / sdfb200_nearest_triangle ON THE GPU.
"""
import ctypes as C
import hashlib

import numpy as np
import pytest

from conftest import assert_bit_equal, golden, displaced_sphere

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def file_sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def topology_mask(words, g3):
    topo = np.zeros(words.size, bool)
    topo[:g3] = True
    level = np.arange(g3, dtype=np.int64)
    while level.size:
        w = words[level]
        base = (w[(w & 0x80000000) == 0] & 0x3FFFFFFF).astype(np.int64)
        level = (base[:, None] + np.arange(8)).reshape(-1)
        topo[level] = True
    return topo


def grid_sample(area, stride, n=256):
    g = (np.arange(n, dtype=np.float32) + np.float32(0.5)) / np.float32(n)
    idx = np.arange(0, n ** 3, stride, dtype=np.int64)
    p = np.stack([g[idx % n], g[(idx // n) % n], g[idx // (n * n)]], -1)
    return (area[:3] + p * (area[3:] - area[:3])).astype(np.float32)


@pytest.fixture(scope="module")
def m1(sdf):
    gold = golden("full_size.npz")
    v, i = sdf.meshes.config_mesh("M1")
    assert sha(v) + sha(i) == str(gold["mesh_sha256"]), "config mesh M1 differs from the one the fixture was generated on"
    box = sdf.meshes.bounding_box_with_margin(v)
    assert_bit_equal(box, gold["box"], "box")
    return gold, v, i, sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])


def reference_gate(d, ref_d, size):
    """North-star tolerance against the reference's own build: ties between triangles (and, for CONTINUITY, the handful
    of leaves whose subdivision decision they flip) may move a point by more, their fraction must stay below 1e-4."""
    tol = 1e-5 * np.maximum(np.abs(ref_d), 1e-3 * size)
    over = np.abs(d - ref_d) > tol
    return float(over.mean()), float(np.abs(d - ref_d).max())


@pytest.mark.parametrize("name,algorithm", [("c2_nocont", 1), ("c2_cont", 2)])
def test_config2_octree_at_full_size(sdf, ref, m1, tmp_path, name, algorithm):
    gold, v, i, mesh, bb = m1
    s = sdf.OctreeSdf(mesh, bb, 8, 3, 1e-3, algorithm, 1)
    d = s.getOctreeData()
    assert d.size == int(gold[name + "_words"])
    assert sha(d) == str(gold[name + "_sha256"]), "octree words differ from the history-free oracle at full size"
    info = s.info()
    assert_bit_equal(np.float32([info.value_range, info.min_border_value]), np.float32([gold[name + "_value_range"], gold[name + "_min_border_value"]]))
    path = str(tmp_path / "gpu.bin")
    assert s.saveToFile(path)
    assert file_sha(path) == str(gold[name + "_bin_sha256"]), ".bin bytes differ from the oracle's"
    if bool(gold[name + "_topology_equals_reference"]):
        topo = topology_mask(d, 512)
        assert sha(d[topo]) == str(gold[name + "_ref_topology_sha256"]), "node words differ from the REFERENCE's single-thread build"
        assert sha(d[~topo].reshape(-1, 64)[:, 0]) == str(gold[name + "_ref_c0_sha256"]), "c0 of the leaves differs from the reference"
    else:
        leaves = int((~topology_mask(d, 512)).sum() // 64)
        assert leaves == int(gold[name + "_leaves"]) and abs(leaves - int(gold[name + "_ref_leaves"])) <= 1e-4 * leaves
    area = s.getSampleArea().as_array()
    assert_bit_equal(area, gold["sample_area"], "sample area")
    size = float(area[3] - area[0])
    q = grid_sample(area, int(gold["stride"]))
    every = int(gold["grad_every"])
    dist, grad = s.getDistance(q, gradient=True, exact_order=True)
    assert sha(dist) == str(gold[name + "_port_distances_sha256"]), "exact-order distances differ from the oracle's query"
    assert sha(grad[::every]) == str(gold[name + "_port_gradients_sha256"]), "exact-order gradients differ from the oracle's query"
    # against the reference's own build + query
    frac, worst = reference_gate(dist, gold[name + "_distances"], size)
    assert frac < 1e-4 and worst < 2e-3 * size, (frac, worst)
    fast, fast_grad = s.getDistance(q, gradient=True)
    frac, worst = reference_gate(fast, gold[name + "_distances"], size)
    assert frac < 2e-4 and worst < 2e-3 * size, (frac, worst)
    ok = np.isfinite(gold[name + "_gradients"]).all(1) & np.isfinite(fast_grad[::every]).all(1)
    assert ok.mean() > 0.99
    assert (np.abs(fast_grad[::every][ok] - gold[name + "_gradients"][ok]) > 1e-3).mean() < 2e-3
    # the compiled reference reads the GPU's file and gives the GPU's bits
    r = ref.load(path)
    rd, rg = r.query(q, True, 8)
    assert_bit_equal(rd, dist, "reference query of the GPU-built .bin")
    assert_bit_equal(rg[::every], grad[::every], "reference gradients of the GPU-built .bin")


def test_config3_exact_octree_at_full_size(sdf, ref, m1, tmp_path):
    gold, v, i, mesh, bb = m1
    s = sdf.ExactOctreeSdf(mesh, bb, 7, 3, 128, 1)
    nodes, sets, masks = s.getOctreeData(), s.getTrianglesSets(), s.getTrianglesMasks()
    assert nodes.shape[0] == int(gold["c3_nodes"]) and sets.size == int(gold["c3_sets_words"]) and masks.size == int(gold["c3_masks_bytes"])
    assert sha(nodes) == str(gold["c3_nodes_sha256"])
    assert sha(sets) == str(gold["c3_sets_sha256"])
    assert sha(masks) == str(gold["c3_masks_sha256"])
    assert sha(s.getTrianglesData()) == str(gold["c3_triangle_data_sha256"])
    path = str(tmp_path / "gpu_exact.bin")
    assert s.saveToFile(path)
    assert file_sha(path) == str(gold["c3_bin_sha256"])
    q = grid_sample(s.getSampleArea().as_array(), int(gold["stride"]))
    dist, grad = s.getDistance(q, gradient=True)
    assert_bit_equal(dist, gold["c3_distances"], "distances against the REFERENCE's own build")       # layout- and history-independent
    assert_bit_equal(grad[::int(gold["grad_every"])], gold["c3_gradients"], "gradients against the REFERENCE's own build")
    r = ref.load(path)
    assert_bit_equal(r.query(q[::16], False, 8), dist[::16], "reference query of the GPU-built .bin")


def test_config4_dragon_class_exact_octree_at_full_size(sdf):
    """BASELINE configs[3]: 5 242 880 triangles, ExactOctreeSdf depth 8 / start depth 3 / minTrianglesPerNode 128 — 11.3 M nodes,
    162 M set words, 812 MB of masks, more than 2^32 (node, triangle) pairs on the two deepest levels. Hashes and a 256^3
    distance sample from the history-free CPU oracle (tests/golden/make_golden_c4.py, about an hour of CPU)."""
    gold = golden("config4.npz")
    v, i = sdf.meshes.config_mesh("M2")
    assert sha(v) + sha(i) == str(gold["mesh_sha256"]), "config mesh M2 differs from the one the fixture was generated on"
    box = gold["box"]
    s = sdf.ExactOctreeSdf(sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:]), 8, 3, 128, 1)
    try:
        nodes, sets, masks = s.getOctreeData(), s.getTrianglesSets(), s.getTrianglesMasks()
        assert nodes.shape[0] == int(gold["nodes"]) and sets.size == int(gold["sets_words"]) and masks.size == int(gold["masks_bytes"])
        assert sha(nodes) == str(gold["nodes_sha256"])
        assert sha(sets) == str(gold["sets_sha256"])
        assert sha(masks) == str(gold["masks_sha256"])
        assert sha(s.getTrianglesData()) == str(gold["triangle_data_sha256"])
        q = grid_sample(s.getSampleArea().as_array(), 61)[::int(gold["sample_every"])]   # make_golden_full.grid_sample (STRIDE = 61), then every 16th
        dist, grad = s.getDistance(q, gradient=True)
        assert_bit_equal(dist, gold["distances"], "distances against the oracle's build")
        assert_bit_equal(grad, gold["gradients"], "gradients against the oracle's build")
    finally:
        s.close()
        sdf.lib().sdfb200_release_cached_memory()   # ~95 GB of level arrays: back to the driver before the next test


def test_exact_octree_against_reference_build(sdf, ref):
    """GPU build <-> the compiled reference's own build (not the port): same distances and gradients bit for bit; the
    node arrays agree except where the reference's vertex cache changed a tie (counted)."""
    v, i = displaced_sphere(4)
    box = sdf.meshes.bounding_box_with_margin(v)
    g = sdf.ExactOctreeSdf(sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:]), 6, 3, 32, 1)
    r = ref.build_exact(v, i, box, 6, 3, 32, 1)
    rng = np.random.default_rng(5)
    area = g.getSampleArea().as_array()
    q = (area[:3] + rng.uniform(-0.05, 1.05, (200000, 3)) * (area[3:] - area[:3])).astype(np.float32)
    (d, gr), (rd, rg) = g.getDistance(q, gradient=True), r.query(q, True, 8)
    assert_bit_equal(d, rd, "distance")
    inside = ~((q < area[:3]) | (q >= area[3:])).any(1)     # the reference leaves out-of-box gradients uninitialised
    assert_bit_equal(gr[inside], rg[inside], "gradient")
    gn, rn = g.getOctreeData(), r.octree_data().reshape(-1, 2)
    assert abs(gn.shape[0] - rn.shape[0]) <= 1e-3 * rn.shape[0]
    if gn.shape == rn.shape:
        assert (gn[:, 0] != rn[:, 0]).mean() < 1e-3


# ---- kernel-level known answers on the GPU (reference: src/tools/TriangleDistanceTest/main.cpp:12-64) ----------------------
def test_point_triangle_kernel_known_answers(sdf):
    k = golden("kernels.npz")
    L = sdf.lib()
    tri, pts = np.ascontiguousarray(k["tet_triangle_data"][0]), np.ascontiguousarray(k["points"])
    w = np.ascontiguousarray(k["tet_vertices"][k["tet_indices"][:3]].reshape(-1))
    n = len(pts)
    p = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    for mode, dist_key, grad_key in ((0, "signed0", None), (1, "signed1", "grad1"), (2, "signed2", "grad2")):
        d = np.empty(n, np.float32)
        g = np.zeros((n, 3), np.float32)
        assert L.sdfb200_point_triangle(p(tri), p(w), p(pts), C.c_uint64(n), C.c_int(mode), p(d), p(g) if grad_key else None) == 0, L.sdfb200_last_error()
        assert_bit_equal(d, k[dist_key], f"signed distance, mode {mode}")
        if grad_key:
            assert_bit_equal(g, k[grad_key], f"gradient, mode {mode}")
    # squared distance (a3) is |signed distance|^2's source: all seven regions of the triangle frame must occur in the vectors
    sq = k["sq_dist"]
    assert np.allclose(np.sqrt(sq), np.abs(k["signed0"]), rtol=2e-6, atol=1e-7)
    td = tri
    origin, tr = td[:3], td[3:12].reshape(3, 3)
    local = (pts - origin) @ tr   # rough region census in the triangle frame: enough to see every branch exercised
    assert len(np.unique(np.sign(local[:, :2]).astype(int), axis=0)) >= 4


def test_nearest_triangle_kernel_known_answers(sdf):
    m = golden("mesh_small.npz")
    L = sdf.lib()
    v, i, q = np.ascontiguousarray(m["vertices"]), np.ascontiguousarray(m["indices"]), np.ascontiguousarray(m["query_points"])
    out = np.empty(len(q), np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    assert L.sdfb200_nearest_triangle(p(v), C.c_uint32(len(v)), p(i), C.c_uint32(i.size), p(q), C.c_uint64(len(q)), p(out)) == 0, L.sdfb200_last_error()
    assert np.array_equal(out, m["nearest"]), "BVH traversal on the GPU picks other triangles than tmd::TriangleMeshDistance"
    td = np.empty((i.size // 3, 37), np.float32)
    assert L.sdfb200_triangle_data(p(v), C.c_uint32(len(v)), p(i), C.c_uint32(i.size), p(td)) == 0
    assert_bit_equal(td, m["triangle_data"], "TriangleData")
