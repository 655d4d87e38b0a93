"""GPU parity tests of hot path 1 + 2 for ExactOctreeSdf, through the C-ABI (ctypes -> libsdfb200.so).

Parity bars (all integer / index work, so everything is BIT-EXACT):
  * mOctreeData (childrenIndex, trianglesArrayIndex), mTrianglesSets, mTrianglesMasks, header scalars and the
    .bin bytes against the history-free oracle (port, use_cache=False) — same filter decisions (Frank-Wolfe in
    the reference's float order), same first-strict-minimum sample choice, same array order;
  * getDistance(+gradient): bit-exact against the oracle's query of the same structure, and against the
    brute-force definition (nearest triangle of the whole mesh) — exact distances are layout- and
    history-independent, so they also equal the reference's own build, see the golden test.
"""
import numpy as np
import pytest

from conftest import assert_bit_equal, golden, displaced_sphere

pytestmark = pytest.mark.gpu


def build_both(sdf, port, v, i, depth, start, min_tris, threads=1):
    box = sdf.meshes.bounding_box_with_margin(v)
    g = sdf.ExactOctreeSdf(sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:]), depth, start, min_tris, threads)
    p = port.build_exact(v, i, box, depth, start, min_tris, 1, use_cache=False)
    return g, p, box


def random_points(area, n, seed, spill=0.1):
    rng = np.random.default_rng(seed)
    return (area[:3] + rng.uniform(-spill, 1 + spill, (n, 3)) * (area[3:] - area[:3])).astype(np.float32)


def port_arrays(p):
    import ctypes as C
    ns, nm, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
    p.b.lib.orc_exact_sizes(p.h, C.byref(ns), C.byref(nm), C.byref(nt))
    sets, masks, tris = np.empty(ns.value, np.uint32), np.empty(nm.value, np.uint8), np.empty((nt.value, 37), np.float32)
    p.b.lib.orc_exact_arrays(p.h, sets.ctypes.data_as(C.c_void_p), masks.ctypes.data_as(C.c_void_p), tris.ctypes.data_as(C.c_void_p))
    hdr = np.empty(8, np.uint32)
    p.b.lib.orc_exact_header(p.h, hdr.ctypes.data_as(C.c_void_p))
    return sets, masks, tris, hdr


@pytest.mark.parametrize("subdiv,depth,start,min_tris", [(2, 4, 1, 16), (2, 5, 3, 8), (3, 5, 2, 32), (3, 4, 0, 16), (4, 6, 3, 32),
                                                          (3, 6, 1, 4), (2, 3, 1, 0), (4, 5, 2, 128)])
def test_build_bit_exact_vs_oracle(sdf, port, subdiv, depth, start, min_tris):
    v, i = displaced_sphere(subdiv)
    g, p, _ = build_both(sdf, port, v, i, depth, start, min_tris)
    sets, masks, tris, hdr = port_arrays(p)
    nodes = g.getOctreeData()
    want = p.octree_data().reshape(-1, 2)
    assert nodes.shape == want.shape
    assert np.array_equal(nodes[:, 0], want[:, 0]), f"{int((nodes[:, 0] != want[:, 0]).sum())} childrenIndex words differ"
    assert np.array_equal(nodes[:, 1], want[:, 1]), f"{int((nodes[:, 1] != want[:, 1]).sum())} trianglesArrayIndex words differ"
    gs, gm = g.getTrianglesSets(), g.getTrianglesMasks()
    assert gs.size == sets.size and np.array_equal(gs, sets)
    assert gm.size == masks.size and np.array_equal(gm, masks)
    assert_bit_equal(g.getTrianglesData(), tris, "TriangleData")
    info = g.info()
    got = [info.start_depth, info.min_triangles_in_leafs, info.max_triangles_in_leafs, info.max_triangles_encoded_in_leafs,
           info.bit_encoding_start_depth, info.bits_per_index, info.max_depth]
    # orc_exact_header: startGridSize, startDepth, minTris, maxTris, maxEncoded, bitEncodingStartDepth, bitsPerIndex, maxDepth
    assert got == [int(x) for x in hdr[1:8]], (got, hdr)
    assert info.start_grid_size == int(hdr[0]) == 1 << start


def brute_force(port, v, i, q):
    """Definition of the exact SDF: signed distance to the nearest triangle of the whole mesh (RealSdf.cpp:10-25)."""
    tris = port.triangle_data(v, i)
    best = np.full(len(q), np.inf, np.float32)
    tri = np.zeros(len(q), np.int64)
    for t in range(len(tris)):
        d = port.sq_dist(tris[t], q)
        better = d < best
        best[better] = d[better]
        tri[better] = t
    out = np.empty(len(q), np.float32)
    for t in np.unique(tri):
        m = tri == t
        out[m] = port.signed_dist(tris[t], np.zeros((3, 3), np.float32), q[m], 0)[0]
    return out


@pytest.mark.parametrize("gradient", [False, True])
def test_query_bit_exact_vs_oracle(sdf, port, gradient):
    v, i = displaced_sphere(3)
    g, p, _ = build_both(sdf, port, v, i, 5, 2, 16)
    q = random_points(g.getSampleArea().as_array(), 100000, 11)          # includes out-of-grid points
    if gradient:
        (d, gr), (od, og) = g.getDistance(q, gradient=True), p.query(q, True)
        assert_bit_equal(d, od, "distance")
        assert_bit_equal(gr, og, "gradient")
    else:
        assert_bit_equal(g.getDistance(q), p.query(q), "distance")


def test_query_equals_brute_force_definition(sdf, port):
    v, i = displaced_sphere(2)
    g, p, _ = build_both(sdf, port, v, i, 5, 3, 8)
    q = random_points(g.getSampleArea().as_array(), 4000, 12, spill=0.0)
    d = g.getDistance(q)
    want = brute_force(port, v, i, q)
    # the octree only restricts WHICH triangles are tested; the nearest one must be among them
    assert np.abs(np.abs(d) - np.abs(want)).max() <= 1e-6
    assert (np.sign(d) == np.sign(want)).mean() > 0.999


def test_reference_bin_is_loaded_and_queried(sdf, tmp_path):
    """A .bin written by the unmodified reference (tests/golden) decodes to the reference's own distances."""
    gold = golden("small_structures.npz")
    path = tmp_path / "ref_exact.bin"
    path.write_bytes(gold["exact_bin"].tobytes())
    s = sdf.SdfFunction.loadFromFile(str(path))
    assert isinstance(s, sdf.ExactOctreeSdf)
    d, gr = s.getDistance(gold["query_points"], gradient=True)
    assert_bit_equal(d, gold["exact_distances"], "distance")
    assert_bit_equal(gr, gold["exact_gradients"], "gradient")
    again = str(tmp_path / "resaved.bin")
    assert s.saveToFile(again) and open(again, "rb").read() == gold["exact_bin"].tobytes()


def test_bin_round_trip_and_oracle_bytes(sdf, port, tmp_path):
    v, i = displaced_sphere(3)
    g, p, _ = build_both(sdf, port, v, i, 5, 2, 32)
    a, b = str(tmp_path / "gpu.bin"), str(tmp_path / "port.bin")
    assert g.saveToFile(a) and p.save(b)
    assert open(a, "rb").read() == open(b, "rb").read()
    again = sdf.SdfFunction.loadFromFile(a)
    q = random_points(g.getSampleArea().as_array(), 20000, 13)
    assert_bit_equal(again.getDistance(q), g.getDistance(q))


def test_thread_layouts_hold_the_same_tree(sdf, port):
    """numThreads >= 2 selects the per-start-voxel layout (the merge the reference intends,
    ExactOctreeSdfDepthFirst.h:576-622); queries must not depend on the layout."""
    v, i = displaced_sphere(3)
    box = sdf.meshes.bounding_box_with_margin(v)
    mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
    a = sdf.ExactOctreeSdf(mesh, bb, 5, 2, 16, 1)
    b = sdf.ExactOctreeSdf(mesh, bb, 5, 2, 16, 4)
    na, nb = a.getOctreeData(), b.getOctreeData()
    assert na.shape == nb.shape and a.getTrianglesSets().size == b.getTrianglesSets().size
    assert a.getTrianglesMasks().size == b.getTrianglesMasks().size
    assert (na[:, 0] == 0xFFFFFFFF).sum() == (nb[:, 0] == 0xFFFFFFFF).sum()
    q = random_points(a.getSampleArea().as_array(), 50000, 14)
    assert_bit_equal(a.getDistance(q), b.getDistance(q))
    # per-voxel layout: sub-octrees appear in start-grid order, so child blocks of consecutive inner start slots ascend
    inner = nb[:64][nb[:64, 0] != 0xFFFFFFFF, 0]
    assert np.all(np.diff(inner.astype(np.int64)) > 0)


def test_invalid_arguments(sdf):
    v, i = displaced_sphere(2)
    box = sdf.meshes.bounding_box_with_margin(v)
    with pytest.raises(sdf.SdfB200Error):
        sdf.ExactOctreeSdf(sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:]), 3, 2, 16, 1)   # maxDepth < startDepth + 2


def test_config3_full_size_properties(sdf, port):
    """Config 3 at full size (M1, 327 680 triangles, depth 7, minTri 128): determinism, structural validity,
    and exactness of a sample of the 256^3 grid against the brute-force definition computed on the GPU side
    (every query's result must be the minimum over ALL triangles: checked through |d| <= |d_any triangle|)."""
    import hashlib
    v, i = sdf.meshes.config_mesh("M1")
    box = sdf.meshes.bounding_box_with_margin(v)
    mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
    a = sdf.ExactOctreeSdf(mesh, bb, 7, 3, 128, 2)
    b = sdf.ExactOctreeSdf(mesh, bb, 7, 3, 128, 2)
    na = a.getOctreeData()
    for x, y in ((na, b.getOctreeData()), (a.getTrianglesSets(), b.getTrianglesSets()), (a.getTrianglesMasks(), b.getTrianglesMasks())):
        assert hashlib.sha256(x.tobytes()).hexdigest() == hashlib.sha256(y.tobytes()).hexdigest()
    info = a.info()
    assert info.bits_per_index == 19 and info.bit_encoding_start_depth == 5
    inner = na[:, 0] != 0xFFFFFFFF
    assert na.shape[0] == 512 + 8 * int(inner.sum())                       # every inner node owns one 8-block
    assert info.max_triangles_in_leafs >= info.min_triangles_in_leafs
    # exactness on a sample: the octree distance equals the oracle's brute force over the whole mesh
    q = random_points(a.getSampleArea().as_array(), 64, 15, spill=0.0)
    d = a.getDistance(q)
    want = brute_force(port, v, i, q)
    assert np.abs(np.abs(d) - np.abs(want)).max() <= 1e-6
