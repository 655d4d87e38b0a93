"""The nearest-triangle BVH built on the device (bvh_device.cu / bvh_build.cuh; reference:
tmd::TriangleMeshDistance::_build_tree, TriangleMeshDistance.h:421-490) against the host builder of mesh_host.cpp, which
calls libstdc++'s own sort routines and was pinned against a serial restatement of the reference builder in round 1
(tests/cpp/bvh_host_main.cpp): every link and every float64 sphere must be identical, because the traversal order of this
tree decides between equidistant triangles. Covers tie-heavy meshes (the plain icosphere: all triangles that start at one
vertex tie on the sort key), generic ones, triangle counts that are not powers of two (ranges of unequal halves), random
soups, the one- and two-triangle trees, and the benchmark meshes at full size (327 680 and 5 242 880 triangles)."""
import numpy as np
import pytest

from conftest import displaced_sphere, edge_case_meshes

pytestmark = pytest.mark.gpu


def assert_same_tree(got, want, what):
    assert got.shape == want.shape, what
    assert np.array_equal(got["left"], want["left"]) and np.array_equal(got["right"], want["right"]), f"{what}: links differ"
    assert np.array_equal(got["leaf"] != 0, want["leaf"] != 0), f"{what}: leaf flags differ"
    inner = want["leaf"] == 0   # a leaf node's own sphere fields are never written (its sphere lives in the parent)
    for f in ("left_sphere", "right_sphere"):
        a, b = got[f][inner].view(np.uint64), want[f][inner].view(np.uint64)
        assert np.array_equal(a, b), f"{what}: {f} differs in {int((a != b).any(axis=1).sum())} of {int(inner.sum())} inner nodes"


def device_tree(sdf, v, i):
    pm = sdf.PreparedMesh(sdf.Mesh(v, i), bvh=True, exact=False)
    try:
        return pm.bvh_nodes(i.size // 3)
    finally:
        pm.close()


@pytest.mark.parametrize("subdiv,displaced", [(0, False), (1, True), (2, False), (3, True), (4, False), (5, True), (6, False)])
def test_icospheres(sdf, subdiv, displaced):
    v, i = displaced_sphere(subdiv) if displaced else sdf.meshes.isosphere(subdiv)
    assert_same_tree(device_tree(sdf, v, i), sdf.bvh_host(v, i), f"icosphere {subdiv}")


def test_odd_triangle_counts_and_soups(sdf):
    v, i = displaced_sphere(4)
    i = i.reshape(-1, 3)
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 5, 16, 17, 33, 100, 2047, 2048, 2049, 4097, 5119):
        keep = np.sort(rng.choice(len(i), n, replace=False))
        sub = np.ascontiguousarray(i[keep]).reshape(-1)
        assert_same_tree(device_tree(sdf, v, sub), sdf.bvh_host(v, sub), f"{n} triangles")
    from test_capi_host import _random_meshes
    checked = 0
    for vv, ii in _random_meshes(np.random.default_rng(78), 40):
        if ii.size // 3 < 1 or not np.isfinite(vv).all():
            continue
        assert_same_tree(device_tree(sdf, vv, ii), sdf.bvh_host(vv, ii), f"soup of {ii.size // 3}")
        checked += 1
    assert checked > 20
    for name, (vv, ii) in edge_case_meshes().items():
        assert_same_tree(device_tree(sdf, vv, ii), sdf.bvh_host(vv, ii), name)


def test_quantised_mesh_with_massive_ties(sdf):
    """Coordinates snapped to a coarse grid: thousands of equal keys per node, the case where an unstable sort's choices matter most."""
    v, i = displaced_sphere(6)
    v = (np.round(v * 16) / 16).astype(np.float32)
    assert_same_tree(device_tree(sdf, v, i), sdf.bvh_host(v, i), "quantised icosphere 6")


def test_coordinates_of_mixed_magnitude(sdf):
    """Centre sums are sequential float64 sums whose roundings depend on the running magnitude: a few coordinates ten orders of
    magnitude below the rest, meshes far from the origin and extreme scales, with the top levels' sums on the host threads
    (default) — both must give the host builder's bits."""
    v, i = displaced_sphere(6)
    cases = []
    for every in (7, 1000):
        w = v.copy()
        k = np.arange(0, len(w), every)
        w[k, k % 3] *= np.float32(1e-10)
        cases.append((f"tiny coordinate every {every} vertices", w))
    for offset, scale in ((123456.7, 1.0), (0.0, 1e-20), (0.0, 1e20), (-3.3, 1e-3)):
        cases.append((f"offset {offset} scale {scale}", (v * np.float32(scale) + np.float32(offset)).astype(np.float32)))
    for what, w in cases:
        assert_same_tree(device_tree(sdf, w, i), sdf.bvh_host(w, i), what)


@pytest.mark.parametrize("name", ["M1", "M2"])
def test_benchmark_meshes_at_full_size(sdf, name):
    v, i = sdf.meshes.config_mesh(name)
    assert_same_tree(device_tree(sdf, v, i), sdf.bvh_host(v, i), name)
