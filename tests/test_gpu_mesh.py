"""Mesh ingestion on the device (mesh_device.cu) against the host restatement and the oracle:
TriangleData must come out bit for bit — constructor, edge pseudo-normals (which corners pair up), angle-weighted vertex
pseudo-normals (acosf bits and addition order), the non-manifold repair — on closed meshes, on random triangle soups with
duplicated vertices / triangles / holes, and on the benchmark mesh at full size (sha256 against the fixture written
from the reference)."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from conftest import assert_bit_equal, golden, displaced_sphere, edge_case_meshes

pytestmark = pytest.mark.gpu


def device_triangle_data(sdf, v, i):
    """TriangleData as the builders compute it: an ExactOctreeSdf build exposes the mesh's array through its getter."""
    box = sdf.meshes.bounding_box_with_margin(v)
    s = sdf.ExactOctreeSdf(sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:]), 3, 1, 8, 1)
    return s.getTrianglesData()


def host_triangle_data(sdf, v, i):
    out = np.empty((i.size // 3, 37), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    v, i = np.ascontiguousarray(v, np.float32), np.ascontiguousarray(i, np.uint32)
    assert sdf.lib().sdfb200_triangle_data(p(v), C.c_uint32(len(v)), p(i), C.c_uint32(i.size), p(out)) == 0
    return out


@pytest.mark.parametrize("subdiv", [1, 3, 5])
def test_closed_meshes(sdf, port, subdiv):
    v, i = displaced_sphere(subdiv)
    got = device_triangle_data(sdf, v, i)
    assert_bit_equal(got, host_triangle_data(sdf, v, i), "device vs host TriangleData")
    assert_bit_equal(got, port.triangle_data(v, i), "device vs oracle TriangleData")


def test_edge_case_meshes(sdf, port):
    for name, (v, i) in edge_case_meshes().items():
        if i.size // 3 < 2:
            continue   # ExactOctreeSdf needs two triangles; the single triangle goes through the OctreeSdf tests
        assert_bit_equal(device_triangle_data(sdf, v, i), port.triangle_data(v, i), name)


def test_random_non_manifold_meshes(sdf, port):
    from test_capi_host import _random_meshes
    rng = np.random.default_rng(77)
    checked = 0
    for v, i in _random_meshes(rng, 60):
        if i.size // 3 < 2:
            continue
        got = device_triangle_data(sdf, v, i)
        want = port.triangle_data(v, i)
        same = got.view(np.uint32) == want.view(np.uint32)
        nan_both = np.isnan(got) & np.isnan(want)          # degenerate triangles: NaN payloads may differ between x86 and sm_100
        assert (same | nan_both).all(), f"mesh with {i.size // 3} triangles: {int((~(same | nan_both)).sum())} values differ"
        checked += 1
    assert checked > 40


def test_benchmark_mesh_at_full_size(sdf):
    gold = golden("full_size.npz")
    v, i = sdf.meshes.config_mesh("M1")
    got = device_triangle_data(sdf, v, i)
    assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == str(gold["c3_triangle_data_sha256"])
