import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def bits(a):
    """Bit pattern view for exact float comparisons (distinguishes -0.0 / +0.0, compares NaN payloads)."""
    return np.ascontiguousarray(a).view(np.uint32)


def assert_bit_equal(a, b, what=""):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    same = bits(a) == bits(b)
    assert same.all(), f"{what}: {int((~same).sum())} of {same.size} values differ (max abs {np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64)))})"


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def port():
    from oracle.binding import port as p
    if not p.available():
        import subprocess
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    return p


@pytest.fixture(scope="session")
def ref():
    from oracle.binding import ref as r
    if not r.available():
        pytest.skip("oracle/_ref/libsdfref.so not built (needs the reference tree)")
    return r


@pytest.fixture(scope="session")
def sdf():
    """The product package; building the CUDA library is part of the session set-up."""
    import sdflib_b200
    if not os.path.exists(sdflib_b200.LIB_PATH):
        from sdflib_b200 import build
        build.build(verbose=False)
    return sdflib_b200


def displaced_sphere(subdiv):
    """Generic-position test mesh: reference icosphere + the closed-form displacement."""
    from sdflib_b200 import meshes
    v, i = meshes.isosphere(subdiv)
    return meshes.displace(v), i


def octree_topology(words, start_grid):
    """Walks an OctreeSdf array: returns (mask of node words, #leaves, #inner)."""
    words = np.asarray(words)
    topo = np.zeros(words.size, bool)
    g3 = start_grid ** 3
    topo[:g3] = True
    stack = list(range(g3))
    leaves = inner = 0
    while stack:
        s = stack.pop()
        w = int(words[s])
        if w & 0x80000000:
            leaves += 1
        else:
            inner += 1
            c = w & 0x3FFFFFFF
            topo[c:c + 8] = True
            stack.extend(range(c, c + 8))
    return topo, leaves, inner


def edge_case_meshes():
    """Small meshes off the beaten path: a closed tetrahedron, a single (open) triangle, two disjoint spheres."""
    from sdflib_b200 import meshes
    tet_v = np.float32([[-0.5, -0.5, 0.0], [0.5, -0.5, 0.0], [0.0, 0.5, 0.0], [0.1, 0.0, 0.7]])
    tet_i = np.uint32([0, 1, 2, 1, 0, 3, 2, 1, 3, 0, 2, 3])
    tri_v = np.float32([[-0.5, -0.4, 0.1], [0.6, -0.5, 0.0], [0.05, 0.5, -0.1]])
    tri_i = np.uint32([0, 1, 2])
    sv, si = meshes.isosphere(1)
    two_v = np.concatenate([sv * np.float32(0.4) + np.float32([-0.6, 0, 0]), sv * np.float32(0.3) + np.float32([0.7, 0.1, 0])]).astype(np.float32)
    two_i = np.concatenate([si, si + len(sv)]).astype(np.uint32)
    return {"tetrahedron": (tet_v, tet_i), "single_triangle": (tri_v, tri_i), "two_spheres": (two_v, two_i)}
