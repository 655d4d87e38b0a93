"""Prepared meshes and the single-process multi-device builds (include/sdfb200.h: sdfb200_mesh_*, sdfb200_build_*_multi):
every replica must be the single-device build bit for bit. On a box with one GPU the thread choreography (sizes summed
on the host, payload all-gather, segmented assembly, the CONTINUITY sample exchange) still runs with several "ranks" on
that GPU through peer copies (SDFB200_ALLOW_DUPLICATE_DEVICES=1); with two or more GPUs the same tests use distinct
devices and NCCL."""
import hashlib
import os

import numpy as np
import pytest

from conftest import displaced_sphere

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def device_lists(sdf):
    n = sdf.device_count()
    lists = [[0, 0], [0, 0, 0]]          # ranks sharing one device (needs the test switch)
    if n >= 2:
        lists.append(list(range(min(n, 4))))
    return lists


@pytest.fixture(autouse=True)
def allow_duplicates():
    os.environ["SDFB200_ALLOW_DUPLICATE_DEVICES"] = "1"
    yield
    os.environ.pop("SDFB200_ALLOW_DUPLICATE_DEVICES", None)


def test_prepared_mesh_serves_every_builder(sdf):
    v, i = displaced_sphere(4)
    box = sdf.meshes.bounding_box_with_margin(v)
    mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
    pm = sdf.PreparedMesh(mesh, bvh=True, exact=True)
    for alg in (sdf.OctreeSdf.NO_CONTINUITY, sdf.OctreeSdf.CONTINUITY):
        a, b = sdf.OctreeSdf(pm, bb, 6, 3, 1e-3, alg, 1), sdf.OctreeSdf(mesh, bb, 6, 3, 1e-3, alg, 1)
        assert sha(a.getOctreeData()) == sha(b.getOctreeData())
    a, b = sdf.ExactOctreeSdf(pm, bb, 6, 3, 32, 1), sdf.ExactOctreeSdf(mesh, bb, 6, 3, 32, 1)
    assert sha(a.getOctreeData()) == sha(b.getOctreeData()) and sha(a.getTrianglesSets()) == sha(b.getTrianglesSets())
    assert sha(a.getTrianglesData()) == sha(b.getTrianglesData())
    with pytest.raises(sdf.SdfB200Error):
        sdf.OctreeSdf(sdf.PreparedMesh(mesh, bvh=False, exact=True), bb, 5, 3)        # OctreeSdf needs the BVH part


def test_mesh_blob_round_trip(sdf):
    import torch
    v, i = displaced_sphere(3)
    box = sdf.meshes.bounding_box_with_margin(v)
    mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
    pm = sdf.PreparedMesh(mesh, bvh=True, exact=True)
    blob = torch.empty(pm.blob_bytes(), dtype=torch.uint8, device="cuda")
    pm.export_blob(blob.data_ptr(), blob.numel())
    clone = sdf.PreparedMesh.from_blob(blob.data_ptr(), blob.numel())
    assert sha(sdf.OctreeSdf(clone, bb, 5, 2).getOctreeData()) == sha(sdf.OctreeSdf(mesh, bb, 5, 2).getOctreeData())
    e, e0 = sdf.ExactOctreeSdf(clone, bb, 5, 2, 16), sdf.ExactOctreeSdf(mesh, bb, 5, 2, 16)
    assert sha(e.getOctreeData()) == sha(e0.getOctreeData()) and sha(e.getTrianglesMasks()) == sha(e0.getTrianglesMasks())
    with pytest.raises(sdf.SdfB200Error):
        sdf.PreparedMesh.from_blob(blob.data_ptr(), 64)                                  # truncated blob


@pytest.mark.parametrize("algorithm", [1, 2])
def test_octree_multi_device_equals_single(sdf, algorithm):
    v, i = displaced_sphere(4)
    box = sdf.meshes.bounding_box_with_margin(v)
    mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
    want = sdf.OctreeSdf(mesh, bb, 6, 3, 1e-3, algorithm, 2)
    rng = np.random.default_rng(3)
    q = (box[:3] + rng.random((50000, 3)) * (box[3:] - box[:3])).astype(np.float32)
    for devices in device_lists(sdf):
        replicas = sdf.OctreeSdf.build_on_devices(mesh, bb, 6, 3, devices, 1e-3, algorithm, 2)
        assert len(replicas) == len(devices)
        for k, r in enumerate(replicas):
            assert r.info().device == devices[k]
            assert sha(r.getOctreeData()) == sha(want.getOctreeData()), (devices, k)
            assert np.array_equal(r.getDistance(q, exact_order=True).view(np.uint32), want.getDistance(q, exact_order=True).view(np.uint32))
            assert r.info().min_border_value == want.info().min_border_value and r.info().value_range == want.info().value_range


def test_exact_multi_device_equals_single(sdf):
    v, i = displaced_sphere(4)
    box = sdf.meshes.bounding_box_with_margin(v)
    mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
    want = sdf.ExactOctreeSdf(mesh, bb, 6, 3, 32, 2)
    rng = np.random.default_rng(4)
    q = (box[:3] + rng.random((50000, 3)) * (box[3:] - box[:3])).astype(np.float32)
    for devices in device_lists(sdf):
        replicas = sdf.ExactOctreeSdf.build_on_devices(mesh, bb, 6, 3, devices, 32, 2)
        for k, r in enumerate(replicas):
            assert sha(r.getOctreeData()) == sha(want.getOctreeData()), (devices, k)
            assert sha(r.getTrianglesSets()) == sha(want.getTrianglesSets()) and sha(r.getTrianglesMasks()) == sha(want.getTrianglesMasks())
            i1, i0 = r.info(), want.info()
            assert (i1.max_triangles_in_leafs, i1.max_triangles_encoded_in_leafs) == (i0.max_triangles_in_leafs, i0.max_triangles_encoded_in_leafs)
            assert np.array_equal(r.getDistance(q).view(np.uint32), want.getDistance(q).view(np.uint32))


def test_multi_device_argument_errors(sdf):
    v, i = displaced_sphere(2)
    box = sdf.meshes.bounding_box_with_margin(v)
    mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
    with pytest.raises(sdf.SdfB200Error):
        sdf.OctreeSdf.build_on_devices(mesh, bb, 5, 3, [99])
    os.environ.pop("SDFB200_ALLOW_DUPLICATE_DEVICES", None)
    with pytest.raises(sdf.SdfB200Error):
        sdf.OctreeSdf.build_on_devices(mesh, bb, 5, 3, [0, 0])
    assert len(sdf.OctreeSdf.build_on_devices(mesh, bb, 5, 3, [0])) == 1
