"""Sharded construction on the GPU. The whole C-ABI protocol (build_*_shard -> shard_sizes ->
shard_finish -> shard_export -> assemble) is exercised on ONE GPU by playing every rank in turn and doing the two
collectives by hand; the assembled arrays must be bit-identical to the single-rank build on every "rank".
With >= 2 GPUs the same is run for real: one process per GPU, NCCL all-reduce + all-gather (sdflib_b200.sharded)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import assert_bit_equal, displaced_sphere, ROOT

pytestmark = pytest.mark.gpu


def play_ranks(sdf, make_shard, world):
    """make_shard(rank, world) -> sharded.Shard in phase 1. Returns the assembled Shard objects."""
    import torch
    shards = [make_shard(r, world) for r in range(world)]
    sizes = [s.sizes() for s in shards]
    for a in range(world):
        for b in range(a + 1, world):
            assert not np.any((sizes[a] != 0) & (sizes[b] != 0)), "two ranks own the same root"
    total = np.sum(np.stack(sizes).astype(np.int64), 0).astype(np.uint32)
    for s in shards:
        s.finish(total)
    counts = np.array([s.payload_words() for s in shards], np.uint64)
    stride = int((int(counts.max()) + 3) // 4 * 4)
    gathered = torch.zeros(world * stride, dtype=torch.int32, device="cuda")
    for r, s in enumerate(shards):
        s.export(gathered[r * stride:(r + 1) * stride])
    for s in shards:
        s.assemble(gathered, counts, stride)
    return shards


@pytest.mark.parametrize("world,threads", [(2, 2), (3, 1), (8, 2)])
def test_octree_shards_assemble_to_the_single_rank_build(sdf, world, threads):
    from sdflib_b200 import sharded, _capi
    v, i = displaced_sphere(3)
    box = sdf.meshes.bounding_box_with_margin(v)
    mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
    single = sdf.OctreeSdf(mesh, bb, 6, 2, 1e-3, sdf.OctreeSdf.NO_CONTINUITY, threads)

    def make(rank, w):
        h = C.c_void_p()
        _capi.check(_capi.lib().sdfb200_build_octree_shard(
            *sharded._mesh_args(mesh, bb), C.c_uint32(6), C.c_uint32(2), C.c_int(1), C.c_float(1e-3), C.c_float(0.0), C.c_int(1),
            C.c_uint32(threads), C.c_uint32(rank), C.c_uint32(w), C.byref(h)))
        return sharded.Shard(h.value)

    want = single.getOctreeData()
    q = (single.getSampleArea().as_array()[:3] + np.random.default_rng(3).uniform(0, 1, (20000, 3)) * (box[3:] - box[:3]).max()).astype(np.float32)
    for s in play_ranks(sdf, make, world):
        o = s.into(sdf.OctreeSdf)
        assert np.array_equal(o.getOctreeData(), want)
        assert_bit_equal(np.float32([o.info().value_range, o.info().min_border_value]),
                         np.float32([single.info().value_range, single.info().min_border_value]))
        assert_bit_equal(o.getDistance(q), single.getDistance(q))


@pytest.mark.parametrize("world,threads", [(2, 2), (4, 1)])
def test_exact_shards_assemble_to_the_single_rank_build(sdf, world, threads):
    from sdflib_b200 import sharded, _capi
    v, i = displaced_sphere(3)
    box = sdf.meshes.bounding_box_with_margin(v)
    mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
    single = sdf.ExactOctreeSdf(mesh, bb, 5, 2, 16, threads)

    def make(rank, w):
        h = C.c_void_p()
        _capi.check(_capi.lib().sdfb200_build_exact_shard(
            *sharded._mesh_args(mesh, bb), C.c_uint32(5), C.c_uint32(2), C.c_uint32(16), C.c_uint32(threads), C.c_uint32(rank),
            C.c_uint32(w), C.byref(h)))
        return sharded.Shard(h.value)

    q = (single.getSampleArea().as_array()[:3] + np.random.default_rng(4).uniform(0, 1, (20000, 3)) * (box[3:] - box[:3]).max()).astype(np.float32)
    for s in play_ranks(sdf, make, world):
        o = s.into(sdf.ExactOctreeSdf)
        assert np.array_equal(o.getOctreeData(), single.getOctreeData())
        assert np.array_equal(o.getTrianglesSets(), single.getTrianglesSets())
        assert np.array_equal(o.getTrianglesMasks(), single.getTrianglesMasks())
        i1, i2 = o.info(), single.info()
        assert (i1.max_triangles_in_leafs, i1.max_triangles_encoded_in_leafs) == (i2.max_triangles_in_leafs, i2.max_triangles_encoded_in_leafs)
        assert_bit_equal(o.getDistance(q), single.getDistance(q))


def test_unassembled_shard_refuses_queries(sdf):
    from sdflib_b200 import sharded, _capi
    v, i = displaced_sphere(2)
    box = sdf.meshes.bounding_box_with_margin(v)
    mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
    h = C.c_void_p()
    _capi.check(_capi.lib().sdfb200_build_octree_shard(
        *sharded._mesh_args(mesh, bb), C.c_uint32(4), C.c_uint32(2), C.c_int(1), C.c_float(1e-3), C.c_float(0.0), C.c_int(1),
        C.c_uint32(2), C.c_uint32(0), C.c_uint32(2), C.byref(h)))
    s = sharded.Shard(h.value)
    n = C.c_uint64()
    assert _capi.lib().sdfb200_shard_words(s._h, C.byref(n)) == _capi.ERR_INVALID       # finish not called yet
    s.finish(np.maximum(s.sizes(), 64))                                                 # plausible sizes for the foreign roots
    d = np.empty(1, np.float32)
    code = _capi.lib().sdfb200_query(s._h, _capi.ptr(np.zeros(3, np.float32)), C.c_uint64(1), _capi.ptr(d), None, C.c_int(0), None)
    assert code == _capi.ERR_INVALID


def test_nccl_multi_process(sdf):
    """One process per GPU over NCCL (needs >= 2 GPUs; the 1-GPU round-end run skips it)."""
    n = sdf.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(ROOT, "scripts", "sharded_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sharded ok" in r.stdout
