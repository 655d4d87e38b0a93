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


@pytest.mark.parametrize("world", [2, 3])
def test_continuity_collective_build_equals_single_rank(sdf, world):
    """InitAlgorithm::CONTINUITY over several ranks, played on ONE GPU: `world` threads build concurrently and the
    all-gather hook exchanges the sample slices through a barrier in this process. Every rank must return exactly the
    single-rank structure (replicated logic + deterministic owners of the de-duplicated samples)."""
    import threading
    import torch
    from sdflib_b200 import sharded, _capi
    v, i = displaced_sphere(3)
    box = sdf.meshes.bounding_box_with_margin(v)
    mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
    single = sdf.OctreeSdf(mesh, bb, 6, 2, 1e-3, sdf.OctreeSdf.CONTINUITY, 1).getOctreeData()
    barrier = threading.Barrier(world)
    staging = {}
    calls = [0] * world

    def make_hook(rank):
        def hook(_user, d_send, d_recv, nbytes):
            try:
                send = torch.as_tensor(sharded._DevicePointer(d_send, nbytes), device="cuda")
                recv = torch.as_tensor(sharded._DevicePointer(d_recv, nbytes * world), device="cuda")
                staging[rank] = send.clone()
                torch.cuda.synchronize()
                barrier.wait(timeout=120)
                recv.copy_(torch.cat([staging[r] for r in range(world)]))
                torch.cuda.synchronize()
                barrier.wait(timeout=120)   # nobody overwrites its staging slot before everybody has read it
                calls[rank] += 1
                return 0
            except Exception as e:      # pragma: no cover
                print("hook failed:", e)
                barrier.abort()
                return 1
        return _capi.ALLGATHER_FN(hook)

    results, errors = [None] * world, []

    def run(rank):
        try:
            results[rank] = sharded.build_octree_collective(mesh, bb, 6, 2, [1e-3, 0.0], sdf.OctreeSdf.TRAPEZOIDAL_RULE, rank, world,
                                                            make_hook(rank)).getOctreeData()
        except Exception as e:
            errors.append(e)
            barrier.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads: t.start()
    for t in threads: t.join(timeout=300)
    assert not any(t.is_alive() for t in threads), "a rank is stuck in the collective build"
    assert not errors, errors
    assert calls[0] > 4 and len(set(calls)) == 1       # one exchange per sampled level + per fix-up round, same on every rank
    for r in range(world):
        assert results[r].size == single.size and np.array_equal(results[r], single), f"rank {r} differs"


def test_collective_entry_checks_arguments(sdf):
    from sdflib_b200 import sharded, _capi
    v, i = displaced_sphere(1)
    box = sdf.meshes.bounding_box_with_margin(v)
    mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
    h = C.c_void_p()
    args = (*sharded._mesh_args(mesh, bb), C.c_uint32(4), C.c_uint32(2), C.c_int(1), C.c_float(1e-3), C.c_float(0))
    L = _capi.lib()
    # NO_CONTINUITY is not a collective build; more than one rank needs a hook; CONTINUITY refuses the voxel-sharded entry
    assert L.sdfb200_build_octree_collective(*args, C.c_int(1), C.c_uint32(1), C.c_uint32(0), C.c_uint32(2), None, None, C.byref(h)) == _capi.ERR_INVALID
    assert L.sdfb200_build_octree_collective(*args, C.c_int(2), C.c_uint32(1), C.c_uint32(0), C.c_uint32(2), None, None, C.byref(h)) == _capi.ERR_INVALID
    assert L.sdfb200_build_octree_shard(*args, C.c_int(2), C.c_uint32(1), C.c_uint32(0), C.c_uint32(2), C.byref(h)) == _capi.ERR_UNSUPPORTED
    # one rank needs no hook and equals the plain constructor
    assert L.sdfb200_build_octree_collective(*args, C.c_int(2), C.c_uint32(1), C.c_uint32(0), C.c_uint32(1), None, None, C.byref(h)) == _capi.OK
    one = sharded.Shard(h.value).into(sdf.OctreeSdf)
    assert np.array_equal(one.getOctreeData(), sdf.OctreeSdf(mesh, bb, 4, 2, 1e-3, sdf.OctreeSdf.CONTINUITY, 1).getOctreeData())
