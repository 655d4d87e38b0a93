"""Sharded construction over the GPUs of one box: one process per GPU, start-depth voxels
partitioned across ranks (the reference's own task decomposition, src/sdf/OctreeSdfDepthFirst.h:433-469,
include/SdfLib/ExactOctreeSdfDepthFirst.h:534-574), one all-reduce of the per-voxel sizes (a few KB) and ONE
all-gather of the payload over NCCL/NVLink to assemble the final arrays on every rank.

The collectives live here (torch.distributed is plumbing); everything else is the C-ABI protocol of
include/sdfb200.h: build_*_shard -> shard_sizes -> shard_finish -> shard_words/export -> assemble.
`exchange()` only needs an object with that protocol, so its host logic is testable with gloo on CPU.
"""
import ctypes as C

import numpy as np

from . import _capi
from .sdf import OctreeSdf, ExactOctreeSdf, PreparedMesh, SdfFunction


class Shard:
    """A phase-1 shard handle of the C-ABI (owns the handle until `into()` hands it to an SdfFunction)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    def sizes(self):
        L = _capi.lib()
        n = C.c_uint64()
        _capi.check(L.sdfb200_shard_sizes(self._h, None, C.c_uint64(0), C.byref(n)))
        out = np.empty(n.value, np.uint32)
        _capi.check(L.sdfb200_shard_sizes(self._h, _capi.ptr(out), C.c_uint64(out.size), C.byref(n)))
        return out

    def finish(self, all_sizes):
        a = _capi.u32(all_sizes)
        _capi.check(_capi.lib().sdfb200_shard_finish(self._h, _capi.ptr(a), C.c_uint64(a.size)))

    def payload_words(self):
        n = C.c_uint64()
        _capi.check(_capi.lib().sdfb200_shard_words(self._h, C.byref(n)))
        return n.value

    def export(self, buf):
        """buf: int32 CUDA tensor on the shard's device with at least payload_words() elements."""
        _capi.check(_capi.lib().sdfb200_shard_export(self._h, C.c_void_p(buf.data_ptr()), C.c_uint64(buf.numel())))

    def assemble(self, gathered, words_per_rank, stride):
        w = np.ascontiguousarray(words_per_rank, dtype=np.uint64)
        _capi.check(_capi.lib().sdfb200_assemble(self._h, C.c_void_p(gathered.data_ptr()), _capi.ptr(w), C.c_uint64(stride),
                                                 C.c_uint32(len(w))))

    def buffer_device(self):
        import torch
        i = _capi.Info()
        _capi.check(_capi.lib().sdfb200_get_info(self._h, C.byref(i)))
        return torch.device("cuda", i.device)

    def into(self, cls):
        obj = cls.__new__(cls)
        SdfFunction.__init__(obj, self._h.value)
        self._h = None
        return obj

    def __del__(self):
        if getattr(self, "_h", None):
            _capi.lib().sdfb200_free(self._h)


def exchange(shard, group=None, device=None):
    """Runs the size all-reduce and the payload all-gather for `shard` over `group` (default process group).
    Returns (bytes all-gathered per rank, stride in words)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    device = device if device is not None else shard.buffer_device()
    own = shard.sizes()
    t = torch.from_numpy(own.astype(np.int64)).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)          # exactly one rank contributes a non-zero size per root
    shard.finish(t.cpu().numpy().astype(np.uint32))
    n = shard.payload_words()
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[dist.get_rank(group)] = n
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    counts = counts.cpu().numpy().astype(np.uint64)
    stride = int((int(counts.max()) + 3) // 4 * 4)
    buf = torch.zeros(stride, dtype=torch.int32, device=device)
    shard.export(buf)
    gathered = torch.empty(world * stride, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(gathered, buf, group=group)            # the one data-path collective
    shard.assemble(gathered, counts, stride)
    return stride * 4, stride


def prepare_mesh(mesh, bvh, exact, group=None):
    """The mesh prepared ONCE per node (the host set-up used to be repeated by every rank with
    cores / ranks threads). Without a BVH every rank ingests the mesh on its own GPU (TriangleData is a few kernels).
    With a BVH — host work whose result depends on std::sort's tie order — rank 0 builds it with all host cores and the
    prepared mesh travels as one device blob through a broadcast over NCCL/NVLink."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1 or not bvh:
        return PreparedMesh(mesh, bvh=bvh, exact=exact)
    rank = dist.get_rank(group)
    device = torch.device("cuda", torch.cuda.current_device())
    own = PreparedMesh(mesh, bvh=True, exact=exact, all_host_threads=True) if rank == 0 else None
    size = torch.tensor([own.blob_bytes() if own is not None else 0], dtype=torch.int64, device=device)
    dist.broadcast(size, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    blob = torch.empty(int(size.item()), dtype=torch.uint8, device=device)
    if own is not None:
        own.export_blob(blob.data_ptr(), blob.numel())
    dist.broadcast(blob, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return own if own is not None else PreparedMesh.from_blob(blob.data_ptr(), blob.numel())


def _mesh_args(mesh, box):
    return (_capi.ptr(mesh.vertices), C.c_uint32(len(mesh.vertices)), _capi.ptr(mesh.indices), C.c_uint32(mesh.indices.size),
            _capi.ptr(_capi.f32(box.as_array())))


def _box_arg(box):
    return _capi.ptr(_capi.f32(box.as_array()))


class _DevicePointer:
    """A library-owned device range as a CUDA array-interface object (torch.as_tensor wraps it without a copy)."""

    def __init__(self, pointer, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(pointer), False), "version": 2}


def torch_allgather_hook(group=None, device=None):
    """sdfb200_allgather_fn over torch.distributed: NCCL directly on the device pointers; gloo through host copies.
    device=None means the two pointers are HOST memory (how the hook's choreography is tested without a GPU)."""
    import torch
    import torch.distributed as dist

    def hook(_user, d_send, d_recv, nbytes):
        try:
            world = dist.get_world_size(group)
            if device is None:
                send = torch.frombuffer((C.c_ubyte * nbytes).from_address(d_send), dtype=torch.uint8)
                recv = torch.frombuffer((C.c_ubyte * (nbytes * world)).from_address(d_recv), dtype=torch.uint8)
            else:
                send = torch.as_tensor(_DevicePointer(d_send, nbytes), device=device)
                recv = torch.as_tensor(_DevicePointer(d_recv, nbytes * world), device=device)
            if dist.get_backend(group) == "nccl":
                dist.all_gather_into_tensor(recv, send, group=group)
            else:   # host-staged collective (tests)
                parts = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
                dist.all_gather(parts, send.cpu(), group=group)
                recv.copy_(torch.cat(parts).to(recv.device))
            return 0
        except Exception as e:   # never let an exception cross the C boundary
            import sys
            print(f"[sdflib_b200] all-gather hook failed: {e}", file=sys.stderr)
            return 1

    return _capi.ALLGATHER_FN(hook)


def build_octree_collective(mesh, box, depth, startDepth, params, terminationRule, rank, world, hook):
    """InitAlgorithm::CONTINUITY over `world` ranks: replicated octree logic, BVH sampling sliced over the ranks and
    all-gathered per depth through `hook` (an _capi.ALLGATHER_FN). Returns the complete OctreeSdf on every rank."""
    h = C.c_void_p()
    pm = mesh if isinstance(mesh, PreparedMesh) else PreparedMesh(mesh, bvh=True, exact=False)
    _capi.check(_capi.lib().sdfb200_build_octree_collective_from_mesh(
        pm._h, _box_arg(box), C.c_uint32(depth), C.c_uint32(startDepth), C.c_int(terminationRule), C.c_float(params[0]),
        C.c_float(params[1]), C.c_uint32(rank), C.c_uint32(world), hook, None, C.byref(h)))
    return Shard(h.value).into(OctreeSdf)


def build_octree_sharded(mesh, box, depth, startDepth, maxError=1e-3, initAlgorithm=OctreeSdf.NO_CONTINUITY, numThreads=2,
                         terminationRule=OctreeSdf.TRAPEZOIDAL_RULE, terminationRuleParams=None, group=None):
    """OctreeSdf(...) built cooperatively by all ranks of `group`; every rank returns the complete structure.
    `mesh`: a Mesh (prepared here, once per node) or a PreparedMesh from prepare_mesh() to reuse across builds."""
    import torch.distributed as dist
    params = list(terminationRuleParams) if terminationRuleParams is not None else [maxError]
    params += [0.0] * (2 - len(params))
    if not isinstance(mesh, PreparedMesh):
        mesh = prepare_mesh(mesh, True, False, group)
    if initAlgorithm == OctreeSdf.CONTINUITY:
        import torch
        hook = torch_allgather_hook(group, torch.device("cuda", torch.cuda.current_device()))
        return build_octree_collective(mesh, box, depth, startDepth, params, terminationRule, dist.get_rank(group),
                                       dist.get_world_size(group), hook)
    h = C.c_void_p()
    _capi.check(_capi.lib().sdfb200_build_octree_from_mesh(
        mesh._h, _box_arg(box), C.c_uint32(depth), C.c_uint32(startDepth), C.c_int(terminationRule), C.c_float(params[0]),
        C.c_float(params[1]), C.c_int(initAlgorithm), C.c_uint32(numThreads), C.c_uint32(dist.get_rank(group)),
        C.c_uint32(dist.get_world_size(group)), C.byref(h)))
    if dist.get_world_size(group) == 1:
        return Shard(h.value).into(OctreeSdf)
    shard = Shard(h.value)
    exchange(shard, group)
    return shard.into(OctreeSdf)


def build_exact_sharded(mesh, box, maxDepth, startDepth=1, minTrianglesPerNode=128, numThreads=2, group=None):
    """ExactOctreeSdf(...) built cooperatively by all ranks of `group`; every rank returns the complete structure."""
    import torch.distributed as dist
    h = C.c_void_p()
    if not isinstance(mesh, PreparedMesh):
        mesh = prepare_mesh(mesh, False, True, group)
    _capi.check(_capi.lib().sdfb200_build_exact_from_mesh(
        mesh._h, _box_arg(box), C.c_uint32(maxDepth), C.c_uint32(startDepth), C.c_uint32(minTrianglesPerNode),
        C.c_uint32(numThreads), C.c_uint32(dist.get_rank(group)), C.c_uint32(dist.get_world_size(group)), C.byref(h)))
    if dist.get_world_size(group) == 1:
        return Shard(h.value).into(ExactOctreeSdf)
    shard = Shard(h.value)
    exchange(shard, group)
    return shard.into(ExactOctreeSdf)
