"""Host-side logic of the sharded build under world_size-2 gloo on CPU: the collective choreography
of sdflib_b200.sharded.exchange (size all-reduce, padded payload all-gather, assembly order) is driven with a
stand-in shard that speaks the C-ABI protocol on numpy arrays. The stand-in is cut out of a complete structure built
by the CPU oracle (test infrastructure), so a wrong ordering, padding or dtype conversion shows up as a wrong array."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def subtree_words(words, slot):
    w = int(words[slot])
    if w & 0x80000000:
        return 64
    c = w & 0x3FFFFFFF
    return 8 + sum(subtree_words(words, c + k) for k in range(8))


class FakeOctreeShard:
    """Speaks the protocol of sdflib_b200.sharded.Shard for an OctreeSdf in per-voxel layout (numThreads >= 2)."""

    def __init__(self, complete, g3, rank, world):
        self.complete, self.g3, self.rank, self.world = complete, g3, rank, world
        self.size = np.array([subtree_words(complete, s) for s in range(g3)], np.uint32)
        self.owned = (np.arange(g3) % world) == rank            # layout order == slot order in this layout
        self.result = None

    def sizes(self):
        return np.where(self.owned, self.size, 0).astype(np.uint32)

    def finish(self, all_sizes):
        assert np.array_equal(all_sizes, self.size), "all-reduced sizes differ from the true sizes"
        self.base = self.g3 + np.concatenate([[0], np.cumsum(all_sizes.astype(np.int64))[:-1]])
        self.all = all_sizes

    def _payload(self):
        parts = [np.array([self.rank + 1, 100 - self.rank], np.uint32)]
        for s in np.nonzero(self.owned)[0]:
            parts += [self.complete[s:s + 1], self.complete[self.base[s]:self.base[s] + self.all[s]]]
        return np.concatenate(parts).astype(np.uint32)

    def payload_words(self):
        return self._payload().size

    def export(self, buf):
        p = self._payload()
        buf.numpy()[:p.size] = p.view(np.int32)

    def assemble(self, gathered, counts, stride):
        g = gathered.numpy().view(np.uint32)
        out = np.zeros_like(self.complete)
        at = [2] * self.world
        self.scalars = (max(int(g[q * stride]) for q in range(self.world)), min(int(g[q * stride + 1]) for q in range(self.world)))
        for s in range(self.g3):
            q = s % self.world
            src = g[q * stride:(q + 1) * stride]
            out[s] = src[at[q]]
            out[self.base[s]:self.base[s] + self.all[s]] = src[at[q] + 1:at[q] + 1 + self.all[s]]
            at[q] += 1 + int(self.all[s])
        assert [int(c) for c in counts] == at
        self.result = out


def _worker(rank, world, port_no, complete, g3, ret):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from sdflib_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shard = FakeOctreeShard(complete, g3, rank, world)
        nbytes, stride = sharded.exchange(shard, device=torch.device("cpu"))
        ok = bool(np.array_equal(shard.result, complete)) and shard.scalars == (world, 100 - (world - 1)) and stride % 4 == 0
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_reassembles_the_structure_under_gloo(port, world):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from conftest import displaced_sphere
    from sdflib_b200 import meshes
    v, i = displaced_sphere(2)
    box = meshes.bounding_box_with_margin(v)
    complete = port.build_octree(v, i, box, 4, 2, 1e-3, 1, 2, use_cache=False).octree_data()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port_no = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port_no, complete, 64, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)


# ---- the all-gather hook of the collective CONTINUITY build (sdfb200_allgather_fn) under gloo --------------------------
def _hook_worker(rank, world, port_no, ret):
    import ctypes as C
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from sdflib_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hook = sharded.torch_allgather_hook(device=None)      # host pointers: the C side of the protocol, without a GPU
        ok = True
        for per in (1, 16, 4096):                              # bytes per rank of three "levels"
            send = np.full(per, rank + 1, np.uint8)
            send[0] = per % 251
            recv = np.zeros(per * world, np.uint8)
            rc = hook(None, send.ctypes.data, recv.ctypes.data, per)
            want = np.concatenate([np.concatenate([[per % 251], np.full(per - 1, r + 1)]) for r in range(world)]).astype(np.uint8)
            ok = ok and rc == 0 and np.array_equal(recv, want)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_allgather_hook_gathers_in_rank_order_under_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port_no = 31500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_hook_worker, args=(r, world, port_no, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert dict(ret) == {r: True for r in range(world)}
