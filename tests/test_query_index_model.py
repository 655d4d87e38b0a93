"""Host model of the EXPERIMENTAL dense leaf index of octree_query.cu (leafIndexKernel / octreeQueryIndexedKernel).

The CUDA code cannot run here; what can be checked on the CPU is the arithmetic it transcribes: the entry packing
((block - G^3) / 8, steps, leaf flag), the cell chosen from the start cell and the leading path bits, and the finish of
the descent for trees deeper than the index — against the plain descent of octreeQueryKernel (itself bit-exact with
the oracle on the GPU, tests/test_gpu_octree.py), on octrees built by the oracle."""
import numpy as np
import pytest

LEAF, MASK = np.uint32(1 << 31), np.uint32(~(3 << 30) & 0xFFFFFFFF)
PATH_BITS = 16


def plain_descent(oct_, G, ix, iy, iz, bx, by, bz):
    node = oct_[(iz * G + iy) * G + ix]
    k = 0
    while not node & LEAF:
        sh = PATH_BITS - 1 - k
        child = ((bx >> sh) & 1) | (((by >> sh) & 1) << 1) | (((bz >> sh) & 1) << 2)
        node = oct_[int(node & MASK) + child]
        k += 1
    return int(node & MASK), k


def build_index(oct_, G, levels):
    N, G3 = G << levels, G ** 3
    index = np.zeros(N ** 3, np.uint32)
    for i in range(N ** 3):
        cx, cy, cz = i % N, (i // N) % N, i // (N * N)
        node = oct_[((cz >> levels) * G + (cy >> levels)) * G + (cx >> levels)]
        k = 0
        while not node & LEAF and k < levels:
            sh = levels - 1 - k
            child = ((cx >> sh) & 1) | (((cy >> sh) & 1) << 1) | (((cz >> sh) & 1) << 2)
            node = oct_[int(node & MASK) + child]
            k += 1
        block = int(node & MASK)
        assert block >= G3 and (block - G3) % 8 == 0      # what the kernel reports through `bad`
        index[i] = ((block - G3) >> 3) | (((1 << 31) | (k << 27)) if node & LEAF else 0)
    return index


def indexed_descent(oct_, index, G, levels, ix, iy, iz, bx, by, bz):
    N, G3 = G << levels, G ** 3
    cx = (ix << levels) | (bx >> (PATH_BITS - levels))
    cy = (iy << levels) | (by >> (PATH_BITS - levels))
    cz = (iz << levels) | (bz >> (PATH_BITS - levels))
    e = int(index[(cz * N + cy) * N + cx])
    block = ((e & ((1 << 27) - 1)) << 3) + G3
    if e & (1 << 31):
        return block, (e >> 27) & 15
    k = levels
    while True:
        sh = PATH_BITS - 1 - k
        child = ((bx >> sh) & 1) | (((by >> sh) & 1) << 1) | (((bz >> sh) & 1) << 2)
        node = oct_[block + child]
        k += 1
        block = int(node & MASK)
        if node & LEAF:
            return block, k


@pytest.mark.parametrize("depth,start,levels", [(5, 2, 3), (5, 2, 1), (5, 2, 0), (6, 2, 2), (5, 3, 2)])
def test_indexed_descent_reaches_the_same_leaf(port, depth, start, levels):
    from sdflib_b200 import meshes
    v, i = meshes.isosphere(2)
    v = (v * np.float32([1.0, 0.8, 0.6]) + np.float32([0.013, -0.007, 0.003])).astype(np.float32)
    box = np.float32([-1.3, -1.3, -1.3, 1.3, 1.3, 1.3])
    sdf = port.build_octree(v, i, box, depth, start, threshold=3e-3, algorithm=1, use_cache=False)
    oct_ = sdf.octree_data()
    G = sdf.header()["start_grid_size"]
    assert G == 1 << start and levels <= depth - start
    index = build_index(oct_, G, levels)
    rng = np.random.default_rng(7)
    cells = rng.integers(0, G, size=(4000, 3))
    bits = rng.integers(0, 1 << PATH_BITS, size=(4000, 3))
    bits[:64] = 0                                  # cell corners
    bits[64:128] = (1 << PATH_BITS) - 1            # just below the next cell
    bits[128:192] = 1 << (PATH_BITS - 1)           # exactly on the first split plane (child = frac >= 0.5)
    depths = set()
    for (ix, iy, iz), (bx, by, bz) in zip(cells.tolist(), bits.tolist()):
        want = plain_descent(oct_, G, ix, iy, iz, bx, by, bz)
        got = indexed_descent(oct_, index, G, levels, ix, iy, iz, bx, by, bz)
        assert got == want
        depths.add(want[1])
    assert len(depths) >= 2                        # leaves above, at and below the index depth were all visited
