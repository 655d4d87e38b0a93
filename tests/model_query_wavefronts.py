"""CPU model of the L1 wavefronts of the OctreeSdf bulk query on the C2 workload (not a test; run by hand):

    python tests/model_query_wavefronts.py

Builds the C2 octree with the oracle (test infrastructure — the compiled reference when present, ~10 s with 8 threads),
walks the 256^3 cell-centre grid the way the kernels do and counts, per warp of 32 consecutive queries,
  * the distinct 128-byte lines of every descent gather,
  * the distinct leaves (one wavefront per distinct leaf for each of the 16 coefficient loads of octreeQueryKernel),
  * the rounds / load wavefronts / shuffles of the quad-cooperative variant for group widths 2, 4, 8.
Source of the numbers quoted in DESIGN.md section 8 and in octree_query.cu."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.binding import port, ref            # noqa: E402
from sdflib_b200 import meshes                  # noqa: E402

LEAF, MASK = np.uint32(1 << 31), np.uint32(0x3FFFFFFF)


def distinct_per_row(a, invalid=None):
    s = np.sort(a, axis=1)
    d = 1 + (s[:, 1:] != s[:, :-1]).sum(1)
    if invalid is not None:
        d -= (s[:, -1] == invalid)
    return d


def main():
    v, i = meshes.config_mesh("M1")
    box = meshes.bounding_box_with_margin(v)
    backend = ref if ref.available() else port
    t = time.time()
    sdf = backend.build_octree(v, i, box, 8, 3, threshold=1e-3, algorithm=1, num_threads=os.cpu_count() or 1)
    oct_, G = sdf.octree_data(), sdf.header()["start_grid_size"]
    print(f"{backend.kind} build {time.time() - t:.1f} s, {len(oct_)} words")
    L, N = 5, 256
    cz, cy, cx = (a.ravel() for a in np.meshgrid(*(np.arange(N, dtype=np.uint32),) * 3, indexing="ij"))
    node = oct_[((cz >> L) * G + (cy >> L)) * G + (cx >> L)]
    steps = np.zeros(node.shape, np.int8)
    gathers = np.ones(N ** 3 // 32, np.int64)            # the start-grid load: one line per warp
    for k in range(L):
        inner = (node & LEAF) == 0
        sh = L - 1 - k
        child = ((cx >> sh) & 1) | (((cy >> sh) & 1) << 1) | (((cz >> sh) & 1) << 2)
        addr = (node & MASK) + child
        gathers += distinct_per_row(np.where(inner, addr >> 5, np.uint32(0xFFFFFFFF)).reshape(-1, 32), 0xFFFFFFFF)
        node = np.where(inner, oct_[np.where(inner, addr, 0)], node)
        steps += inner
    block = node & MASK
    print("queries by leaf depth (steps below the start grid):", np.bincount(steps).tolist())
    lw = distinct_per_row(block.reshape(-1, 32))
    print(f"descent gathers: {gathers.mean():.1f} wavefronts per warp")
    print(f"distinct leaves per warp: mean {lw.mean():.2f} -> coefficient fill of octreeQueryKernel: {16 * lw.mean():.1f} wavefronts per warp")
    nw = block.size // 32
    for w in (2, 4, 8):
        grp = block.reshape(-1, w)
        cls = np.zeros(grp.shape, np.int8)                # class of a lane = rank of its leaf's first occurrence in the group
        ncls = np.ones(len(grp), np.int8)
        for l in range(1, w):
            seen, c = np.zeros(len(grp), bool), np.zeros(len(grp), np.int8)
            for m in range(l):
                eq = (grp[:, l] == grp[:, m]) & ~seen
                c, seen = np.where(eq, cls[:, m], c), seen | eq
            cls[:, l] = np.where(seen, c, ncls)
            ncls = ncls + ~seen
        rounds = ncls.reshape(-1, 32 // w).max(1)         # the warp loops until its slowest group is done
        loads = 0
        for t in range(int(rounds.max())):
            active = ncls > t
            leaf_t = np.where(active, np.take_along_axis(grp, (cls == t).argmax(1)[:, None], 1)[:, 0], np.uint32(0xFFFFFFFF))
            loads += (16 // w) * distinct_per_row(leaf_t.reshape(-1, 32 // w), 0xFFFFFFFF).sum()
        shuffles = rounds.sum() * (4 * int(np.log2(w)) + 4)   # 3 broadcasts + ballot + log2(w) stages of 4 sums
        print(f"cooperative width {w}: {rounds.mean():.2f} rounds per warp, {loads / nw:.1f} load wavefronts + {shuffles / nw:.1f} shuffles"
              f" = {(loads + shuffles) / nw:.1f} per warp (queries of a class share leaf, y and z: rows of the grid)")


if __name__ == "__main__":
    main()
