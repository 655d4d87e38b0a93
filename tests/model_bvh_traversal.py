"""CPU model of warp occupancy in the BVH sampler of the OctreeSdf builds on the C2 workload (not a test; run by hand):

    python tests/model_bvh_traversal.py [max_depth_to_model]

The device sampler (octree_device.cuh, one thread per distinct sample position, one BVH node per loop trip) is bound
by FP64 issue at ~9 active threads per instruction. This model separates the two causes: lanes that finished their
traversal waiting for the longest one of their warp (what a lane-refill schedule would recover), and trips in which
the live lanes of a warp are split between inner-node steps and leaf (point-triangle) steps.
The octree comes from the oracle (test infrastructure; the compiled reference when present), the per-query visit
counts from the oracle port's BVH (same traversal order as the device: near child first, far child re-tested)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.binding import port, ref            # noqa: E402
from sdflib_b200 import meshes                  # noqa: E402

LEAF, MASK = np.uint32(1 << 31), np.uint32(0x3FFFFFFF)
# the 19 mid-points of a node in units of its half size: 12 edge centres, 6 face centres, the centre
OFFSETS = np.array([(x, y, z) for z in (-1, 0, 1) for y in (-1, 0, 1) for x in (-1, 0, 1) if (x == 0) + (y == 0) + (z == 0) >= 1],
                   np.float32)


def main():
    last = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    v, i = meshes.config_mesh("M1")
    box = meshes.bounding_box_with_margin(v)
    backend = ref if ref.available() else port
    sdf = backend.build_octree(v, i, box, 8, 3, threshold=1e-3, algorithm=1, num_threads=os.cpu_count() or 1)
    oct_, G = sdf.octree_data(), sdf.header()["start_grid_size"]
    area = sdf.sample_area()
    size = float(area[3] - area[0])
    # level by level, children in parent order (the device's level order)
    gz, gy, gx = (a.ravel() for a in np.meshgrid(*(np.arange(G),) * 3, indexing="ij"))
    half = np.float32(size / G / 2)
    centre = (area[:3] + (np.stack([gx, gy, gz], 1).astype(np.float32) * 2 + 1) * half).astype(np.float32)
    word = oct_[:G ** 3]
    depth = 3
    print("depth  nodes  samples  distinct  visits/sample (inner+leaf)  leaf share  tail utilisation (lanes busy until the warp's longest traversal ends)")
    while depth <= last:
        pts = (centre[:, None, :] + OFFSETS[None] * half).reshape(-1, 3).astype(np.float32)
        _, first = np.unique(pts.view([("", np.float32)] * 3).ravel(), return_index=True)   # exact-bit duplicates: first occurrence owns
        owners = np.sort(first)
        t = time.time()
        _, visits = port.nearest_triangle_visits(v, i, pts[owners])
        trips = visits.sum(1).astype(np.int64)
        pad = (-len(trips)) % 32
        per_warp = np.concatenate([trips, np.zeros(pad, np.int64)]).reshape(-1, 32)
        util = trips.sum() / (32 * per_warp.max(1).sum())
        print(f"{depth:5d} {len(word):6d} {len(pts):8d} {len(owners):9d}  {trips.mean():7.1f} ({visits[:, 0].mean():.1f} + {visits[:, 1].mean():.1f})"
              f"  {visits[:, 1].sum() / trips.sum():10.2f}  {util:8.2f}     [{time.time() - t:.0f} s]", flush=True)
        inner = (word & LEAF) == 0
        base = (word[inner] & MASK).astype(np.int64)
        word = oct_[(base[:, None] + np.arange(8)[None]).ravel()]
        c = np.arange(8)
        dirs = np.stack([(c & 1) * 2 - 1, ((c >> 1) & 1) * 2 - 1, ((c >> 2) & 1) * 2 - 1], 1).astype(np.float32)
        half = np.float32(half / 2)
        centre = (centre[inner][:, None, :] + dirs[None] * half).reshape(-1, 3).astype(np.float32)
        depth += 1


if __name__ == "__main__":
    main()
