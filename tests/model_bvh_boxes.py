"""CPU model (not a test; run by hand): BVH nodes visited per nearest-triangle search on the C2 workload with and without
the extra axis-aligned-box pruning of the device sampler (bvh_sampler.cuh), and whether any winner changes.
    python tests/model_bvh_boxes.py [samples per depth]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.binding import port            # noqa: E402
from sdflib_b200 import meshes             # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for name in ("M1",):
    v, i = meshes.config_mesh(name)
    rng = np.random.default_rng(3)
    tri = i.reshape(-1, 3)
    box = meshes.bounding_box_with_margin(v)
    size = float((box[3:] - box[:3]).max())
    sets = {}
    for depth in (3, 5, 7, 8):
        h = size / 2 ** depth / 2
        t = tri[rng.integers(0, len(tri), n)]
        w = rng.dirichlet((1, 1, 1), n).astype(np.float32)
        sets[f"depth {depth} (within ~1.5 half-sizes of the surface)"] = ((v[t[:, 0]] * w[:, :1] + v[t[:, 1]] * w[:, 1:2] + v[t[:, 2]] * w[:, 2:]) + rng.normal(0, 1.5 * h, (n, 3))).astype(np.float32)
    sets["uniform in the box"] = (box[:3] + rng.random((n, 3)) * (box[3:] - box[:3])).astype(np.float32)
    sets["exactly on vertices (ties between the triangles of a fan)"] = v[rng.integers(0, len(v), n)]
    e = tri[rng.integers(0, len(tri), n)]
    sets["edge mid-points (two-triangle ties)"] = ((v[e[:, 0]].astype(np.float64) + v[e[:, 1]]) / 2).astype(np.float32)
    for what, p in sets.items():
        win, vis = port.nearest_triangle_visits(v, i, p)
        win2, vis2 = port.nearest_triangle_visits_seeded(v, i, p, None)
        a, b = vis.sum(1).mean(), vis2.sum(1).mean()
        print(f"{name} {what:62s}: {a:7.1f} visits ({vis[:,0].mean():.0f} + {vis[:,1].mean():.0f}) -> {b:6.1f} ({vis2[:,0].mean():.0f} + {vis2[:,1].mean():.0f})  x{a / b:.2f}   winners changed: {int((win != win2).sum())}", flush=True)
