"""CPU model (not a test; run by hand): how many BVH nodes a nearest-triangle search visits on the C2 workload when its
running best starts from an UPPER BOUND instead of DBL_MAX.   python tests/model_bvh_seed.py [samples]
Seeds tried: the exact distance (the ceiling of the idea), and |d(q)| + |p - q| from a point q at one node half-size
(what the builder knows when it samples the mid-points of a node: its corner values). Winners must not change."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.binding import port            # noqa: E402
from sdflib_b200 import meshes             # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
v, i = meshes.config_mesh("M1")
rng = np.random.default_rng(3)
tri = i.reshape(-1, 3)
box = meshes.bounding_box_with_margin(v)
size = float((box[3:] - box[:3]).max())
for depth in (5, 6, 7, 8):
    h = size / 2 ** depth / 2                       # node half size
    # sample positions like the build's: near the surface, within a few node sizes
    t = tri[rng.integers(0, len(tri), n)]
    w = rng.dirichlet((1, 1, 1), n).astype(np.float32)
    p = (v[t[:, 0]] * w[:, :1] + v[t[:, 1]] * w[:, 1:2] + v[t[:, 2]] * w[:, 2:]) + rng.normal(0, 1.5 * h, (n, 3))
    p = p.astype(np.float32)
    win, vis = port.nearest_triangle_visits(v, i, p)
    # exact unsigned distance through a second query set: distance of p = |sdf| from the port's point-triangle kernel
    d = np.abs(port.mesh_distance(v, i, p)) if hasattr(port, "mesh_distance") else None
    if d is None:
        # distance to the winner triangle, float64 numpy (closest point on triangle by clamped barycentric projection is overkill: use
        # the seeded query itself to bracket: bisection on the seed would be slow, so take the vertex-based bound of the winner)
        a, b, c = (v[tri[win][:, k]].astype(np.float64) for k in range(3))
        pp = p.astype(np.float64)
        # exact point-triangle distance (Ericson)
        ab, ac, ap = b - a, c - a, pp - a
        d1, d2 = (ab * ap).sum(1), (ac * ap).sum(1)
        bp = pp - b; d3, d4 = (ab * bp).sum(1), (ac * bp).sum(1)
        cp = pp - c; d5, d6 = (ab * cp).sum(1), (ac * cp).sum(1)
        vc = d1 * d4 - d3 * d2; vb = d5 * d2 - d1 * d6; va = d3 * d6 - d5 * d4
        denom = va + vb + vc
        vv = np.where(denom != 0, vb / np.where(denom == 0, 1, denom), 0); ww = np.where(denom != 0, vc / np.where(denom == 0, 1, denom), 0)
        q = a + ab * vv[:, None] + ac * ww[:, None]
        m = (d1 <= 0) & (d2 <= 0); q[m] = a[m]
        m = (d3 >= 0) & (d4 <= d3); q[m] = b[m]
        m = (d6 >= 0) & (d5 <= d6); q[m] = c[m]
        m = (vc <= 0) & (d1 >= 0) & (d3 <= 0); tt = d1 / np.where(d1 - d3 == 0, 1, d1 - d3); q[m] = (a + ab * tt[:, None])[m]
        m = (vb <= 0) & (d2 >= 0) & (d6 <= 0); tt = d2 / np.where(d2 - d6 == 0, 1, d2 - d6); q[m] = (a + ac * tt[:, None])[m]
        m = (va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0); tt = (d4 - d3) / np.where((d4 - d3) + (d5 - d6) == 0, 1, (d4 - d3) + (d5 - d6)); q[m] = (b + (c - b) * tt[:, None])[m]
        d = np.linalg.norm(pp - q, axis=1)
    base = vis.sum(1).mean()
    out = [f"depth {depth}: h = {h:.5f}  unseeded {base:7.1f} visits ({vis[:,0].mean():.0f} inner + {vis[:,1].mean():.0f} leaf)"]
    for name, seed in (("exact * (1 + 1e-6)", d * (1 + 1e-6) + 1e-9), ("d + h", d + h), ("d + 2h", d + 2 * h), ("d + 3.5h", d + 3.5 * h)):
        w2, vis2 = port.nearest_triangle_visits_seeded(v, i, p, seed)
        same = (w2 == win).mean()
        out.append(f"   seed {name:20s}: {vis2.sum(1).mean():7.1f} visits ({vis2[:,0].mean():.0f} + {vis2[:,1].mean():.0f})  x{base / vis2.sum(1).mean():.2f}  winners equal {same:.4f}")
    print("\n".join(out), flush=True)
