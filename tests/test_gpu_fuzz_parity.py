"""GPU product against the oracle on random NON-MANIFOLD meshes (a `-m gpu` test with 12 meshes; by hand for more):

    python tests/test_gpu_fuzz_parity.py [meshes=40] [seed=21]

The committed GPU parity tests use sphere-like and hand-made edge-case meshes. This loop throws triangle soups,
duplicated vertices, duplicated triangles and holes at all three builders (random depth / start depth / rule /
threshold / minTrianglesPerNode) and at the queries, and compares with the history-free oracle bit for bit — the same
comparison as tests/test_gpu_octree.py::test_edge_case_meshes_bit_exact. The oracle itself is pinned against the
compiled reference on the same family of meshes (tests/test_oracle.py). First passed on a B200 in round 2 (40 meshes,
0 differing builds: profiles/r2_summary.md)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdflib_b200 as sdf                       # noqa: E402
from oracle.binding import port                 # noqa: E402
from test_capi_host import _random_meshes       # noqa: E402


import pytest                                   # noqa: E402


@pytest.mark.gpu
def test_random_non_manifold_meshes_bit_exact():
    assert run(12, 21) == 0


def run(count, seed):
    rng = np.random.default_rng(seed)
    bad = 0
    for n, (v, i) in enumerate(_random_meshes(rng, count)):
        lo, hi = v.min(0), v.max(0)
        m = 0.2 * float((hi - lo).max())
        box = np.concatenate([lo - m, hi + m]).astype(np.float32)
        mesh, bb = sdf.Mesh(v, i), sdf.BoundingBox(box[:3], box[3:])
        depth, start = int(rng.integers(3, 6)), int(rng.integers(0, 4))
        thr, rule = float(rng.choice([1e-3, 1e-2, 1e-1])), int(rng.integers(1, 4))
        for alg in (sdf.OctreeSdf.NO_CONTINUITY, sdf.OctreeSdf.CONTINUITY):
            g = sdf.OctreeSdf(mesh, bb, depth, start, thr, alg, 1, terminationRule=rule, terminationRuleParams=[thr, 0.1])
            p = port.build_octree(v, i, box, depth, start, thr, alg, 1, termination_rule=rule, param1=0.1, use_cache=False)
            same = np.array_equal(g.getOctreeData(), p.octree_data())
            q = (box[:3] + rng.random((20000, 3)) * (box[3:] - box[:3])).astype(np.float32)
            same_q = np.array_equal(g.getDistance(q, exact_order=True).view(np.uint32), p.query(q).view(np.uint32))
            if not (same and same_q):
                bad += 1
                print(f"mesh {n}: OctreeSdf differs (words {same}, queries {same_q}): tris {i.size // 3} depth {depth} start {start} thr {thr} rule {rule} alg {alg}", flush=True)
            g.close()
        if i.size // 3 >= 2:
            start = int(rng.integers(0, 3))
            depth, min_tris = start + int(rng.integers(2, 4)), int(rng.choice([1, 4, 16, 64]))
            e = sdf.ExactOctreeSdf(mesh, bb, depth, start, min_tris, 1)
            pe = port.build_exact(v, i, box, depth, start, min_tris, 1, use_cache=False)
            same = np.array_equal(e.getOctreeData().reshape(-1), pe.octree_data())
            q = (box[:3] + rng.random((20000, 3)) * (box[3:] - box[:3])).astype(np.float32)
            same_q = np.array_equal(e.getDistance(q).view(np.uint32), pe.query(q).view(np.uint32))
            if not (same and same_q):
                bad += 1
                print(f"mesh {n}: ExactOctreeSdf differs (nodes {same}, queries {same_q}): tris {i.size // 3} depth {depth} start {start} minTris {min_tris}", flush=True)
            e.close()
    print(f"{count} meshes, {bad} differing builds")
    return bad


if __name__ == "__main__":
    sys.exit(1 if run(int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 21) else 0)
