"""A/B of the EXPERIMENTAL dense leaf index of the OctreeSdf bulk query (octree_query.cu, SDFB200_QUERY_INDEX=1).

For the C2 octree (and a depth-9 one, whose index is 512 MB): results with the index must be bit-identical to the plain
kernel — value, gradient, FMA and exact-order variants, grid-ordered and random points, points outside the box — and the
kernel time of both is printed (CUDA events on the launching stream, L2 flushed between launches).

    gpurun --timeout 900 -- 'python scripts/gpu_ab_query_index.py'
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdflib_b200 as S                      # noqa: E402
from sdflib_b200 import meshes               # noqa: E402


def timed(fn, reps=10):
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    best, total = 1e9, 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best, total = min(best, ms), total + ms
    return best, total / reps


def main():
    v, i = meshes.config_mesh("M1")
    box = meshes.bounding_box_with_margin(v)
    mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])
    ok = True
    for depth, n_grid in ((8, 256), (9, 512)):
        sdf = S.OctreeSdf(mesh, bb, depth, 3, 1e-3, S.OctreeSdf.NO_CONTINUITY, 1)
        area = sdf.getGridBoundingBox().as_array()
        rng = np.random.default_rng(42)
        sets = {
            f"grid {n_grid}^3": torch.from_numpy(meshes.cell_centre_grid(area, n_grid)).cuda(),
            "random 2^24": torch.from_numpy((area[:3] + rng.random((1 << 24, 3), np.float32) * (area[3:] - area[:3])).astype(np.float32)).cuda(),
            "around the box": torch.from_numpy((area[:3] - 0.2 + rng.random((1 << 20, 3), np.float32) * (area[3:] - area[:3] + 0.4)).astype(np.float32)).cuda(),
        }
        for name, pts in sets.items():
            for gradient in (False, True):
                for exact in (False, True):
                    out = {}
                    for sw in ("0", "1"):
                        os.environ["SDFB200_QUERY_INDEX"] = sw
                        r = sdf.getDistance(pts, gradient=gradient, exact_order=exact)
                        torch.cuda.synchronize()
                        out[sw] = [t.clone() for t in (r if gradient else (r,))]
                    same = all(torch.equal(a.view(torch.int32), b.view(torch.int32)) for a, b in zip(out["0"], out["1"]))
                    ok &= same
                    line = f"depth {depth} {name:15s} grad={int(gradient)} exact={int(exact)} identical={same}"
                    if not exact:
                        dist = torch.empty(pts.shape[0], dtype=torch.float32, device="cuda")
                        grad = torch.empty((pts.shape[0], 3), dtype=torch.float32, device="cuda") if gradient else None
                        for sw in ("0", "1"):
                            os.environ["SDFB200_QUERY_INDEX"] = sw
                            best, mean = timed(lambda: sdf.getDistance(pts, gradient=gradient, out=dist, out_gradient=grad))
                            line += f" | index={sw}: best {best:.3f} ms mean {mean:.3f} ms ({pts.shape[0] / best / 1e6:.1f} Gq/s)"
                    print(line, flush=True)
        sdf.close()
    print("ALL IDENTICAL" if ok else "MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
