"""A/B of the OctreeSdf bulk-query kernels (octree_query.cu):

  tile  : the default FMA kernel (octreeQueryTileKernel: TMA-staged tiles, top index, quad-cooperative evaluation)
  plain : one query per thread (octreeQueryKernel; SDFB200_QUERY_PLAIN=1 when the structure is built)
  exact : the reference-order kernel behind SDFB200_QUERY_EXACT_ORDER (bit-exact)

on the C2 octree and a depth-9 one, for grid-ordered, random and partly-outside point sets, with and without
gradients. Checks: tile and plain within the FMA tolerance of the exact-order result (1e-5 relative, floor 1e-3 of the
box; points beyond it are counted), identical in/out-of-box classification. Kernel times by CUDA
events on the launching stream with L2 flushed between launches.

    gpurun --timeout 1200 -- 'python scripts/gpu_ab_query_variants.py'
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


def run(sdf, pts, gradient, exact):
    r = sdf.getDistance(pts, gradient=gradient, exact_order=exact)
    torch.cuda.synchronize()
    return [t.clone() for t in (r if gradient else (r,))]


def main():
    v, i = meshes.config_mesh("M1")
    box = meshes.bounding_box_with_margin(v)
    mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])
    ok = True
    for depth, n_grid in ((8, 256), (9, 512)):
        os.environ["SDFB200_QUERY_PLAIN"] = "1"
        plain_sdf = S.OctreeSdf(mesh, bb, depth, 3, 1e-3, S.OctreeSdf.NO_CONTINUITY, 1)
        os.environ["SDFB200_QUERY_PLAIN"] = "0"
        sdf = S.OctreeSdf(mesh, bb, depth, 3, 1e-3, S.OctreeSdf.NO_CONTINUITY, 1)
        area = sdf.getGridBoundingBox().as_array()
        size = float(area[3] - area[0])
        rng = np.random.default_rng(42)
        sets = {
            f"grid {n_grid}^3": torch.from_numpy(meshes.cell_centre_grid(area, n_grid)).cuda(),
            "random 2^24": torch.from_numpy((area[:3] + rng.random((1 << 24, 3), np.float32) * (area[3:] - area[:3])).astype(np.float32)).cuda(),
            "around the box": torch.from_numpy((area[:3] - 0.2 + rng.random(((1 << 20) + 5, 3), np.float32) * (area[3:] - area[:3] + 0.4)).astype(np.float32)).cuda(),
        }
        for name, pts in sets.items():
            for gradient in (False, True):
                exact = run(sdf, pts, gradient, True)
                plain = run(plain_sdf, pts, gradient, False)
                tile = run(sdf, pts, gradient, False)
                tol = 1e-5 * torch.clamp(exact[0].abs(), min=1e-3 * size)
                line = f"depth {depth} {name:15s} grad={int(gradient)}"
                for what, got in (("tile", tile), ("plain", plain)):
                    rel = (got[0] - exact[0]).abs() / tol
                    over = int((rel > 1.0).sum())
                    good = over <= 1e-5 * pts.shape[0] and rel.max().item() < 8.0 and (got[0] - exact[0]).abs().max().item() < 1e-6 * size
                    line += f" | {what}: max err/tol {rel.max().item():.2f}, {over} over, max abs {(got[0] - exact[0]).abs().max().item():.2e}"
                    if gradient:
                        fin = torch.isfinite(exact[1]).all(1) & torch.isfinite(got[1]).all(1)
                        gerr = (got[1][fin] - exact[1][fin]).abs().max().item()
                        good &= gerr <= 1e-3 and bool((torch.isfinite(exact[1]).all(1) == torch.isfinite(got[1]).all(1)).all())
                        line += f", grad max abs {gerr:.2e}"
                    ok &= good
                    line += f" ok={good}"
                print(line, flush=True)
                dist = torch.empty(pts.shape[0], dtype=torch.float32, device="cuda")
                grad = torch.empty((pts.shape[0], 3), dtype=torch.float32, device="cuda") if gradient else None
                line = "    kernel time"
                for what, obj, ex in (("tile", sdf, False), ("plain", plain_sdf, False), ("exact order", sdf, True)):
                    best, mean = timed(lambda: obj.getDistance(pts, gradient=gradient, exact_order=ex, out=dist, out_gradient=grad))
                    line += f" | {what}: best {best:.3f} ms, mean {mean:.3f} ms ({pts.shape[0] / best / 1e6:.1f} Gq/s)"
                print(line, flush=True)
        # a batch that does not start on a 16-byte boundary takes the kernel's plain-load path: same bits as the TMA path
        pts = sets["around the box"]
        shifted = torch.empty(pts.numel() + 1, dtype=torch.float32, device="cuda")[1:].view(-1, 3)
        shifted.copy_(pts)
        a, b = run(sdf, pts, True, False), run(sdf, shifted, True, False)
        same = all(torch.equal(x.view(torch.int32), y.view(torch.int32)) for x, y in zip(a, b))
        print(f"depth {depth} unaligned batch identical to aligned: {same}", flush=True)
        ok &= same
        sdf.close(); plain_sdf.close()
    print("ALL CHECKS PASSED" if ok else "CHECK FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
