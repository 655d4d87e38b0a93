"""Kernel times of the reference-order (bit-exact) OctreeSdf query on the C2 structure: 256^3 grid and 2^24 random points, value and
value + gradient, with a sha1 of the results (must not move when only the load width / pass structure of the kernel changes).
    gpurun --timeout 200 -- 'python scripts/gpu_exact_order_timing.py'"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sdflib_b200 as S
from sdflib_b200 import meshes

v, i = meshes.config_mesh("M1"); box = meshes.bounding_box_with_margin(v)
sdf = S.OctreeSdf(S.Mesh(v, i), S.BoundingBox(box[:3], box[3:]), 8, 3, 1e-3, S.OctreeSdf.NO_CONTINUITY, 1)
area = sdf.getGridBoundingBox().as_array()
rng = np.random.default_rng(42)
sets = {"grid 256^3": torch.from_numpy(meshes.cell_centre_grid(area, 256)).cuda(),
        "random 2^24": torch.from_numpy((area[:3] + rng.random((1 << 24, 3), np.float32) * (area[3:] - area[:3])).astype(np.float32)).cuda()}
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for name, pts in sets.items():
    for gradient in (False, True):
        dist = torch.empty(pts.shape[0], dtype=torch.float32, device="cuda")
        grad = torch.empty((pts.shape[0], 3), dtype=torch.float32, device="cuda") if gradient else None
        fn = lambda: sdf.getDistance(pts, gradient=gradient, exact_order=True, out=dist, out_gradient=grad)
        for _ in range(2):
            fn()
        best = 1e9
        for _ in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        h = hashlib.sha1(dist.cpu().numpy().tobytes() + (grad.cpu().numpy().tobytes() if gradient else b"")).hexdigest()[:12]
        print("exact order", name, "grad=%d" % gradient, "best %.4f ms (%.1f Gq/s)" % (best, pts.shape[0] / best / 1e6), h, flush=True)
