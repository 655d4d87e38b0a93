"""A/B timing of the OctreeSdf query kernels on config 2 (run once per SDFB200_QUERY_KERNEL value)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sdflib_b200 as S
from sdflib_b200 import meshes
v, i = meshes.config_mesh("M1")
box = meshes.bounding_box_with_margin(v)
sdf = S.OctreeSdf(S.Mesh(v, i), S.BoundingBox(box[:3], box[3:]), 8, 3, 1e-3, S.OctreeSdf.NO_CONTINUITY, 2)
area = sdf.getSampleArea().as_array()
grid = torch.from_numpy(meshes.cell_centre_grid(area, 256)).cuda()
rng = np.random.default_rng(42)
rnd = torch.from_numpy((area[:3] + rng.random((1 << 24, 3), dtype=np.float32) * (area[3:] - area[:3])).astype(np.float32)).cuda()
res = {"kernel": os.environ.get("SDFB200_QUERY_KERNEL", "coop")}
for name, pts in (("grid256", grid), ("random16M", rnd)):
    for gradient in (False, True):
        out = torch.empty(len(pts), dtype=torch.float32, device="cuda")
        og = torch.empty((len(pts), 3), dtype=torch.float32, device="cuda")
        f = (lambda: sdf.getDistance(pts, gradient=True, out=out, out_gradient=og)) if gradient else (lambda: sdf.getDistance(pts, out=out))
        for _ in range(5): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        res[f"{name}{'_grad' if gradient else ''}"] = {"ms": round(ms, 4), "Gq_s": round(len(pts) / ms / 1e6, 2), "checksum": float(out.double().sum())}
print(json.dumps(res))
