"""Small driver for ncu captures (one GPU): runs ONE build or ONE query pass of the bench workloads.
    python scripts/profile_kernels.py exact_build | exact_build_c4 | octree_build | octree_cont | exact_query | octree_query | octree_query_random | octree_query_grad"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdflib_b200 as S
from sdflib_b200 import meshes

what = sys.argv[1]
v, i = meshes.config_mesh("M2" if what.endswith("c4") else "M1")
box = meshes.bounding_box_with_margin(v)
mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])
if what == "exact_build_c4":
    sdf = S.ExactOctreeSdf(mesh, bb, 8, 3, 128, 2)
    print(sdf.build_stats())
    sdf.close()
    sdf = S.ExactOctreeSdf(mesh, bb, 8, 3, 128, 2)   # the second build is the warm one (device pool filled)
    print(sdf.build_stats())
elif what.startswith("exact"):
    sdf = S.ExactOctreeSdf(mesh, bb, 7, 3, 128, 2)
elif what == "octree_cont":
    sdf = S.OctreeSdf(mesh, bb, 8, 3, 1e-3, S.OctreeSdf.CONTINUITY, 2)
else:
    sdf = S.OctreeSdf(mesh, bb, 8, 3, 1e-3, S.OctreeSdf.NO_CONTINUITY, 2)
if "query" in what:
    area = sdf.getSampleArea().as_array()
    if what.endswith("random"):
        import numpy as np
        pts = torch.from_numpy((area[:3] + np.random.default_rng(42).random((1 << 24, 3), np.float32) * (area[3:] - area[:3])).astype(np.float32)).cuda()
    else:
        pts = torch.from_numpy(meshes.cell_centre_grid(area, 256)).cuda()
    out = torch.empty(len(pts), dtype=torch.float32, device="cuda")
    grad = torch.empty((len(pts), 3), dtype=torch.float32, device="cuda") if what.endswith("grad") else None
    for _ in range(3):
        sdf.getDistance(pts, out=out, gradient=grad is not None, out_gradient=grad)
    torch.cuda.synchronize()
print("done", what)
