"""Exact-octree timing probe on the GPU box: config 3 (M1, depth 7, minTri 128) build + 256^3 queries."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sdflib_b200 as S
from sdflib_b200 import meshes

name, depth, start, mintri = (sys.argv[1:] + ["M1", "7", "3", "128"])[:4]
depth, start, mintri = int(depth), int(start), int(mintri)
v, i = meshes.config_mesh(name)
box = meshes.bounding_box_with_margin(v)
mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])
for rep in range(2):
    t0 = time.perf_counter()
    sdf = S.ExactOctreeSdf(mesh, bb, depth, start, mintri, 2)
    dt = time.perf_counter() - t0
    info = sdf.info()
    print(json.dumps({"build_s": dt, "stats": sdf.build_stats(), "nodes": int(info.octree_words), "sets": int(info.triangle_sets_words),
                      "masks": int(info.triangle_masks_bytes), "max_leaf": info.max_triangles_in_leafs,
                      "max_encoded": info.max_triangles_encoded_in_leafs}), flush=True)
    if rep == 0:
        sdf.close()
area = sdf.getSampleArea().as_array()
for n in (64, 256):
    pts = torch.from_numpy(meshes.cell_centre_grid(area, n)).cuda()
    out = torch.empty(len(pts), dtype=torch.float32, device="cuda")
    sdf.getDistance(pts, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sdf.getDistance(pts, out=out); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"grid": n, "ms": ms, "Mq_per_s": len(pts) / ms / 1e3, "checksum": float(out.double().sum())}), flush=True)
