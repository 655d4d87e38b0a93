"""BASELINE.json configs 4 and 5 at their full sizes on ONE B200 (synthetic Dragon-/Lucy-class meshes):
    python scripts/gpu_big_configs.py c4 | c5 [reference]
  c4: M2 (5 242 880 triangles), ExactOctreeSdf(depth 8, startDepth 3, minTrianglesPerNode 128), 256^3 queries
  c5: M3 (20 971 520 triangles), OctreeSdf(depth 9, startDepth 3, threshold 1e-4, CONTINUITY), 512^3 value+gradient
Prints one JSON line per config (build seconds and phases, structure sizes, query throughput, free HBM)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sdflib_b200 as S
from sdflib_b200 import meshes

what = sys.argv[1]
c5_depth = int(os.environ.get("C5_DEPTH", "9"))
c5_thr = float(os.environ.get("C5_THRESHOLD", "1e-4"))
t0 = time.perf_counter()
v, i = meshes.config_mesh("M2" if what == "c4" else "M3")
box = meshes.bounding_box_with_margin(v)
mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])
mesh_s = time.perf_counter() - t0
out = {"config": what, "triangles": int(i.size // 3), "mesh_generation_s": round(mesh_s, 2)}
if what == "c5":
    out.update(depth=c5_depth, threshold=c5_thr)
if len(sys.argv) > 2 and sys.argv[2] == "reference":
    from oracle.binding import ref
    cores = os.cpu_count()
    r = ref.build_exact(v, i, box, 8, 3, 128, cores) if what == "c4" else ref.build_octree(v, i, box, c5_depth, 3, c5_thr, 2, cores)
    out.update(reference_build_s=r.build_seconds, threads=cores)
    print(json.dumps(out)); sys.exit(0)

def build():
    torch.cuda.synchronize(); t = time.perf_counter()
    s = S.ExactOctreeSdf(mesh, bb, 8, 3, 128, 2) if what == "c4" else S.OctreeSdf(mesh, bb, c5_depth, 3, c5_thr, S.OctreeSdf.CONTINUITY, 2)
    torch.cuda.synchronize()
    return s, time.perf_counter() - t

sdf, first = build()
out["build_first_s"] = round(first, 3)
out["stats_ms"] = {k: round(x, 1) for k, x in sdf.build_stats().items() if k.endswith("_ms")}
out["stats_counts"] = {k: int(x) for k, x in sdf.build_stats().items() if not k.endswith("_ms")}
sdf.close()
sdf, second = build()
out["build_s"] = round(second, 3)
info = sdf.info()
out["octree_words"] = int(info.octree_words)
if what == "c4":
    out.update(set_words=int(info.triangle_sets_words), mask_bytes=int(info.triangle_masks_bytes), max_triangles_in_leafs=int(info.max_triangles_in_leafs))
free, total = torch.cuda.mem_get_info()
out["hbm_used_gb"] = round((total - free) / 1e9, 2)
n = 256 if what == "c4" else 512
area = sdf.getSampleArea().as_array()
pts = torch.from_numpy(meshes.cell_centre_grid(area, n)).cuda()
dist = torch.empty(len(pts), dtype=torch.float32, device="cuda")
grad = torch.empty((len(pts), 3), dtype=torch.float32, device="cuda") if what == "c5" else None
def q():
    if grad is None: sdf.getDistance(pts, out=dist)
    else: sdf.getDistance(pts, out=dist, gradient=True, out_gradient=grad)
for _ in range(2): q()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 5
e0.record()
for _ in range(steps): q()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
out.update(query_grid=n, query_ms=round(ms, 3), queries_per_s=len(pts) / ms * 1e3, gradient=grad is not None,
           checksum=float(dist.double().sum().item()), nan=int(torch.isnan(dist).sum().item()))
print(json.dumps(out))
