"""Per-rank phase times of one warm single-process multi-device build (SDFB200_TIMING=1 is set here).
    gpurun --gpus 8 -- 'python scripts/gpu_multi_timing.py c4 8'"""
import os, sys
os.environ["SDFB200_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdflib_b200 as S
from sdflib_b200 import meshes
cfg, n = sys.argv[1], int(sys.argv[2])
v, i = meshes.config_mesh("M2" if cfg == "c4" else "M1")
mesh, bb = S.Mesh(v, i), S.BoundingBox(*np.split(meshes.bounding_box_with_margin(v), 2))
for rep in range(3):
    print(f"==== build {rep}", file=sys.stderr, flush=True)
    if cfg in ("c4", "c3"):
        r = S.ExactOctreeSdf.build_on_devices(mesh, bb, 8 if cfg == "c4" else 7, 3, list(range(n)), 128, 2)
    else:
        r = S.OctreeSdf.build_on_devices(mesh, bb, 8, 3, list(range(n)), 1e-3, 2 if cfg == "c2cont" else 1, 2)
    for x in r:
        x.close()
