"""Single-process multi-device builds (sdfb200_build_*_multi) at 1 .. N GPUs of one box: seconds (first call and best of 3
warm), parity of every replica with the 1-device build (sha256 of the arrays), per-phase statistics of rank 0.

    gpurun --gpus 8 -- 'python scripts/gpu_multi_scaling.py c4 c3 c2 c2cont'          # default: all four
  c4     BASELINE config 4: M2 (5 242 880 triangles), ExactOctreeSdf depth 8, start depth 3, minTrianglesPerNode 128
  c3     config 3: M1, ExactOctreeSdf depth 7        c2 / c2cont: config 2, OctreeSdf depth 8 NO_CONTINUITY / CONTINUITY
Prints one JSON line per (config, N)."""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdflib_b200 as S                      # noqa: E402
from sdflib_b200 import meshes               # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def digest(s):
    if isinstance(s, S.ExactOctreeSdf):
        return sha(s.getOctreeData()) + sha(s.getTrianglesSets()) + sha(s.getTrianglesMasks())
    return sha(s.getOctreeData())


def main():
    configs = sys.argv[1:] or ["c4", "c3", "c2", "c2cont"]
    n_gpus = torch.cuda.device_count()
    counts = [n for n in (1, 2, 4, 8) if n <= n_gpus]
    print(json.dumps({"gpus": n_gpus, "nccl": bool(S.lib().sdfb200_nccl_available()), "host_cores": os.cpu_count()}), flush=True)
    meshes_cache = {}
    for cfg in configs:
        name = "M2" if cfg == "c4" else "M1"
        if name not in meshes_cache:
            v, i = meshes.config_mesh(name)
            meshes_cache[name] = (S.Mesh(v, i), S.BoundingBox(*np.split(meshes.bounding_box_with_margin(v), 2)))
        mesh, bb = meshes_cache[name]

        def build(devices):
            if cfg == "c4":
                return S.ExactOctreeSdf.build_on_devices(mesh, bb, 8, 3, devices, 128, 2)
            if cfg == "c3":
                return S.ExactOctreeSdf.build_on_devices(mesh, bb, 7, 3, devices, 128, 2)
            return S.OctreeSdf.build_on_devices(mesh, bb, 8, 3, devices, 1e-3, 2 if cfg == "c2cont" else 1, 2)

        reference_digest = None
        for n in counts:
            devices = list(range(n))
            times = []
            stats = None
            parity = None
            for rep in range(4):
                for d in devices:
                    torch.cuda.synchronize(d)
                t0 = time.perf_counter()
                replicas = build(devices)
                for d in devices:
                    torch.cuda.synchronize(d)
                times.append(time.perf_counter() - t0)
                if rep == 3:
                    stats = {k: round(x, 1) for k, x in replicas[0].build_stats().items() if k.endswith("_ms")}
                    per_rank = [r.build_stats() for r in replicas]
                    stats["levels_ms_max"] = round(max(s["levels_ms"] for s in per_rank), 1)       # load balance of the voxel plan
                    stats["levels_ms_min"] = round(min(s["levels_ms"] for s in per_rank), 1)
                    stats["total_ms_max"] = round(max(s["total_ms"] for s in per_rank), 1)
                    digests = [digest(r) for r in (replicas if cfg != "c4" else replicas[:2])]   # c4: 1.5 GB per download, two replicas suffice
                    if reference_digest is None:
                        reference_digest = digests[0]
                    parity = all(x == reference_digest for x in digests)
                for r in replicas:
                    r.close()
            print(json.dumps({"config": cfg, "n_gpus": n, "first_call_s": round(times[0], 4), "build_s": round(min(times[1:]), 4),
                              "all_s": [round(t, 4) for t in times], "stats_rank0_ms": stats, "parity_with_1gpu": parity, "sha": reference_digest}), flush=True)
        S.lib().sdfb200_release_cached_memory()


if __name__ == "__main__":
    main()
