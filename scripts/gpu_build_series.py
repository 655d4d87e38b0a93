"""Series of identical builds in one process (the device block cache of host_mem.cpp must settle): public-API seconds and
phase times of 1 + N builds of the Dragon-class ExactOctreeSdf (C4) and of the C2 OctreeSdf builds.
    gpurun --timeout 300 -- 'python scripts/gpu_build_series.py [N]'"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdflib_b200 as S
from sdflib_b200 import meshes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
only_c4 = len(sys.argv) > 2 and sys.argv[2] == "c4"   # with SDFB200_TIMING=1 the per-depth times of every build go to stderr


def series(name, make):
    out = []
    for k in range(n + 1):
        t0 = time.perf_counter()
        s = make()
        dt = time.perf_counter() - t0
        st = s.build_stats()
        out.append("%.3f (total %.0f: levels %.0f, layout %.0f, download %.0f, mesh %.0f)" % (dt, st["total_ms"], st["levels_ms"], st["layout_ms"], st["download_ms"],
                                                                                               st["triangle_data_ms"] + st["bvh_ms"] + st["upload_ms"]))
        s.close()
    print(name, " | ".join(out), flush=True)


if not only_c4:
    v, i = meshes.config_mesh("M1"); box = meshes.bounding_box_with_margin(v)
    mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])
    series("octree_c2", lambda: S.OctreeSdf(mesh, bb, 8, 3, 1e-3, S.OctreeSdf.NO_CONTINUITY, 2))
    series("octree_c2_continuity", lambda: S.OctreeSdf(mesh, bb, 8, 3, 1e-3, S.OctreeSdf.CONTINUITY, 2))
    series("exact_c3", lambda: S.ExactOctreeSdf(mesh, bb, 7, 3, 128, 2))
    S.lib().sdfb200_release_cached_memory()
v, i = meshes.config_mesh("M2"); box = meshes.bounding_box_with_margin(v)
mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])
series("exact_c4", lambda: S.ExactOctreeSdf(mesh, bb, 8, 3, 128, 2))
