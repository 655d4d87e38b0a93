"""Device BVH build (bvh_device.cu) against the host builder: wall time of the BVH phase of sdfb200_mesh_create on the
benchmark meshes.   python scripts/gpu_bvh_timing.py [M1 M2 M3]      (SDFB200_HOST_BVH=1 in the environment: host builder)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdflib_b200 as S
from sdflib_b200 import meshes

names = [a for a in sys.argv[1:] if a.startswith("M")] or ["M1", "M2"]
for name in names:
    v, i = meshes.config_mesh(name)
    mesh = S.Mesh(v, i)
    rows = []
    for rep in range(4):
        t0 = time.perf_counter()
        pm = S.PreparedMesh(mesh, bvh=True, exact=False)
        wall = time.perf_counter() - t0
        st = pm.stats()
        rows.append({"wall_ms": round(wall * 1e3, 2), **{k: round(x, 2) for k, x in st.items()}})
        pm.close()
    print(json.dumps({"mesh": name, "triangles": int(i.size // 3), "host_bvh": bool(os.environ.get("SDFB200_HOST_BVH")), "runs": rows}))
