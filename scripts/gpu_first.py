"""First GPU check: kernel-level parity, octree build parity vs the oracles, query parity, rough timings."""
import ctypes as C, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sdflib_b200 as S
from sdflib_b200 import _capi, meshes
from oracle.binding import ref, port

out = {}
def eq(a, b): return bool(np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32)))

print("device", torch.cuda.get_device_name(0), "count", S.device_count(), flush=True)
L = _capi.lib()
rng = np.random.default_rng(0)
v, i = meshes.isosphere(4); v = meshes.displace(v)
# kernels
pts = rng.uniform(-2, 2, (200000, 3)).astype(np.float32)
tri = np.empty(len(pts), np.uint32)
_capi.check(L.sdfb200_nearest_triangle(_capi.ptr(v), C.c_uint32(len(v)), _capi.ptr(i), C.c_uint32(i.size), _capi.ptr(pts), C.c_uint64(len(pts)), _capi.ptr(tri)))
rt = port.nearest_triangle(v, i, pts)
print("bvh nearest equal:", np.array_equal(tri, rt), (tri != rt).sum(), flush=True)
td = port.triangle_data(v, i)
for t in (0, 17, 333):
    w = v[i[3*t:3*t+3]].reshape(-1).copy()
    for mode in (0, 1, 2, 3):
        d = np.empty(len(pts), np.float32); g = np.zeros((len(pts), 3), np.float32)
        _capi.check(L.sdfb200_point_triangle(_capi.ptr(td[t].copy()), _capi.ptr(w), _capi.ptr(pts), C.c_uint64(len(pts)), C.c_int(mode), _capi.ptr(d), _capi.ptr(g)))
        if mode == 3: rd = port.sq_dist(td[t], pts); rg = g
        else: rd, rg = port.signed_dist(td[t], w, pts, mode)
        assert eq(d, rd) and eq(g, rg), (t, mode, np.abs(d-rd).max())
print("point-triangle kernels bit-exact", flush=True)

def run(name, sub, disp, depth, start, thr, with_ref=True, nthreads=1):
    v, i = meshes.isosphere(sub)
    if disp: v = meshes.displace(v)
    box = meshes.bounding_box_with_margin(v)
    mesh = S.Mesh(v, i); bb = S.BoundingBox(box[:3], box[3:])
    t = time.time(); sdf = S.OctreeSdf(mesh, bb, depth, start, thr, S.OctreeSdf.NO_CONTINUITY, nthreads); tg = time.time() - t
    t = time.time(); sdf2 = S.OctreeSdf(mesh, bb, depth, start, thr, S.OctreeSdf.NO_CONTINUITY, nthreads); tg2 = time.time() - t
    st = sdf2.build_stats()
    data = sdf.getOctreeData()
    res = dict(tris=i.size // 3, words=int(data.size), gpu_build_s=tg, gpu_build2_s=tg2, stats=st)
    print(name, res, flush=True)
    if with_ref:
        t = time.time(); p = port.build_octree(v, i, box, depth, start, thr, 1, nthreads, use_cache=False); tp = time.time() - t
        pd = p.octree_data()
        res["port_nocache_equal"] = bool(pd.size == data.size and np.array_equal(pd, data))
        res["port_s"] = tp
        if pd.size == data.size: res["words_differing"] = int((pd != data).sum())
        hp = p.header(); info = sdf.info()
        res["header"] = (hp, info.value_range, info.min_border_value)
        res["header_equal"] = bool(np.float32(hp["value_range"]) == np.float32(info.value_range) and np.float32(hp["min_border_value"]) == np.float32(info.min_border_value))
        r = ref.build_octree(v, i, box, depth, start, thr, 1, nthreads); rd = r.octree_data()
        res["ref_s"] = r.build_seconds
        res["ref_size_equal"] = bool(rd.size == data.size)
        # queries
        area = sdf.getSampleArea().as_array()
        q = (area[:3] + rng.uniform(-0.1, 1.1, (300000, 3)) * (area[3:] - area[:3])).astype(np.float32)
        dg, gg = sdf.getDistance(q, gradient=True, exact_order=True)
        dp, gp = p.query(q, True)
        res["query_exact_equal"] = (eq(dg, dp), eq(gg, gp))
        df, gf = sdf.getDistance(q, gradient=True)
        res["query_fast_maxabs"] = (float(np.abs(df - dp).max()), float(np.abs(gf - gp).max()))
        dr = r.query(q)
        res["vs_ref_maxabs"] = float(np.abs(df - dr).max())
        print("   ", {k: res[k] for k in res if k not in ("stats",)}, flush=True)
    # timing of device-resident queries
    N = 256
    grid = torch.from_numpy(meshes.cell_centre_grid(sdf.getSampleArea().as_array(), N)).cuda()
    o = torch.empty(len(grid), device="cuda")
    for exact in (False, True):
        for gradient in (False, True):
            og = torch.empty((len(grid), 3), device="cuda") if gradient else None
            for _ in range(3): sdf.getDistance(grid, gradient=gradient, exact_order=exact, out=o, out_gradient=og)
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): sdf.getDistance(grid, gradient=gradient, exact_order=exact, out=o, out_gradient=og)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            res[f"q256_exact{int(exact)}_grad{int(gradient)}_Gqps"] = len(grid) / ms / 1e6
    print("   query", {k: round(vv, 3) for k, vv in res.items() if k.startswith("q256")}, flush=True)
    out[name] = res

run("C1", 2, False, 5, 3, 1e-3)
run("C1mt", 2, False, 5, 3, 1e-3, nthreads=2)
run("s4", 4, True, 6, 2, 1e-3)
run("s5", 5, True, 7, 3, 1e-3)
run("C2", 7, True, 8, 3, 1e-3, with_ref=False)
json.dump(out, open("gpurun_out/gpu_first.json", "w"), indent=1, default=str)
