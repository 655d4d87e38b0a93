"""A/B of the two late round-2 changes, one process per setting (the switches are read once per process / structure):

  build switches (any ENV=VALUE of the builders)         best-of-4 GPU phase of the C2 builds (both algorithms) + sha1 of the array (must
                                                         not move); profiles/r2_ab_screened_sampler_packed_query.log holds the float-screened
                                                         BVH traversal measured this way (dropped: bvh_sampler.cuh)
  SDFB200_QUERY_PACKED=0|1                               packed float32 instructions in the tile query kernel: kernel times on the
                                                         256^3 grid / 2^24 random points (value, value + gradient) + sha1 of the results

    gpurun --timeout 600 -- 'python scripts/gpu_ab_r2_late.py'
"""
import os, sys, subprocess
if len(sys.argv) > 1 and sys.argv[1] == 'run':
    sys.path.insert(0, '.')
    import hashlib
    import numpy as np
    import torch
    import sdflib_b200 as S
    from sdflib_b200 import meshes
    what = sys.argv[2]
    v, i = meshes.config_mesh("M1"); box = meshes.bounding_box_with_margin(v)
    mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])
    if what == 'build':
        for alg, name in ((S.OctreeSdf.NO_CONTINUITY, 'no_continuity'), (S.OctreeSdf.CONTINUITY, 'continuity')):
            best, total = 1e9, 1e9
            for _ in range(4):
                s = S.OctreeSdf(mesh, bb, 8, 3, 1e-3, alg, 2)
                st = s.build_stats()
                best, total = min(best, st['levels_ms']), min(total, st['total_ms'])
                h = hashlib.sha1(s.getOctreeData().tobytes()).hexdigest()[:12]
                s.close()
            print(sys.argv[3:], name, 'levels_ms %.1f total_ms %.1f' % (best, total), h, flush=True)
    else:
        sdf = S.OctreeSdf(mesh, bb, 8, 3, 1e-3, S.OctreeSdf.NO_CONTINUITY, 1)
        area = sdf.getGridBoundingBox().as_array()
        rng = np.random.default_rng(42)
        sets = {"grid 256^3": torch.from_numpy(meshes.cell_centre_grid(area, 256)).cuda(),
                "random 2^24": torch.from_numpy((area[:3] + rng.random((1 << 24, 3), np.float32) * (area[3:] - area[:3])).astype(np.float32)).cuda()}
        flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
        for name, pts in sets.items():
            for gradient in (False, True):
                dist = torch.empty(pts.shape[0], dtype=torch.float32, device="cuda")
                grad = torch.empty((pts.shape[0], 3), dtype=torch.float32, device="cuda") if gradient else None
                fn = lambda: sdf.getDistance(pts, gradient=gradient, out=dist, out_gradient=grad)
                for _ in range(3):
                    fn()
                best, total = 1e9, 0.0
                for _ in range(20):
                    flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fn(); b.record(); torch.cuda.synchronize()
                    ms = a.elapsed_time(b); best = min(best, ms); total += ms
                h = hashlib.sha1(dist.cpu().numpy().tobytes() + (grad.cpu().numpy().tobytes() if gradient else b'')).hexdigest()[:12]
                print(sys.argv[3:], name, 'grad=%d' % gradient, 'best %.4f ms mean %.4f ms (%.1f Gq/s)' % (best, total / 20, pts.shape[0] / best / 1e6), h, flush=True)
        sdf.close()
else:
    # (the SDFB200_SAMPLE_FAST / SDFB200_SAMPLE_CTAS settings of the committed log existed at commit 12a8917 only)
    runs = [('build', {'SDFB200_SAMPLE_REFILL': '1'}), ('query', {'SDFB200_QUERY_PACKED': '0'}), ('query', {'SDFB200_QUERY_PACKED': '1'})]
    for what, env in runs:
        subprocess.run([sys.executable, __file__, 'run', what] + ['%s=%s' % kv for kv in env.items()], env=dict(os.environ, **env))
