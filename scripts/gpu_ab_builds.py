"""A/B of a build-path switch given as ENV=VALUE pairs on the command line: best-of-4 GPU phase of the C2 builds + output hash."""
import os, sys, subprocess
if len(sys.argv) > 1 and sys.argv[1] == 'run':
    sys.path.insert(0, '.')
    import hashlib, time, torch
    import sdflib_b200 as S
    from sdflib_b200 import meshes
    v, i = meshes.config_mesh("M1"); box = meshes.bounding_box_with_margin(v)
    mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])
    for alg, name in ((S.OctreeSdf.NO_CONTINUITY, 'no_continuity'), (S.OctreeSdf.CONTINUITY, 'continuity')):
        best = 1e9
        for _ in range(4):
            s = S.OctreeSdf(mesh, bb, 8, 3, 1e-3, alg, 2)
            best = min(best, s.build_stats()['levels_ms'])
            h = hashlib.sha1(s.getOctreeData().tobytes()).hexdigest()[:10]
            s.close()
        print(sys.argv[2:], name, 'levels_ms %.1f' % best, h, flush=True)
else:
    for setting in sys.argv[1:]:
        k, v = setting.split('=')
        subprocess.run([sys.executable, __file__, 'run', setting], env=dict(os.environ, **{k: v}))
