"""A/B of the BVH traversal scheduling variants on the C2 OctreeSdf build (levels_ms = GPU phase)."""
import os, sys, subprocess, json
if len(sys.argv) > 1:
    sys.path.insert(0, '.')
    import sdflib_b200 as S
    from sdflib_b200 import meshes
    v, i = meshes.config_mesh("M1"); box = meshes.bounding_box_with_margin(v)
    mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])
    best = None
    for _ in range(3):
        s = S.OctreeSdf(mesh, bb, 8, 3, 1e-3, S.OctreeSdf.NO_CONTINUITY, 2)
        st = s.build_stats(); best = st['levels_ms'] if best is None else min(best, st['levels_ms'])
        import hashlib; h = hashlib.sha1(s.getOctreeData().tobytes()).hexdigest()[:10]
        s.close()
    print('variant', os.environ.get('SDFB200_BVH_VARIANT'), 'levels_ms %.1f' % best, h)
else:
    for var in ('x',):
        subprocess.run([sys.executable, __file__, 'run'], env=dict(os.environ, SDFB200_BVH_VARIANT=var))
