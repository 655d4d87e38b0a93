"""GPU vs oracle diagnosis of the CONTINUITY build (first differing word, per-level sizes)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import sdflib_b200 as S
from sdflib_b200 import meshes
from oracle.binding import port

cases = [(2, 5, 3, 1e-3), (2, 4, 0, 1e-3), (3, 6, 3, 1e-3), (3, 6, 1, 2e-3), (4, 7, 3, 3e-4)]
for sub, depth, start, thr in cases:
    v, i = meshes.isosphere(sub); v = meshes.displace(v); box = meshes.bounding_box_with_margin(v)
    t = time.time()
    try:
        g = S.OctreeSdf(S.Mesh(v, i), S.BoundingBox(box[:3], box[3:]), depth, start, thr, S.OctreeSdf.CONTINUITY, 1)
    except Exception as e:
        print((sub, depth, start), 'GPU build failed:', e); continue
    tg = time.time() - t
    t = time.time()
    p = port.build_octree(v, i, box, depth, start, thr, 2, 1, use_cache=False)
    tp = time.time() - t
    a, b = g.getOctreeData(), p.octree_data()
    print((sub, depth, start, thr), 'gpu words', a.size, 'oracle words', b.size, 'gpu %.3fs oracle %.3fs' % (tg, tp))
    n = min(a.size, b.size)
    diff = np.nonzero(a[:n] != b[:n])[0]
    if diff.size == 0 and a.size == b.size:
        print('   IDENTICAL', g.info().value_range, p.header())
        continue
    print('   differing words', diff.size, 'first', diff[:8], 'gpu', a[diff[:8]], 'oracle', b[diff[:8]])
    fa, fb = a[:n].view(np.float32)[diff], b[:n].view(np.float32)[diff]
    print('   as float: max abs diff', np.nanmax(np.abs(fa - fb)))
