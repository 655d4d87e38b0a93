import sys, time, hashlib
sys.path.insert(0, '.')
import torch
import sdflib_b200 as S
from sdflib_b200 import meshes
v, i = meshes.config_mesh("M1"); box = meshes.bounding_box_with_margin(v)
mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])
for alg, name in ((S.OctreeSdf.NO_CONTINUITY, 'no_continuity'), (S.OctreeSdf.CONTINUITY, 'continuity')):
    for rep in range(4):
        torch.cuda.synchronize(); t = time.perf_counter()
        s = S.OctreeSdf(mesh, bb, 8, 3, 1e-3, alg, 2)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        st = s.build_stats()
        h = hashlib.sha1(s.getOctreeData().tobytes()).hexdigest()[:10]
        print(name, '%.3f s' % dt, {k: round(x, 1) for k, x in st.items() if k.endswith('_ms')}, 'traversals', int(st['leaves']), 'of', int(st['nodes_processed']) * 19, h)
        s.close()
