"""Generates tests/golden/config4.npz: BASELINE configs[3] (Dragon-class mesh M2 = getIsosphere(9) + displacement,
5 242 880 triangles; ExactOctreeSdf depth 8, start depth 3, minTrianglesPerNode 128) built by the HISTORY-FREE CPU oracle
(port, use_cache=0, single thread: about an hour) — sha256 and sizes of its arrays, and its exact distances on a sample
of the 256^3 grid. The GPU build must reproduce the hashes (tests/test_gpu_full_size.py::test_config4_*).
    python tests/golden/make_golden_c4.py"""
import hashlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle.binding import ref, port                       # noqa: E402
from make_golden import displace, box_of, sha              # noqa: E402
from make_golden_full import exact_arrays, grid_sample     # noqa: E402

v, i = ref.isosphere(9)
v = displace(v)
box = box_of(v)
t = time.time()
p = port.build_exact(v, i, box, 8, 3, 128, 1, use_cache=False)
nodes = p.octree_data()
sets, masks, tris = exact_arrays(p)
print("port build", round(time.time() - t), "s", nodes.size // 2, sets.size, masks.size, flush=True)
q = grid_sample(p.sample_area())[::16]
d, g = p.query(q, True, os.cpu_count())
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "config4.npz"), box=box, mesh_sha256=sha(v) + sha(i),
                    nodes=np.int64(nodes.size // 2), nodes_sha256=sha(nodes), sets_words=np.int64(sets.size), sets_sha256=sha(sets),
                    masks_bytes=np.int64(masks.size), masks_sha256=sha(masks), triangle_data_sha256=sha(tris),
                    sample_every=np.int64(16), distances=d, gradients=g)
print("done", flush=True)
