"""torchrun entry: sharded OctreeSdf / ExactOctreeSdf builds over NCCL must equal the single-rank builds bit for bit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import sdflib_b200 as S
from sdflib_b200 import meshes, sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
S.lib().sdfb200_set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "s4"
if name == "s4":
    v, i = meshes.isosphere(4); v = meshes.displace(v)
else:
    v, i = meshes.config_mesh(name)
box = meshes.bounding_box_with_margin(v)
mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])
od, ed = (6, 6) if name == "s4" else ((8, 8) if name == "M2" else (8, 7))   # M2: BASELINE config 4 (ExactOctreeSdf depth 8)
for rep in range(2):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    o = sharded.build_octree_sharded(mesh, bb, od, 3, 1e-3, numThreads=2)
    torch.cuda.synchronize(); dist.barrier(); t_oct = time.perf_counter() - t0
    t0 = time.perf_counter()
    e = sharded.build_exact_sharded(mesh, bb, ed, 3, 128 if name != "s4" else 32, numThreads=2)
    torch.cuda.synchronize(); dist.barrier(); t_ex = time.perf_counter() - t0
# InitAlgorithm::CONTINUITY: replicated logic, BVH sampling sliced over the ranks and all-gathered per depth
for rep in range(2):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    c = sharded.build_octree_sharded(mesh, bb, od, 3, 1e-3, initAlgorithm=S.OctreeSdf.CONTINUITY)
    torch.cuda.synchronize(); dist.barrier(); t_cont = time.perf_counter() - t0
    cont_stats = c.build_stats()
if rank == 0:
    t0 = time.perf_counter(); c1 = S.OctreeSdf(mesh, bb, od, 3, 1e-3, S.OctreeSdf.CONTINUITY, 2); t_c1 = time.perf_counter() - t0
    assert np.array_equal(c.getOctreeData(), c1.getOctreeData())
    print(f"collective CONTINUITY ok world={world} {t_cont:.3f}s (single {t_c1:.3f}s) levels {cont_stats['levels_ms']:.1f} ms (single {c1.build_stats()['levels_ms']:.1f} ms)", flush=True)
    t0 = time.perf_counter(); o1 = S.OctreeSdf(mesh, bb, od, 3, 1e-3, S.OctreeSdf.NO_CONTINUITY, 2); t_o1 = time.perf_counter() - t0
    t0 = time.perf_counter(); e1 = S.ExactOctreeSdf(mesh, bb, ed, 3, 128 if name != "s4" else 32, 2); t_e1 = time.perf_counter() - t0
    assert np.array_equal(o.getOctreeData(), o1.getOctreeData())
    assert np.array_equal(e.getOctreeData(), e1.getOctreeData())
    assert np.array_equal(e.getTrianglesSets(), e1.getTrianglesSets()) and np.array_equal(e.getTrianglesMasks(), e1.getTrianglesMasks())
    print(f"sharded ok world={world} octree {t_oct:.3f}s (single {t_o1:.3f}s) exact {t_ex:.3f}s (single {t_e1:.3f}s)", flush=True)
dist.barrier()
dist.destroy_process_group()
