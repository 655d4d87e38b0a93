"""torchrun entry: per-depth phase times (SDFB200_TIMING=1) of the collective CONTINUITY build on the C2 mesh."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import sdflib_b200 as S
from sdflib_b200 import meshes, sharded
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); S.lib().sdfb200_set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
v, i = meshes.config_mesh("M1"); box = meshes.bounding_box_with_margin(v)
mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])
for rep in range(3):
    if rank == 0: print("=== build", rep, file=sys.stderr, flush=True)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    c = sharded.build_octree_sharded(mesh, bb, 8, 3, 1e-3, initAlgorithm=S.OctreeSdf.CONTINUITY)
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0: print("=== total %.3f s" % (time.perf_counter() - t0), {k: round(x, 1) for k, x in c.build_stats().items() if k.endswith('_ms')}, file=sys.stderr, flush=True)
    c.close()
dist.destroy_process_group()
