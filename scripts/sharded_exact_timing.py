"""torchrun entry (one process per GPU): ExactOctreeSdf builds of config 4 (M2, depth 8) or config 3 (M1, depth 7) through
sdflib_b200.sharded, warm, with per-rank phase statistics gathered on rank 0.
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sharded_exact_timing.py c4"""
import json, os, sys, time
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
cfg = sys.argv[1] if len(sys.argv) > 1 else "c4"
v, i = meshes.config_mesh("M2" if cfg == "c4" else "M1")
mesh, bb = S.Mesh(v, i), S.BoundingBox(*np.split(meshes.bounding_box_with_margin(v), 2))
times = []
for rep in range(4):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    e = sharded.build_exact_sharded(mesh, bb, 8 if cfg == "c4" else 7, 3, 128, numThreads=2)
    torch.cuda.synchronize(); dist.barrier(); times.append(time.perf_counter() - t0)
    stats = e.build_stats()
    if rep < 3:
        e.close()
all_stats = [None] * world
dist.all_gather_object(all_stats, {k: round(x, 1) for k, x in stats.items() if k.endswith("_ms")})
if rank == 0:
    print(json.dumps({"config": cfg, "mode": "process per GPU", "n_gpus": world, "first_call_s": round(times[0], 4), "build_s": round(min(times[1:]), 4),
                      "levels_ms": [s["levels_ms"] for s in all_stats], "layout_ms": [s["layout_ms"] for s in all_stats]}), flush=True)
dist.barrier()
dist.destroy_process_group()
