#!/usr/bin/env python
"""Headline benchmark: SDF queries/sec on a 256^3 grid (+ octree build seconds), B200 vs the reference's CPU path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]; SURVEY.md §8d "C2"): Armadillo-class mesh M1 (displaced icosphere,
327 680 triangles, synthetic — the real scan is not available offline), OctreeSdf(depth=8, startDepth=3,
threshold=1e-3, NO_CONTINUITY), bulk getDistance over the 256^3 cell centres of the octree box.
A step = one pass of getDistance over the whole 16.7 M-point grid.
  value : queries/s with points, structure and outputs resident in HBM (kernel launched through the C-ABI
          with device pointers, timed with CUDA events on the launching stream, max over ranks)
  e2e   : the same through the C-ABI with HOST buffers (pinned): H2D of the points and D2H of the distances
          inside the timed region
  N > 1 : weak scaling — the structure is replicated, every rank queries its own 256^3 grid, no data-path
          collective; value = total queries of all ranks / max-over-ranks time
The working set (201 MB points + 67 MB distances + 82.5 MB structure) exceeds the 126 MB L2, so successive
steps cannot be served from cache ("inputs larger than L2").

Besides the headline line the same JSON object carries the other half of BASELINE.json's metric ("octree build sec"):
  build.octree_c2 / build.octree_c2_continuity / build.exact_c3 : wall-clock seconds of the OctreeSdf (C2, NO_CONTINUITY and
          CONTINUITY) and ExactOctreeSdf (C3: depth 7, minTri 128) constructors through the public API; with N > 1 the
          builds are cooperative (sdflib_b200.sharded: start-depth voxels sharded + one NCCL all-gather of the payload;
          CONTINUITY: sampling sliced over the ranks + one all-gather per depth), seconds = max over ranks
  exact_query : ExactOctreeSdf bulk getDistance over the same 256^3 grid, device-resident

--impl reference times the UNMODIFIED reference (oracle/_ref/libsdfref.so: OctreeSdf built by its own
OpenMP builder, getDistance driven from an `omp parallel for` over all host threads) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(mesh="M1 displaced icosphere, 327680 triangles (Armadillo-class, synthetic)", depth=8, start_depth=3,
                threshold=1e-3, algorithm="NO_CONTINUITY", grid=256)
EXACT = dict(depth=7, start_depth=3, min_triangles=128)   # BASELINE.json configs[2] ("C3"), same mesh
METRIC = "sdf_queries_per_sec_256cubed_grid"
UNIT = "queries/s"


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture of this same command (profiles/query_kernel_traffic.json); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "query_kernel_traffic.json")
    try:
        return float(json.load(open(p))["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the benchmark runs (B200_PROFILING.md clocks line)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first(self, timeout=5.0):
        """nvidia-smi needs a few hundred ms to start; the timed region (tens of ms) must not begin before it samples."""
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.06)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 - 0.05 <= ts <= t1 + 0.05]
        if not rows and self.rows:   # region shorter than the sampling period: the samples closest to it (warm-up runs the same kernel)
            mid = 0.5 * (t0 + t1)
            rows = [r for _, r in sorted(self.rows, key=lambda x: abs(x[0] - mid))[:3]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            c = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def build_inputs(rank=0):
    from sdflib_b200 import meshes
    v, i = meshes.config_mesh("M1")
    box = meshes.bounding_box_with_margin(v)
    return v, i, box


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.binding import ref
    from sdflib_b200 import meshes
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libsdfref.so is not present"}))
        return
    cores = os.cpu_count() or 1
    v, i, box = build_inputs()
    sdf = ref.build_octree(v, i, box, WORKLOAD["depth"], WORKLOAD["start_depth"], WORKLOAD["threshold"], 1, max(cores, 2))
    build_s = sdf.build_seconds
    exact_build_s = None
    if not args.no_exact_reference:
        ex = ref.build_exact(v, i, box, EXACT["depth"], EXACT["start_depth"], EXACT["min_triangles"], max(cores, 2))
        exact_build_s = ex.build_seconds
        ex.close()
    cont_build_s = None
    if not args.no_continuity_reference:
        co = ref.build_octree(v, i, box, WORKLOAD["depth"], WORKLOAD["start_depth"], WORKLOAD["threshold"], 2, max(cores, 2))
        cont_build_s = co.build_seconds
        co.close()
    area = sdf.sample_area()
    # bounded sample of the workload: every 8th point of the 256^3 cell-centre grid (2.1 M queries) per step
    n = WORKLOAD["grid"]
    pts = np.ascontiguousarray(meshes.cell_centre_grid(area, n)[::8])
    for _ in range(max(args.warmup, 1)):
        sdf.query(pts[: len(pts) // 8], num_threads=cores)
    t = 0.0
    for _ in range(args.steps):
        sdf.query(pts, num_threads=cores)
        t += sdf.last_query_seconds
    value = len(pts) * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2: " + json.dumps(WORKLOAD), "sample": "every 8th point of the 256^3 grid per step"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": f"every 8th point of the 256^3 grid ({len(pts)} queries) x {args.steps} steps, omp parallel for over getDistance",
                             "build_s": build_s, "build_threads": max(cores, 2)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "build": {"octree_c2": {"seconds": build_s, "threads": max(cores, 2)},
                      "octree_c2_continuity": {"seconds": cont_build_s, "threads": max(cores, 2)},
                      "exact_c3": {"seconds": exact_build_s, "threads": max(cores, 2), "config": EXACT}},
            "build_s": build_s}
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import sdflib_b200 as S
    from sdflib_b200 import meshes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    S.lib().sdfb200_set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # NCCL prints its version banner on stdout; stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    v, i, box = build_inputs()
    mesh, bb = S.Mesh(v, i), S.BoundingBox(box[:3], box[3:])

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed_build(make):
        """(object, seconds): wall clock around the constructor, barrier on both sides, max over ranks."""
        sync_all()
        t0 = time.perf_counter()
        obj = make()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return obj, float(t.item())

    if world > 1:
        from sdflib_b200 import sharded
        make_oct = lambda: sharded.build_octree_sharded(mesh, bb, WORKLOAD["depth"], WORKLOAD["start_depth"], WORKLOAD["threshold"], numThreads=2)
        make_ex = lambda: sharded.build_exact_sharded(mesh, bb, EXACT["depth"], EXACT["start_depth"], EXACT["min_triangles"], numThreads=2)
    else:
        make_oct = lambda: S.OctreeSdf(mesh, bb, WORKLOAD["depth"], WORKLOAD["start_depth"], WORKLOAD["threshold"], S.OctreeSdf.NO_CONTINUITY, 2)
        make_ex = lambda: S.ExactOctreeSdf(mesh, bb, EXACT["depth"], EXACT["start_depth"], EXACT["min_triangles"], 2)
    sdf, build_first_s = timed_build(make_oct)
    sdf.close()
    builds = []
    for _ in range(3):
        sdf, t = timed_build(make_oct)
        builds.append(t)
        if _ < 2:
            sdf.close()
    build_s = min(builds)
    stats = sdf.build_stats()
    info = sdf.info()
    cont = None
    if True:   # InitAlgorithm::CONTINUITY (the reference's CLI / Unity default), same mesh and depth
        if world > 1:   # replicated logic, BVH sampling sliced over the ranks, one all-gather per depth
            make_cont = lambda: sharded.build_octree_sharded(mesh, bb, WORKLOAD["depth"], WORKLOAD["start_depth"], WORKLOAD["threshold"],
                                                             initAlgorithm=S.OctreeSdf.CONTINUITY)
        else:
            make_cont = lambda: S.OctreeSdf(mesh, bb, WORKLOAD["depth"], WORKLOAD["start_depth"], WORKLOAD["threshold"], S.OctreeSdf.CONTINUITY, 2)
        c, _t = timed_build(make_cont)
        c.close()
        cont_builds = []
        for _ in range(3):
            c, t = timed_build(make_cont)
            cont_builds.append(t)
            if _ == 2:
                cont = {"seconds": min(cont_builds), "all_seconds": cont_builds, "stats_ms_rank0": c.build_stats(),
                        "octree_words": int(c.info().octree_words)}
            c.close()
    exact, _t = timed_build(make_ex)
    exact.close()
    ex_builds = []
    for _ in range(3):
        exact, t = timed_build(make_ex)
        ex_builds.append(t)
        if _ < 2:
            exact.close()
    exact_build_s = min(ex_builds)
    exact_stats = exact.build_stats()
    exact_info = exact.info()
    area = sdf.getSampleArea().as_array()
    n = WORKLOAD["grid"]
    host_pts = meshes.cell_centre_grid(area, n)
    if world > 1:   # every rank its own grid: shift by a rank-dependent fraction of a cell (still inside the box)
        host_pts = (host_pts + np.float32(0.25 * rank / world) * (area[3:] - area[:3]) / np.float32(n)).astype(np.float32)
    nq = len(host_pts)
    pts = torch.from_numpy(host_pts).cuda()
    out = torch.empty(nq, dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        sdf.getDistance(pts, out=out)
    if sampler:
        sampler.wait_first()
        for _ in range(max(args.warmup, 20)):   # keep the GPU under the same load while the first samples come in
            sdf.getDistance(pts, out=out)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        sdf.getDistance(pts, out=out)
    e1.record(stream)
    barrier()
    wall1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(wall0, wall1) if sampler else None
    checksum = float(out.double().sum().item())

    # e2e: host (pinned) buffers through the C-ABI, copies inside the timed region
    pin_pts = torch.from_numpy(host_pts).pin_memory()
    pin_out = torch.empty(nq, dtype=torch.float32).pin_memory()
    np_pts, np_out = pin_pts.numpy(), pin_out.numpy()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        sdf.getDistance(np_pts, out=np_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sdf.getDistance(np_pts, out=np_out)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * nq * e2e_steps / float(e2e_s.item())
    e2e_ok = bool(np.array_equal(np_out, out.cpu().numpy()))

    # ExactOctreeSdf bulk queries on the same grid (device-resident), a handful of steps
    ex_out = torch.empty(nq, dtype=torch.float32, device="cuda")
    ex_steps = max(3, min(args.steps, 10))
    for _ in range(3):
        exact.getDistance(pts, out=ex_out)
    barrier()
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x0.record(stream)
    for _ in range(ex_steps):
        exact.getDistance(pts, out=ex_out)
    x1.record(stream)
    barrier()
    ex_ms = torch.tensor([x0.elapsed_time(x1)], device="cuda")
    if world > 1:
        dist.all_reduce(ex_ms, op=dist.ReduceOp.MAX)
    ex_ms_step = float(ex_ms.item()) / ex_steps
    # the exact field and the tri-cubic approximation of the same mesh must agree to the octree's error threshold
    approx_vs_exact = float((out - ex_out).abs().max().item())

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        ms_step = ms_total / args.steps
        algo_bytes = nq * 16 + 4 * info.octree_words   # N_q*(12+4) + 4*#mOctreeData, per launch
        achieved = algo_bytes / (ms_step * 1e-3) / 1e9
        line = {"metric": METRIC, "value": world * nq * args.steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C2: " + json.dumps(WORKLOAD), "queries_per_step_per_gpu": nq,
                           "octree_words": int(info.octree_words), "l2": "inputs larger than L2 (350 MB working set)",
                           "parallelism": f"replicated structure, {world} independent query shards"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nq * 12), "d2h_bytes_per_step": int(nq * 4),
                        "steps": e2e_steps, "matches_device_run": e2e_ok},
                "gpu_launches": args.steps,   # one octreeQueryKernel launch per timed step (value region)
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": measured_traffic(), "peak_source": peak_src, "kernel": "octreeQueryKernel<false>",
                             "algorithmic_bytes_per_launch": int(algo_bytes)},
                "build": {"scaling": "strong (one mesh, start-depth voxels sharded over the ranks)" if world > 1 else "single GPU",
                          "octree_c2": {"seconds": build_s, "first_call_seconds": build_first_s, "all_seconds": builds, "stats_ms_rank0": stats},
                          "octree_c2_continuity": cont,
                          "exact_c3": {"seconds": exact_build_s, "all_seconds": ex_builds, "config": EXACT, "stats_ms_rank0": exact_stats,
                                       "nodes": int(exact_info.octree_words), "set_words": int(exact_info.triangle_sets_words),
                                       "mask_bytes": int(exact_info.triangle_masks_bytes),
                                       "max_triangles_in_leafs": int(exact_info.max_triangles_in_leafs)}},
                "exact_query": {"value": world * nq / (ex_ms_step * 1e-3), "unit": UNIT, "ms_per_step": ex_ms_step, "steps": ex_steps,
                                "max_abs_difference_to_octree_sdf": approx_vs_exact},
                "build_s": build_s,
                "clocks": clocks, "checksum": checksum}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(sdf, area)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(sdf, area):
    """The reference's own getDistance (oracle/_ref) on the host cores, on a bounded sample of the same workload:
    the structure built on the GPU is handed over through the .bin format, the reference loads and queries it."""
    from oracle.binding import ref, port
    from sdflib_b200 import meshes
    backend = ref if ref.available() else port
    cores = os.cpu_count() or 1
    path = f"/tmp/sdfb200_bench_{os.getpid()}.bin"
    sdf.saveToFile(path)
    r = backend.load(path)
    os.remove(path)
    pts = np.ascontiguousarray(meshes.cell_centre_grid(area, WORKLOAD["grid"])[::8])
    r.query(pts[: len(pts) // 8], num_threads=cores)
    t, reps = 0.0, 3
    for _ in range(reps):
        r.query(pts, num_threads=cores)
        t += r.last_query_seconds
    r.query(pts[: len(pts) // 4], num_threads=1)
    single = (len(pts) // 4) / r.last_query_seconds
    return {"value": len(pts) * reps / t, "unit": UNIT, "cores": cores, "kind": backend.kind,
            "sample": f"every 8th point of the 256^3 grid ({len(pts)} queries) x {reps}, external omp parallel for over "
                      f"getDistance on all cores; structure = GPU-built .bin loaded by the reference",
            "single_thread_value": single}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-continuity-reference", action="store_true", help="reference arm: skip the CONTINUITY OctreeSdf build")
    ap.add_argument("--no-exact-reference", action="store_true", help="reference arm: skip the (tens of seconds) ExactOctreeSdf build")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
