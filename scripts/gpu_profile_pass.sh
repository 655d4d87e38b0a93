#!/bin/bash
# ncu evidence pass, round 2 (one GPU): launch lists of the bench command and of the builds at HEAD.
# (--set full captures of the hot kernels: scripts/gpu_r2_query_ncu.sh and the history of this file.)
mkdir -p gpurun_out
P=scripts/profile_kernels.py
NCU="timeout 900 ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config4 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_octree_build_c2.csv python $P octree_build > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_octree_cont_c2.csv python $P octree_cont > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_bvh_m1.csv python scripts/gpu_bvh_timing.py M1 > /dev/null 2>&1
ls -la gpurun_out/r2_launches_*.csv
