#!/bin/bash
# round 2: ncu --set full of the OctreeSdf query kernels (plain / cooperative, grid / random points)
mkdir -p gpurun_out
P=scripts/profile_kernels.py
NCU="timeout 600 ncu --clock-control none --set full --import-source on"
SDFB200_QUERY_COOP=1 $NCU -k regex:octreeQueryCoopKernel -s 2 -c 1 -o gpurun_out/r2_query_coop_grid -f python $P octree_query > gpurun_out/ncu_q1.log 2>&1
SDFB200_QUERY_COOP=1 $NCU -k regex:octreeQueryCoopKernel -s 2 -c 1 -o gpurun_out/r2_query_coop_random -f python $P octree_query_random > gpurun_out/ncu_q2.log 2>&1
$NCU -k regex:octreeQueryKernel -s 2 -c 1 -o gpurun_out/r2_query_plain_random -f python $P octree_query_random > gpurun_out/ncu_q3.log 2>&1
tail -2 gpurun_out/ncu_q*.log
