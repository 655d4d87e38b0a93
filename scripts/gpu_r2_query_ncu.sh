#!/bin/bash
# round 2: ncu --set full of the OctreeSdf query kernels (tile = default, plain = SDFB200_QUERY_PLAIN=1; grid / random points)
mkdir -p gpurun_out
P=scripts/profile_kernels.py
NCU="timeout 600 ncu --clock-control none --set full --import-source on"
$NCU -k regex:octreeQueryTileKernel -s 2 -c 1 -o gpurun_out/r2_query_tile_grid -f python $P octree_query > gpurun_out/ncu_q1.log 2>&1
$NCU -k regex:octreeQueryTileKernel -s 2 -c 1 -o gpurun_out/r2_query_tile_random -f python $P octree_query_random > gpurun_out/ncu_q2.log 2>&1
tail -n 2 gpurun_out/ncu_q*.log
