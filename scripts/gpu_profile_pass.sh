#!/bin/bash
# ncu evidence pass, round 2 (one GPU): launch lists of the three builds + the bench command, --set full captures of the hot kernels.
mkdir -p gpurun_out
P=scripts/profile_kernels.py
NCU="timeout 900 ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_octree_build_c2.csv python $P octree_build > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_octree_cont_c2.csv python $P octree_cont > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_exact_build_c3.csv python $P exact_build > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config4 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:'trianglePassKernel|edgePairKernel|vertexNormalKernel|toFrameKernel|DeviceRadixSort' -c 12 -o gpurun_out/r2_mesh_ingest_full -f python $P exact_build > gpurun_out/ncu_r2_1.log 2>&1
$NCU --set full --import-source on -k regex:sampleOwnersRefillKernel -s 6 -c 1 -o gpurun_out/r2_sample_refill_full -f python $P octree_build > gpurun_out/ncu_r2_2.log 2>&1
$NCU --set full --import-source on -k regex:'filterRefillKernel|sampleKernel' -s 8 -c 6 -o gpurun_out/r2_exact_build_full -f python $P exact_build > gpurun_out/ncu_r2_3.log 2>&1
$NCU --set full --import-source on -k regex:octreeQueryTileKernel -s 2 -c 1 -o gpurun_out/r2_query_tile_grad_full -f python $P octree_query_grad > gpurun_out/ncu_r2_4.log 2>&1
tail -n 2 gpurun_out/ncu_r2_*.log
ls -la gpurun_out/ | tail -20
