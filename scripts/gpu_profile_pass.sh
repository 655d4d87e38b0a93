#!/bin/bash
# ncu evidence pass (one GPU): launch lists of the three builds + the bench, --set full captures of the hot kernels.
mkdir -p gpurun_out
P=scripts/profile_kernels.py
NCU="timeout 900 ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_octree_build.csv python $P octree_build > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_octree_cont.csv python $P octree_cont > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_exact_build.csv python $P exact_build > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$NCU --set full --import-source on -k regex:sampleOwnersKernel -s 6 -c 1 -o gpurun_out/sample_owners_full -f python $P octree_build > gpurun_out/ncu1.log 2>&1
$NCU --set full --import-source on -k regex:'filterRefillKernel|sampleKernel|compactKernel' -s 15 -c 9 -o gpurun_out/exact_build_full -f python $P exact_build > gpurun_out/ncu2.log 2>&1
$NCU --set full --import-source on -k regex:'exactBinnedKernel|exactWalkKernel|exactScatterKernel' -s 6 -c 3 -o gpurun_out/exact_query_full -f python $P exact_query > gpurun_out/ncu3.log 2>&1
$NCU --set full --import-source on -k regex:octreeQueryKernel -s 2 -c 1 -o gpurun_out/octree_query_full -f python $P octree_query > gpurun_out/ncu4.log 2>&1
$NCU --set full --import-source on -k regex:'contDecideKernel|contJunctionKernel|contEmitKernel|fixValuesKernel|samplePointsKernel|dedupeInsertKernel' -s 20 -c 10 -o gpurun_out/octree_cont_full -f python $P octree_cont > gpurun_out/ncu5.log 2>&1
tail -2 gpurun_out/ncu*.log
ls -la gpurun_out/
