#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), ncu launch list, ncu --set full captures of the hot kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gpu_tests.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_exact_build.csv python scripts/profile_kernels.py exact_build > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_octree_build.csv python scripts/profile_kernels.py octree_build > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'filterKernel|sampleKernel' -c 16 -o gpurun_out/exact_build_full -f python scripts/profile_kernels.py exact_build > gpurun_out/ncu_exact_build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'exactQueryKernel' -s 2 -c 1 -o gpurun_out/exact_query_full -f python scripts/profile_kernels.py exact_query > gpurun_out/ncu_exact_query.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'octreeQueryKernel' -s 2 -c 1 -o gpurun_out/octree_query_full -f python scripts/profile_kernels.py octree_query > gpurun_out/ncu_octree_query.log 2>&1
cat gpurun_out/gpu_tests.log
head -c 3000 gpurun_out/bench.json
tail -3 gpurun_out/bench.err
head -c 1500 gpurun_out/bench_ref.json
