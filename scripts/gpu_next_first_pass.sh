#!/bin/bash
# First GPU call of the next session: everything that was written after the last GPU minute of round 1 and is still
# unmeasured, in one box (about 6-8 minutes):
#   1. parity tests of HEAD (the default paths must still be green: host BVH builder and C-ABI query entry changed),
#   2. BVH build time on the B200 host with the threaded std::sort (was 57 ms inside the C2 build),
#   2b. random non-manifold meshes against the oracle (tests/fuzz_gpu_parity.py; promote to a gpu test once green),
#   3. A/B of the lane-refill BVH sampler (hash of the octree must not change),
#   4. A/B of the two experimental query variants (bit identity / tolerance + kernel times),
#   5. the quick bench, for the new baseline of the build phases.
# Everything lands in gpurun_out/next_*.log.
#     gpurun --timeout 1500 -- 'bash scripts/gpu_next_first_pass.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/next_gpu_tests.log

NVCC=$(command -v nvcc || echo /usr/local/cuda/bin/nvcc)
$NVCC -ccbin /usr/bin/g++ -x cu -std=c++17 -O2 -Wno-deprecated-gpu-targets -Iinclude -Isdflib_b200/csrc tests/cpp/bvh_host_main.cpp \
      -o /tmp/bvh_host_main -Lsdflib_b200 -lsdfb200 -Xlinker -rpath,$PWD/sdflib_b200 \
  && { echo "bvh_host_main <subdivisions> <displaced> -> ok <nodes> <serial reference ms> <product ms, best of 5>"; nproc;
       /tmp/bvh_host_main 7 1; SDFB200_TIMING=1 /tmp/bvh_host_main 7 1 2>&1 | tail -4; /tmp/bvh_host_main 9 1; } 2>&1 | tee gpurun_out/next_bvh_host.log

timeout 600 python tests/fuzz_gpu_parity.py 40 2>&1 | tail -12 | tee gpurun_out/next_fuzz_parity.log
timeout 600 python scripts/gpu_ab_builds.py SDFB200_SAMPLE_REFILL=0 SDFB200_SAMPLE_REFILL=1 2>&1 | tee gpurun_out/next_ab_refill.log
timeout 900 python scripts/gpu_ab_query_variants.py 2>&1 | tee gpurun_out/next_ab_query.log
SKIP_TESTS=1 bash scripts/gpu_quick.sh 2>&1 | tail -40 | tee gpurun_out/next_quick.log
