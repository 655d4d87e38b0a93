#!/bin/bash
# Quick GPU pass: parity tests + bench (ours only).
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/gpu_tests.log
  cat gpurun_out/gpu_tests.log
fi
SDFB200_TIMING=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value %.2f Gq/s  e2e %.2f Gq/s  frac %.3f' % (d['value']/1e9, d['e2e']['value']/1e9, d['roofline']['frac']))
for k in ('octree_c2','octree_c2_continuity','exact_c3'):
    b=d['build'][k]; print(k, b['seconds'], b['all_seconds'], {a:round(v,1) for a,v in b['stats_ms_rank0'].items() if a.endswith('_ms')})
print('exact_query', d['exact_query'])
PY
tail -25 gpurun_out/bench_quick.err
