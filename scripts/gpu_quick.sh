#!/bin/bash
# Quick GPU pass: parity tests + bench (ours, then the reference arm unless SKIP_REFERENCE is set).
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/gpu_tests.log
  cat gpurun_out/gpu_tests.log
fi
timeout 900 python bench.py --steps ${STEPS:-50} --warmup 5 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value %.2f Gq/s  e2e %.2f Gq/s (pageable %.2f)  frac %.3f  cpu %.3f Gq/s on %d cores' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e_pageable']['value']/1e9, d['roofline']['frac'],
      d.get('cpu_baseline',{}).get('value',0)/1e9, d.get('cpu_baseline',{}).get('cores',0)))
for k in ('octree_c2','octree_c2_continuity','exact_c3','exact_c4'):
    b=d['build'].get(k)
    if b: print(k, 'best %.4f first %.4f' % (b['seconds'], b['first_call_seconds']), {a:round(v,1) for a,v in b['stats_ms_rank0'].items() if a.endswith('_ms')})
print('exact_query', d['exact_query'])
print('queries', {k: round(v['value']/1e9,2) for k,v in d['queries'].items()})
print('clocks', d['clocks'])
PY
tail -5 gpurun_out/bench_quick.err
if [ -z "$SKIP_REFERENCE" ]; then
  timeout 900 python bench.py --impl reference --steps ${REF_STEPS:-10} --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
  python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_reference.json'))
print('reference arm: %.4f Gq/s on %d cores; builds' % (d['value']/1e9, d['cpu_baseline']['cores']), {k:(v or {}).get('seconds') for k,v in d['build'].items()})
PY
fi
