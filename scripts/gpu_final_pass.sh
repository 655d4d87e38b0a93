#!/bin/bash
# Final pass of a round on ONE GPU: full parity suite, the bench (ours, then the reference arm), the launch list of the bench
# command, one ncu --set full capture of the headline kernel, and a compute-sanitizer memcheck of smoke().
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_final_gpu_tests.log
cat gpurun_out/r2_final_gpu_tests.log
timeout 600 python bench.py --steps ${STEPS:-50} --warmup 5 > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2_final_bench_1gpu.json') if l.startswith('{')][-1]
print('value %.2f Gq/s  e2e %.2f Gq/s (pageable %.2f)  frac %.3f  cpu %.3f Gq/s on %d cores' % (d['value']/1e9, d['e2e']['value']/1e9, d['e2e_pageable']['value']/1e9, d['roofline']['frac'],
      d.get('cpu_baseline',{}).get('value',0)/1e9, d.get('cpu_baseline',{}).get('cores',0)))
for k in ('octree_c2','octree_c2_continuity','exact_c3','exact_c4'):
    b=d['build'].get(k)
    if b: print(k, 'best %.4f first %.4f' % (b['seconds'], b['first_call_seconds']), [round(x,3) for x in b.get('all_seconds',[])], {a:round(v,1) for a,v in b['stats_ms_rank0'].items() if a.endswith('_ms')})
print('exact_query', d['exact_query'])
print('queries', {k: round(v['value']/1e9,2) for k,v in d['queries'].items()})
print('clocks', d['clocks'])
PY
tail -3 gpurun_out/r2_final_bench_1gpu.err
NCU="timeout 400 ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_final_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config4 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:octreeQueryTileKernel -s 2 -c 1 -o gpurun_out/r2_final_query_tile_grid -f python scripts/profile_kernels.py octree_query > gpurun_out/ncu_final_q.log 2>&1
tail -n 2 gpurun_out/ncu_final_q.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_sanitizer_memcheck_smoke.log 2>&1
echo "memcheck rc=$?"; tail -4 gpurun_out/r2_final_sanitizer_memcheck_smoke.log
if [ -z "$SKIP_REFERENCE" ]; then
  timeout 600 python bench.py --impl reference --steps ${REF_STEPS:-10} --warmup 3 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
  python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r2_final_bench_reference.json') if l.startswith('{')][-1]
print('reference arm: %.4f Gq/s on %d cores; builds' % (d['value']/1e9, d['cpu_baseline']['cores']), {k:(v or {}).get('seconds') for k,v in d.get('build',{}).items()})
PY
fi
ls -la gpurun_out | tail -12
