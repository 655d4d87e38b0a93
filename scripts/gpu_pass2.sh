#!/bin/bash
# tests, quick bench, reference arm (incl. CONTINUITY build), launch list of the CONTINUITY build
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/gpu_tests.log
cat gpurun_out/gpu_tests.log
SDFB200_TIMING=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print('value %.2f Gq/s  e2e %.2f Gq/s  frac %.3f' % (d['value']/1e9, d['e2e']['value']/1e9, d['roofline']['frac']))
for k in ('octree_c2','octree_c2_continuity','exact_c3'):
    b=d['build'][k]; print(k, b['seconds'], b['all_seconds'], {a:round(v,1) for a,v in b['stats_ms_rank0'].items() if a.endswith('_ms')}, b.get('octree_words'))
print('exact_query', d['exact_query'])
PY
tail -5 gpurun_out/bench_quick.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python -c "
import json; d=json.load(open('gpurun_out/bench_ref.json')); print('reference', d['value']/1e9, 'Gq/s cores', d['cpu_baseline']['cores'], d['build'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_octree_cont.csv python scripts/profile_kernels.py octree_cont > /dev/null 2>&1
tail -3 gpurun_out/launches_octree_cont.csv
