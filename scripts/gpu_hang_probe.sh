#!/bin/bash
# Runs a command; if it is still alive after $1 seconds, dumps the host stacks of all its threads (cuda-gdb as gdb) and kills it.
#   bash scripts/gpu_hang_probe.sh 120 python scripts/gpu_multi_scaling.py c3 c2 c2cont
limit=$1; shift
mkdir -p gpurun_out
"$@" > gpurun_out/hang_probe_stdout.log 2>&1 &
pid=$!
for ((s = 0; s < limit; s++)); do
    if ! kill -0 $pid 2>/dev/null; then wait $pid; echo "finished rc=$? after ${s}s"; tail -5 gpurun_out/hang_probe_stdout.log; exit 0; fi
    sleep 1
done
echo "still running after ${limit}s: dumping stacks"
tail -5 gpurun_out/hang_probe_stdout.log
timeout 120 /usr/local/cuda/bin/cuda-gdb -p $pid -batch -ex "set pagination off" -ex "thread apply all bt 14" > gpurun_out/hang_probe_stacks.log 2>&1
grep -E "^Thread|^#" gpurun_out/hang_probe_stacks.log | grep -v "futex_wait\|pthread_cond\|in ?? ()" | head -150
kill -9 $pid
