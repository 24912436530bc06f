#!/bin/bash
# ncu captures of the production kernel: launch list + one full capture (f64; "fast" with a 2nd arg).
tag=${1:-r2a}
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e --no-other-math --stars 16777216 --steps 3 --warmup 3"
$B > gpurun_out/${tag}_plain_f64.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_ncu_launches.log 2>&1
$B > gpurun_out/${tag}_plain_f64b.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -f -o gpurun_out/${tag}_prof_f64 $B > gpurun_out/${tag}_ncu_f64.log 2>&1
if [ -n "$2" ]; then
$B --math fast > gpurun_out/${tag}_plain_fast.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -f -o gpurun_out/${tag}_prof_fast $B --math fast > gpurun_out/${tag}_ncu_fast.log 2>&1
fi
ls -la gpurun_out | tail -8
