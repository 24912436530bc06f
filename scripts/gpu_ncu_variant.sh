#!/bin/bash
# one full ncu capture of k_fused per build/variants/*.so (development builds)
tag=${1:-nv}
mkdir -p gpurun_out
B="python bench.py --no-cpu --no-e2e --no-other-math --stars 16777216 --steps 3 --warmup 3"
for lib in build/variants/*.so; do
  [ -f "$lib" ] || continue
  name=$(basename $lib .so)
  STREAMSIM_B200_LIB=$PWD/$lib ncu --set full --clock-control none -k regex:k_fused -s 3 -c 1 -f -o gpurun_out/${tag}_${name} $B > gpurun_out/${tag}_${name}.log 2>&1
done
ls gpurun_out | grep ${tag}_ | tr '\n' ' '
