#!/bin/bash
# Default bench (all legs), reference arm, census.  Outputs gpurun_out/<tag>_*.
tag=${1:-b}
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ) 2>&1 | grep real
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err ) 2>&1 | grep real
( time timeout 900 python bench.py --census ${2:-100000000} > gpurun_out/${tag}_census.json 2> gpurun_out/${tag}_census.err ) 2>&1 | grep real
tail -c 2500 gpurun_out/${tag}_bench.json; echo; cat gpurun_out/${tag}_ref.json | cut -c1-400; cat gpurun_out/${tag}_census.json; tail -3 gpurun_out/${tag}_bench.err gpurun_out/${tag}_census.err
