#!/bin/bash
# Round-end GPU session: all GPU tests, smoke, the default bench + reference arm, the census,
# the ncu launch list and one full capture per math mode.
tag=${1:-fin}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log ); tail -3 gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err; echo "ref exit $?"
timeout 900 python bench.py --census 100000000 > gpurun_out/${tag}_census.json 2> gpurun_out/${tag}_census.err; echo "census exit $?"; tail -c 700 gpurun_out/${tag}_census.json
bash scripts/gpu_ncu.sh ${tag} fast > /dev/null 2>&1
ls gpurun_out | grep ${tag}_ | tr '\n' ' '
