#!/bin/bash
tag=${1:-s3}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log ); tail -3 gpurun_out/${tag}_pytest.log
# the paired-store variant: parity subset, then speed
STREAMSIM_B200_LIB=$PWD/build/variants/lib_vecstores.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "fused_free_running or shard_invariance or empty_and_tiny" > gpurun_out/${tag}_pytest_vec.log 2>&1; tail -2 gpurun_out/${tag}_pytest_vec.log
B="python bench.py --no-cpu --no-e2e --steps 10"
$B > gpurun_out/${tag}_intree.json 2> gpurun_out/${tag}_intree.err
STREAMSIM_B200_LIB=$PWD/build/variants/lib_vecstores.so $B > gpurun_out/${tag}_vecstores.json 2> gpurun_out/${tag}_vecstores.err
$B --workload wholesky > gpurun_out/${tag}_wholesky_packed.json 2> gpurun_out/${tag}_wholesky_packed.err
$B --workload wholesky --packed-maps 0 > gpurun_out/${tag}_wholesky_separate.json 2> gpurun_out/${tag}_wholesky_separate.err
$B --packed-maps 0 > gpurun_out/${tag}_pal5_separate.json 2> gpurun_out/${tag}_pal5_separate.err
timeout 600 python scripts/bench_paths.py inject 1e7 > gpurun_out/${tag}_inject_paths.json 2> gpurun_out/${tag}_inject_paths.err
python - <<PY
import glob, json
for f in sorted(glob.glob("gpurun_out/${tag}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        o = d.get("other_math", {})
        print("%-50s %s %.4g (frac %.4f)   %s %.4g" % (f, d["config"]["math"], d["value"], d["roofline"]["frac"], o.get("math"), o.get("value", 0)))
    except Exception as e:
        print(f, "unreadable", e)
PY
cat gpurun_out/${tag}_inject_paths.json
