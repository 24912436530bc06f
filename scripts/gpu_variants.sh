#!/bin/bash
# Quick perf comparison: in-tree library + every build/variants/*.so, both math modes.
tag=${1:-v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "fused or fast_math or lsst_configuration or summary_only or shard or empty_and_tiny or nside_4096" > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --no-cpu --no-e2e --steps 10 > gpurun_out/${tag}_intree.json 2> gpurun_out/${tag}_intree.err
for lib in build/variants/*.so; do
  [ -f "$lib" ] || continue
  name=$(basename $lib .so)
  STREAMSIM_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu --no-e2e --no-other-math --steps 10 > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
done
for extra in "$@"; do :; done
if [ -n "$SS_EXTRA_ENVS" ]; then
  for e in $SS_EXTRA_ENVS; do
    env $e timeout 300 python bench.py --no-cpu --no-e2e --no-other-math --steps 10 > gpurun_out/${tag}_env_${e//=/_}.json 2> gpurun_out/${tag}_env_${e//=/_}.err
  done
fi
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
