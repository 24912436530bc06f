#!/bin/bash
# One GPU session: tests, bench, geometry variants.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log )
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
for lib in build/variants/*.so; do
  name=$(basename $lib .so)
  STREAMSIM_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu --no-e2e --steps 10 > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  echo "$name exit $?"
done
timeout 300 python bench.py --no-cpu --no-e2e --steps 10 --fused 0 > gpurun_out/${tag}_generic.json 2> gpurun_out/${tag}_generic.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/*_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.4g" % d["value"], "frac %.4f" % d["roofline"]["frac"], "other", d.get("other_math", {}).get("value"), "e2e %.4g" % d["e2e"]["value"])
    except Exception as e:
        print(f, "unreadable", e)
PY
