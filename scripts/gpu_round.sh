#!/bin/bash
# One GPU session: full GPU tests, then the bench (device leg only) for the in-tree library and variants.
tag=${1:-r2a}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log )
tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --no-cpu --no-e2e --steps 10 > gpurun_out/${tag}_intree.json 2> gpurun_out/${tag}_intree.err
for lib in build/variants/*.so; do
  [ -f "$lib" ] || continue
  name=$(basename $lib .so)
  STREAMSIM_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-cpu --no-e2e --no-other-math --steps 10 > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
done
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
