#!/bin/bash
# bench (device leg, f64 only) for every build/variants/*.so
tag=${1:-wi}
mkdir -p gpurun_out
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
        print("%-50s %.4g stars/s  %.3f ms" % (f, d["value"], d["ms_per_step"]))
    except Exception as e:
        print(f, "unreadable", e, open(f.replace(".json", ".err")).read()[-300:])
PY
