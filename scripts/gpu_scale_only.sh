#!/bin/bash
# bench at each N given (no tests): gpu_scale_only.sh <tag> N [N ...]
tag=${1:-m}; shift
mkdir -p gpurun_out
for n in "$@"; do
  port=$((29500 + n))
  if [ "$n" = "1" ]; then
    timeout 900 python bench.py --gpus 1 > gpurun_out/${tag}_scale_$n.json 2> gpurun_out/${tag}_scale_$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n > gpurun_out/${tag}_scale_$n.json 2> gpurun_out/${tag}_scale_$n.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_scale_$n.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("N=%d value %.4g frac %.4f ms %.3f obs_frac %.5f | e2e %.4g (%.1f GB/s, alone %.1f, ceiling %.1f, frac %.3f) | cpu %s" % (
        d["n_gpus"], d["value"], d["roofline"]["frac"], d["ms_per_step"], d["config"]["observed_fraction"],
        e["value"], e["gbs"], e["pcie_gbs"], e["copy_ceiling_gbs"], e["frac_of_ceiling"], d.get("cpu_baseline", {}).get("value")))
except Exception as ex:
    print("unreadable", ex)
PY
done
