#!/bin/bash
# A/B of an environment switch on the default bench line (short): usage bash tools/gpu_ab.sh <tag> VAR=a VAR=b
tag=$1; shift
for kv in "$@"; do
  env $kv timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --no-eager --no-cpu-baseline > gpurun_out/${tag}_${kv//[^A-Za-z0-9_]/_}.json 2>gpurun_out/${tag}_ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_${kv//[^A-Za-z0-9_]/_}.json"))
print("$kv", "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "clk", d["clocks"]["sm_mhz"], d["ms_per_step_stats_rank0"]["device_loop"])
PY
done
