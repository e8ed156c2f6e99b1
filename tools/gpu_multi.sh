#!/bin/bash
# Multi-GPU checks on one box: bit-exactness of the sharded path, phase table, bench line.  usage: bash tools/gpu_multi.sh <tag> <G> [bench]
tag=$1; G=$2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29577"
timeout 600 $TR tools/check_sharded.py > gpurun_out/${tag}_check_sharded_${G}.log 2>&1; echo "check_sharded rc=$?"
grep -E "bit-exact" gpurun_out/${tag}_check_sharded_${G}.log | grep "rank 0" | head -8
timeout 600 $TR tools/diag_scale.py > gpurun_out/${tag}_diag_${G}.log 2>&1; echo "diag rc=$?"
grep -E "^---|^\{|^===" gpurun_out/${tag}_diag_${G}.log | head -44
if [ "$3" == "bench" ]; then
  timeout 900 $TR bench.py --gpus $G --steps 20 --warmup 5 > gpurun_out/${tag}_bench_${G}gpu.json 2> gpurun_out/${tag}_bench_${G}gpu.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench_${G}gpu.json"))
print("G=$G ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "value", round(d["value"]/1e6,2), "M upd/s; clocks", d["clocks"].get("per_gpu_sm_mhz"))
for k in ("strong_cfg3","cfg4_sharded"):
    b=d.get(k,{}); print(k, b.get("ms_per_step"), b.get("e2e",{}).get("ms_per_step"), b.get("error"))
PY
  tail -3 gpurun_out/${tag}_bench_${G}gpu.err
fi
