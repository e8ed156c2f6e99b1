#!/bin/bash
# One gpurun call: the -m gpu suite, micro-benchmarks, then the default bench line.  Outputs under gpurun_out/<tag>_*.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_check.sh r02a [pytest -k expr]'
tag=${1:-check}; kexpr=${2:-}
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q -s -k "$kexpr" > gpurun_out/${tag}_pytest.log 2>&1
else
  timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/${tag}_pytest.log 2>&1
fi
echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
grep -E "passed|failed|Error|error" gpurun_out/${tag}_pytest.log | tail -8
PIPS_B200_TOKENMIX=simt timeout 120 python tools/bench_tokenmix.py 4096 > gpurun_out/${tag}_tokenmix.txt 2>&1
timeout 120 python tools/bench_tokenmix.py 4096 >> gpurun_out/${tag}_tokenmix.txt 2>&1
cat gpurun_out/${tag}_tokenmix.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "clk", d["clocks"]["sm_mhz"])
print(d["kernel_ms_per_iteration"], d["loop_only"], "fnet", d.get("fnet_ms"))
for k in ("n4096_1gpu","cfg1_demo_shape_1gpu","cfg4_1gpu"):
    print(k, d.get(k,{}).get("ms_per_step"), d.get(k,{}).get("error"))
print(d.get("cfg5_chain_1gpu"))
PY
tail -5 gpurun_out/${tag}_bench.err
