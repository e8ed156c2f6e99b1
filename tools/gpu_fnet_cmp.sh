#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -p no:cacheprovider -s -k "fnet_modes" 2>&1 | grep -E "fnet|passed|failed|Error" | tail -n 20
for m in fast tc; do
  PIPS_B200_FNET=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$m.json 2> gpurun_out/bench_$m.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$m.json')); print('$m', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'loop ms/iter', round(d['loop_only']['ms_per_iteration'],3), 'clk', d['clocks']['sm_mhz'])"
done
