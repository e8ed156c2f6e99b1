#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -4
for t in 0 1; do
PIPS_B200_GEMM_TAIL=$t timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tail$t.json 2> gpurun_out/bench_tail$t.err; echo "bench tail=$t rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_tail$t.json')); print('tail=$t ms/step', round(d['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], 'fc1', round(d['kernel_ms_per_iteration']['gemm_fc1'],3), 'fc2', round(d['kernel_ms_per_iteration']['gemm_fc2'],3))"
done
