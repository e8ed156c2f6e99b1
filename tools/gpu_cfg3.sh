#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --particles 4096 > gpurun_out/bench_cfg3_1gpu.json 2> gpurun_out/bench_cfg3.err; echo "cfg3 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg3_1gpu.json')); print('cfg3 1GPU value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'loop ms/iter', round(d['loop_only']['ms_per_iteration'],3), d['kernel_ms_per_iteration'], 'roofline', round(d['roofline']['achieved'],1), 'corr', round(d['roofline_corr']['achieved']), 'clk', d['clocks'])"
tail -n 3 gpurun_out/bench_cfg3.err
timeout 600 python tools/bench_chain.py 2>&1 | tail -n 3
