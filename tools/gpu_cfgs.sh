#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python tools/bench_configs.py 2>&1 | tail -n 8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_iter.json')); print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'launches', d['gpu_launches'])"
tail -n 3 gpurun_out/bench_iter.err
