#!/bin/bash
# quick iteration: all GPU tests (one process), then the two bench lines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -n 15
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_iter.json'))
    print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'loop ms/iter', round(d['loop_only']['ms_per_iteration'],3))
    print('kernels', d['kernel_ms_per_iteration'])
    print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'corr', d['roofline_corr']['achieved'], 'clocks', d['clocks'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_iter.err').read()[-3000:])
PY
