#!/bin/bash
# round-end sanity on one GPU: full -m gpu suite, smoke(), bench (with the eager-torch context leg), launch list, other configs
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
grep -E "per-iter|max err|max\|err|chain|bf16 features|teacher|score maps|sparse vs" gpurun_out/pytest_gpu.log | head -70 > gpurun_out/parity_log.txt; wc -l gpurun_out/parity_log.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 900 python bench.py --steps 10 --warmup 3 --with-eager > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -n 1 | cut -c1-300
bash tools/gpu_launchlist.sh 2>&1 | tail -n 24
timeout 600 python tools/bench_configs.py 2>&1 | tail -n 4
