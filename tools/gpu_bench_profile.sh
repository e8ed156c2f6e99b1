#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -n 2 gpurun_out/smoke.log)"
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "update" 2>&1 | tail -n 2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16x3.json 2> gpurun_out/bench_bf16x3.err; echo "bench rc=$?"; cat gpurun_out/bench_bf16x3.json; tail -n 5 gpurun_out/bench_bf16x3.err
timeout 600 python bench.py --steps 10 --warmup 3 --precision bf16 --feat bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 rc=$?"; cat gpurun_out/bench_bf16.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bf16x3.csv python tools/profile_step.py bf16x3 fp32 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 2 -o gpurun_out/prof_gemm -f python tools/profile_step.py bf16x3 fp32 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:corr_gather -c 1 -o gpurun_out/prof_corr -f python tools/profile_step.py bf16x3 fp32 > gpurun_out/ncu_corr.log 2>&1; echo "ncu corr rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tokenmix -c 1 -o gpurun_out/prof_tokenmix -f python tools/profile_step.py bf16x3 fp32 > gpurun_out/ncu_tok.log 2>&1; echo "ncu tokenmix rc=$?"
ls -la gpurun_out | tail -n 20
