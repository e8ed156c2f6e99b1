#!/bin/bash
# Run the GPU kernel tests one group per process (a device-side trap poisons the CUDA context of
# its process only) and keep the logs under gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for k in test_pyramid_build test_init_gather test_corr_gather test_gemm_f32 test_tokenmix test_update "test_gemm_tc and 3-128-256-64" "test_gemm_tc and 1-128-256-64" test_gemm_tc; do
  name=$(echo "$k" | tr ' ' '_')
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -s -p no:cacheprovider -k "$k" > "gpurun_out/k_${name}.log" 2>&1
  echo "== $k -> rc=$? : $(tail -n 1 gpurun_out/k_${name}.log)"
done
timeout 900 python -m pytest tests/test_parity_gpu.py -q -s -p no:cacheprovider -k "fp32 or cpu_inputs" > gpurun_out/parity_fp32.log 2>&1
echo "== parity fp32 rc=$? : $(tail -n 1 gpurun_out/parity_fp32.log)"
timeout 900 python -m pytest tests/test_parity_gpu.py -q -s -p no:cacheprovider -k "not fp32" > gpurun_out/parity_tc.log 2>&1
echo "== parity tc rc=$? : $(tail -n 1 gpurun_out/parity_tc.log)"
grep -h -E "FAILED|Error|error|max err|per-iter|corr err" gpurun_out/*.log | head -80
