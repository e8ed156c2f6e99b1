#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for spec in "gemm_tc2_kernel prof_gemm_final 3 2" "conv_tc_kernel prof_conv_final 1 2" "corr_gather prof_corr_final 0 1" "tokenmix prof_tok_final 2 1"; do
  set -- $spec
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$1 -s $3 -c $4 -o gpurun_out/$2 -f python tools/profile_step.py bf16x3 fp32 > gpurun_out/ncu_$2.log 2>&1; echo "$1 rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail -n 6
