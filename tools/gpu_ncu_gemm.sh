#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 3 -c 2 -o gpurun_out/prof_gemm2 -f python tools/profile_step.py bf16x3 fp32 > gpurun_out/ncu_gemm2.log 2>&1; echo "ncu gemm2 rc=$?"
tail -n 3 gpurun_out/ncu_gemm2.log
