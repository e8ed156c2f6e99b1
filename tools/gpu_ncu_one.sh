#!/bin/bash
# usage: gpu_ncu_one.sh <kernel regex> <out name> [skip] [count]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$1 -s ${3:-2} -c ${4:-1} -o gpurun_out/$2 -f python tools/profile_step.py bf16x3 fp32 > gpurun_out/ncu_$2.log 2>&1; echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_$2.log
