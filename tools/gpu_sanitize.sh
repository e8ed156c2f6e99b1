#!/bin/bash
# compute-sanitizer over the tiny parity configuration; logs under gpurun_out/<tag>_sanitize_*.log
# usage: bash tools/gpu_sanitize.sh <tag> [2gpu]
tag=$1
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > gpurun_out/${tag}_sanitize_${tool}.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize target' gpurun_out/${tag}_sanitize_${tool}.log | tr '\n' ' ')"
done
if [ "$2" == "2gpu" ]; then
  for tool in memcheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool --target-processes all --print-limit 20 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/sanitize_target.py > gpurun_out/${tag}_sanitize2_${tool}.log 2>&1
    echo "2gpu $tool rc=$? $(grep -E 'ERROR SUMMARY|sanitize target' gpurun_out/${tag}_sanitize2_${tool}.log | tr '\n' ' ')"
  done
fi
