#!/bin/bash
# ncu --set full captures of single kernels on the micro-benchmarks (one GPU).  usage: bash tools/gpu_ncu.sh <tag> <what...>
tag=$1; shift
mkdir -p gpurun_out
for what in "$@"; do
  case $what in
    tokenmix) timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokenmix -s 3 -c 1 -f -o gpurun_out/${tag}_tokenmix python tools/bench_tokenmix.py 4096 > gpurun_out/${tag}_ncu_tokenmix.log 2>&1 ;;
    tokenmix_simt) PIPS_B200_TOKENMIX=simt timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokenmix -s 3 -c 1 -f -o gpurun_out/${tag}_tokenmix_simt python tools/bench_tokenmix.py 4096 > gpurun_out/${tag}_ncu_tokenmix_simt.log 2>&1 ;;
    corr) timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_gather -s 3 -c 1 -f -o gpurun_out/${tag}_corr python tools/bench_corr.py > gpurun_out/${tag}_ncu_corr.log 2>&1 ;;
    corr_big) timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_gather -s 3 -c 1 -f -o gpurun_out/${tag}_corr_big python tools/bench_corr.py --big > gpurun_out/${tag}_ncu_corr_big.log 2>&1 ;;
  esac
  echo "$what rc=$?"
done
timeout 120 python tools/bench_corr.py; timeout 120 python tools/bench_corr.py --big; timeout 120 python tools/bench_corr.py --bf16
ls -la gpurun_out/${tag}_*.ncu-rep
