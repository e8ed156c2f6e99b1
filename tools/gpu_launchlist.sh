#!/bin/bash
# ncu launch list (device time of every launch) of ONE forward at the bench configuration.  usage: bash tools/gpu_launchlist.sh <tag>
tag=$1; mkdir -p gpurun_out
PIPS_B200_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
   --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py > gpurun_out/${tag}_launches.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${tag}_launches.log
python tools/summarize_launches.py gpurun_out/${tag}_launches.csv "ncu --metrics gpu__time_duration.sum --clock-control none, one Pips.forward (BASELINE cfg2: B=4,S=8,384x512,N=1024,iters=6; bf16x3, fp32 pyramid, fnet tc; eager launches)" | tee gpurun_out/${tag}_launches_summary.txt | head -30
