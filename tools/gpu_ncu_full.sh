#!/bin/bash
# ncu --set full of the dominant kernels inside one forward at the bench configuration.  usage: bash tools/gpu_ncu_full.sh <tag>
tag=$1; mkdir -p gpurun_out
cap() {  # name regex skip count
  PIPS_B200_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c $4 -f \
     -o gpurun_out/${tag}_full_$1 python tools/profile_step.py > gpurun_out/${tag}_full_$1.log 2>&1; echo "$1 rc=$?"
}
cap gemm gemm_tc2 3 2          # skip the first Linear + FC1/FC2 of layer 0's neighbours: launches 3,4 = FC1, FC2 of layer 1
cap corr corr_gather 1 1
cap tokenmix tokenmix_tc 2 1
cap convrows conv_rows 1 1
cap convtc conv_tc_kernel 13 1  # the 416 -> 256 head convolution is launch 17; 13 = a layer-3 3x3
ls -la gpurun_out/${tag}_full_*.ncu-rep
