#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "test_gemm_tc and True-3-128-256-64" 2>&1 | tail -n 5
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "test_gemm_tc" 2>&1 | tail -n 12
