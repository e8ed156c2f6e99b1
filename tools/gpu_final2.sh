#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cur.csv python tools/profile_step.py bf16x3 fp32 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --with-eager > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cat gpurun_out/bench_final.json | cut -c1-400
timeout 600 python tools/bench_chain.py 2>&1 | tail -n 2
