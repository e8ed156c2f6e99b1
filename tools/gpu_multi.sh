#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
G=${1:-2}
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 tools/check_sharded.py 2>&1 | grep -v Warning | tail -n 12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/bench_g$G.json 2> gpurun_out/bench_g$G.err; echo "bench x$G rc=$?"
tail -n 3 gpurun_out/bench_g$G.err; python -c "
import json; d=json.load(open('gpurun_out/bench_g$G.json')); print('n_gpus',d['n_gpus'],'value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus $G --steps 1 --warmup 1 | tail -n 1 | cut -c1-300
