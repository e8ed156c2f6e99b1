#!/bin/bash
# 2+ GPU check of the peer-slab result exchange: sharded == single (both gather modes), then the sharded bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
G=${1:-2}
export PIPS_B200_PEER_TIMEOUT_MS=20000
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 tools/check_sharded.py 2>&1 | grep -v Warning | tail -n 30
echo "check rc=${PIPESTATUS[0]}"
for mode in ${2:-p2p nccl}; do
PIPS_B200_GATHER=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/bench_g${G}_$mode.json 2> gpurun_out/bench_g${G}_$mode.err; echo "bench $mode x$G rc=$?"
tail -n 3 gpurun_out/bench_g${G}_$mode.err; python -c "
import json; d=json.load(open('gpurun_out/bench_g${G}_$mode.json')); print('$mode n_gpus',d['n_gpus'],'value',round(d['value']),'ms/step',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value']))"
done
