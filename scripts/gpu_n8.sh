#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep -E '^\{|Error|error' | tail -3 | tee gpurun_out/bench_n$N.json
free -g | head -2
