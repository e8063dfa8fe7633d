#!/bin/bash
# 2 GPUs: the multi-GPU tests (group / comm_init / C++ editor), then the bench under torchrun (both gather modes) and the reference arm
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nvidia-smi topo -m 2>/dev/null | head -12
NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r2_multi_pytest_n$N.log
NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.log 2>&1
tail -3 gpurun_out/r2_bench_n$N.log | cut -c1-2500
