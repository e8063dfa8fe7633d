#!/bin/bash
# N GPUs of one box: multi-GPU tests (group / comm_init / C++ editor --gpus N), the bench line (weak), the strong-scaling arm of the ray-set metric
mkdir -p gpurun_out
N=${1:-8}
NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02_multi_pytest_n$N.log
NCCL_DEBUG=WARN timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.log 2>&1
grep '^{"metric"' gpurun_out/r02_bench_n$N.log > gpurun_out/r02_bench_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29578 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --no-cpu-baseline --no-extra-sets --spp 0 > gpurun_out/r02_bench_strong_n$N.log 2>&1
grep '^{"metric"' gpurun_out/r02_bench_strong_n$N.log > gpurun_out/r02_bench_strong_n$N.json
cut -c1-300 gpurun_out/r02_bench_n$N.json; cut -c1-300 gpurun_out/r02_bench_strong_n$N.json
