#!/bin/bash
# first GPU contact: parity tests, then perf probes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python scripts/probe.py --scene terrain --n 500 --rays 4194304 --spp 2 2>&1 | tee gpurun_out/probe_small.log
timeout 900 python scripts/probe.py --scene terrain --n 2236 --rays 16777216 --spp 2 2>&1 | tee gpurun_out/probe_terrain.log
