#!/bin/bash
# round 2: GPU tests (incl. the full-size configs), smoke, the bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tail -2
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r2_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r2_smoke.log
( time timeout 1200 python bench.py ) > gpurun_out/r2_bench.log 2>&1
tail -4 gpurun_out/r2_bench.log | cut -c1-3000
