#!/bin/bash
# quick end-to-end check of HEAD: GPU tests, smoke, both bench arms with default flags
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/check_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/check_smoke.log
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 2 ) 2>&1 | tail -5 > gpurun_out/check_bench_reference.log
( time timeout 1200 python bench.py ) 2>&1 | tail -5 > gpurun_out/check_bench.log
cat gpurun_out/check_bench_reference.log gpurun_out/check_bench.log
