#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/last_pytest.log
timeout 100 python bench.py 2>/dev/null | tail -1 > gpurun_out/last_bench_n1.json
cut -c1-200 gpurun_out/last_bench_n1.json
