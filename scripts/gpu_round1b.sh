#!/bin/bash
# parity tests + smoke + bench (both arms) + ncu launch list + ncu full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_reference.json
timeout 1200 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_n1.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_rays -s 2 -c 2 -o gpurun_out/prof_trace -f python scripts/probe.py --rays 16777216 --reps 4 > gpurun_out/ncu_trace.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_shade|k_connect|k_key_scatter" -s 8 -c 8 -o gpurun_out/prof_wave -f python scripts/probe.py --scene spheres --rays 1048576 --reps 1 --spp 1 > gpurun_out/ncu_wave.log 2>&1
ls -la gpurun_out
