#!/bin/bash
mkdir -p gpurun_out
# launch list of a pure render (wavefront breakdown)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_render.csv python scripts/probe.py --scene terrain --rays 1024 --reps 1 --spp 8 > gpurun_out/render_under_ncu.log 2>&1
# the bench's own trace kernel (2^26 rays per launch): full set, one launch
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_rays -s 3 -c 1 -o gpurun_out/prof_trace_bench -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --spp 0 --e2e-steps 1 > gpurun_out/ncu_trace_bench.log 2>&1
timeout 1200 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n1_b.json
