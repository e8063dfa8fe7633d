#!/bin/bash
# end-of-round evidence: tests, smoke, both bench arms, launch list, full capture of the traversal kernel at bench size
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tail -1 > gpurun_out/bench_reference.json
timeout 1200 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_n1.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --spp 8 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_rays -s 3 -c 1 -o gpurun_out/prof_trace_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --spp 0 --e2e-steps 1 > gpurun_out/ncu_trace_final.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_render.csv python scripts/probe.py --scene terrain --rays 1024 --reps 1 --spp 8 > gpurun_out/render_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_build.csv python scripts/build_bench.py terrain > gpurun_out/build_under_ncu.log 2>&1
cat gpurun_out/bench_n1.json
