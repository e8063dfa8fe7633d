#!/bin/bash
# end-of-round-2 evidence on one B200: GPU tests, smoke, both bench arms, ncu launch lists (bench / render / build), one `--set full` capture of the traversal kernel
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r02_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tail -1 > gpurun_out/r02_bench_reference.json
timeout 1200 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r02_bench_n1.json
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra-sets --e2e-steps 1 --spp 8 > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_rays -s 3 -c 1 -o gpurun_out/r02_prof_trace -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-sets --spp 0 --e2e-steps 1 > gpurun_out/r02_ncu_trace.log 2>&1
ncu -i gpurun_out/r02_prof_trace.ncu-rep --page raw --csv > gpurun_out/r02_trace_kernel_bench_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_prof_trace.ncu-rep --page source --csv > gpurun_out/r02_trace_kernel_source.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_render.csv python scripts/probe.py --scene terrain --rays 1024 --reps 1 --spp 8 > gpurun_out/r02_render_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_build.csv python scripts/build_bench.py terrain > gpurun_out/r02_build_under_ncu.log 2>&1
for s in spheres instanced cornell; do timeout 300 python scripts/probe.py --scene $s --rays 16777216 --reps 3 --spp 16 2>&1 | grep -E "closest|any:|render|hit fraction|commit" | tail -5; done | tee gpurun_out/r02_other_configs.log
cut -c1-400 gpurun_out/r02_bench_n1.json
