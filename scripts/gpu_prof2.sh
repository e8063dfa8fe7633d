#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for ws in 1 4 8; do echo "== wave samples $ws"; FOUNDATION_PT_WAVE_SAMPLES=$ws timeout 300 python scripts/probe.py --scene terrain --n 2236 --rays 16777216 --reps 2 --spp 16 2>&1 | grep -E "closest|render|commit" | tail -3; done
timeout 300 python scripts/probe.py --scene spheres --rays 16777216 --reps 2 --spp 16 2>&1 | grep -E "closest|any|render|commit" | tail -4
timeout 300 python scripts/probe.py --scene instanced --rays 16777216 --reps 2 --spp 16 2>&1 | grep -E "closest|any|render|commit" | tail -4
timeout 300 python scripts/probe.py --scene cornell --rays 16777216 --reps 2 --spp 16 2>&1 | grep -E "closest|any|render|commit" | tail -4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_rays -s 2 -c 1 -o gpurun_out/prof_trace2 -f python scripts/probe.py --rays 16777216 --reps 4 > gpurun_out/ncu_trace2.log 2>&1
