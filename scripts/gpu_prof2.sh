#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_refit -c 1 -o gpurun_out/p_refit -f python scripts/probe.py --scene terrain --rays 1024 --reps 1 > gpurun_out/p_refit.log 2>&1
ncu -i gpurun_out/p_refit.ncu-rep --page raw --csv > gpurun_out/p_refit_raw.csv 2>/dev/null
ncu -i gpurun_out/p_refit.ncu-rep --page source --csv > gpurun_out/p_refit_source.csv 2>/dev/null
rm -f gpurun_out/p_refit.ncu-rep
