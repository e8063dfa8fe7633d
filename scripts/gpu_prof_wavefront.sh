#!/bin/bash
# ncu --set full of the three wavefront kernels inside a 16-spp config-3 render (one wave): k_extend at bounce 1 (surface-started rays), k_connect and k_shade at bounce 0
mkdir -p gpurun_out
prof() {  # name regex skip
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o gpurun_out/r02_wf_$1 -f python scripts/probe.py --scene terrain --rays 1024 --reps 1 --spp 16 > gpurun_out/r02_wf_$1.log 2>&1
  ncu -i gpurun_out/r02_wf_$1.ncu-rep --page raw --csv > gpurun_out/r02_wf_$1_raw.csv 2>/dev/null
}
prof extend '^k_extend' 10
prof connect '^k_connect' 9
prof shade '^k_shade' 9
ls -la gpurun_out/r02_wf_*_raw.csv
