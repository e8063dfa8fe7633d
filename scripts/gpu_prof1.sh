#!/bin/bash
# ncu captures (--set full) of the traversal, wavefront and build kernels; raw / source pages exported as CSV
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
timeout 300 $N -k regex:k_trace_rays -s 1 -c 1 -o gpurun_out/p_trace -f python scripts/probe.py --scene terrain --rays 67108864 --reps 2 > gpurun_out/p_trace.log 2>&1
timeout 300 $N -k "regex:k_extend|k_connect|k_shade" -c 6 -o gpurun_out/p_render -f python scripts/probe.py --scene terrain --rays 1024 --reps 1 --spp 8 > gpurun_out/p_render.log 2>&1
timeout 300 $N -k "regex:k_refit|k_collapse|k_rs_|k_karras|k_write_tris|k_tri_boxes|k_morton" -c 60 -o gpurun_out/p_build -f python scripts/probe.py --scene terrain --rays 1024 --reps 1 > gpurun_out/p_build.log 2>&1
for f in p_trace p_render p_build; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/p_trace.ncu-rep --page source --csv > gpurun_out/p_trace_source.csv 2>/dev/null
ncu -i gpurun_out/p_render.ncu-rep --page source --csv --kernel-name regex:k_extend > gpurun_out/p_extend_source.csv 2>/dev/null
ls -la gpurun_out
rm -f gpurun_out/p_build.ncu-rep gpurun_out/p_render.ncu-rep
