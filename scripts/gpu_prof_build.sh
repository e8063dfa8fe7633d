#!/bin/bash
# ncu --set full captures of the build's heaviest kernels in steady state (third build of scripts/build_bench.py)
mkdir -p gpurun_out
prof() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/r2m_$1 -f python scripts/build_bench.py terrain > gpurun_out/r2m_$1.log 2>&1
  ncu -i gpurun_out/r2m_$1.ncu-rep --page raw --csv > gpurun_out/r2m_$1_raw.csv 2>/dev/null
}
prof scatter '^k_rs_scatter' 18 1
prof hist '^k_rs_hist' 18 1
prof refit '^k_refit_agg$' 2 1
prof collapse '^k_collapse_levels' 2 1
ls -la gpurun_out/r2m_*
