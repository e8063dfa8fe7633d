#!/bin/bash
# after the sort / e2e changes: GPU test tier, one ncu capture of the one-sweep pass, bench line
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/r2q_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_rs_onesweep' -s 18 -c 1 -o gpurun_out/r2q_onesweep -f python scripts/build_bench.py terrain > gpurun_out/r2q_onesweep.log 2>&1
ncu -i gpurun_out/r2q_onesweep.ncu-rep --page raw --csv > gpurun_out/r2q_onesweep_raw.csv 2>/dev/null
ncu -i gpurun_out/r2q_onesweep.ncu-rep --page source --csv > gpurun_out/r2q_onesweep_source.csv 2>/dev/null
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r2q_bench.json
cut -c1-700 gpurun_out/r2q_bench.json
