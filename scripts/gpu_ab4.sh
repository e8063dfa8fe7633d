#!/bin/bash
mkdir -p gpurun_out
bb() {
  echo "== build_bench lib=$1 scene=$2"
  FOUNDATION_PT_LIB=$PWD/ab_libs/$1.so timeout 300 python scripts/build_bench.py $2 2>&1 | tail -2
}
{
bb refit2 terrain
bb refit3 terrain
bb refit2 terrain
bb refit3 terrain
} 2>&1 | tee gpurun_out/ab4.log
FOUNDATION_PT_LIB=$PWD/ab_libs/refit3.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ab4_launches_build.csv python scripts/probe.py --scene terrain --rays 1024 --reps 1 > gpurun_out/ab4_ncu.log 2>&1
grep -E "k_refit|k_karras|k_write_tris" gpurun_out/ab4_launches_build.csv | cut -d, -f5,12- | head
