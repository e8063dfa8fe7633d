#!/bin/bash
mkdir -p gpurun_out
bb() {
  echo "== build_bench lib=$1 scene=$2"
  FOUNDATION_PT_LIB=$PWD/ab_libs/$1.so timeout 300 python scripts/build_bench.py $2 2>&1 | tail -2
}
{
bb prev terrain
bb rt256 terrain
bb rt128 terrain
bb rt64 terrain
bb prev terrain
bb rt256 terrain
bb rt128 terrain
bb rt64 terrain
} 2>&1 | tee gpurun_out/ab10.log
for T in 64 128; do
FOUNDATION_PT_LIB=$PWD/ab_libs/rt$T.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ab10_launches_$T.csv python scripts/probe.py --scene terrain --rays 1024 --reps 1 > /dev/null 2>&1
grep -E "k_refit" gpurun_out/ab10_launches_$T.csv | cut -d, -f5,12- | head -3
done
