#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/ab12_pytest.log
bb() {
  echo "== build_bench lib=$1 scene=$2"
  FOUNDATION_PT_LIB=$PWD/ab_libs/$1.so timeout 120 python scripts/build_bench.py $2 2>&1 | tail -2
}
{
bb prev terrain
bb lb terrain
bb lb_rs6 terrain
bb lb_rs8 terrain
bb prev terrain
bb lb terrain
bb lb_rs6 terrain
bb lb_rs8 terrain
bb prev spheres
bb lb spheres
bb lb_rs8 spheres
bb prev instanced
bb lb instanced
} 2>&1 | tee gpurun_out/ab12.log
