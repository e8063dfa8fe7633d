#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab7_pytest.log
cat gpurun_out/ab7_pytest.log
bb() {
  echo "== build_bench lib=$1 scene=$2"
  FOUNDATION_PT_LIB=$PWD/ab_libs/$1.so timeout 300 python scripts/build_bench.py $2 2>&1 | tail -2
}
{
bb cur terrain
bb top terrain
bb cur terrain
bb top terrain
bb cur spheres
bb top spheres
bb cur instanced
bb top instanced
} 2>&1 | tee gpurun_out/ab7.log
FOUNDATION_PT_LIB=$PWD/ab_libs/top.so timeout 300 python scripts/tlas_bench.py 2>&1 | tail -4 | tee -a gpurun_out/ab7.log
