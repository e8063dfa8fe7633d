#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/ab13_pytest.log
bb() {
  echo "== build_bench lib=$1 scene=$2"
  FOUNDATION_PT_LIB=$PWD/ab_libs/$1.so timeout 120 python scripts/build_bench.py $2 2>&1 | tail -2
}
{
bb prev terrain
bb lb_rs6 terrain
bb lbw_rs4 terrain
bb lbw_rs5 terrain
bb lbw_rs6 terrain
bb prev terrain
bb lb_rs6 terrain
bb lbw_rs4 terrain
bb lbw_rs5 terrain
bb lbw_rs6 terrain
bb lbw_rs6 spheres
bb lbw_rs6 instanced
} 2>&1 | tee gpurun_out/ab13.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_build_is_byte_identical_to_oracle" 2>&1 | tail -4 | tee gpurun_out/ab13_racecheck.log
