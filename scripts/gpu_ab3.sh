#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab3_pytest.log
cat gpurun_out/ab3_pytest.log
bb() {
  echo "== build_bench lib=$1 scene=$2"
  FOUNDATION_PT_LIB=$PWD/ab_libs/$1.so timeout 300 python scripts/build_bench.py $2 2>&1 | tail -3
}
{
bb refit terrain
bb refit2 terrain
bb refit terrain
bb refit2 terrain
bb refit spheres
bb refit2 spheres
bb refit2 instanced
} 2>&1 | tee gpurun_out/ab3.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_build_is_byte_identical_to_oracle" 2>&1 | tail -12 | tee gpurun_out/ab3_racecheck.log
