#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab5_pytest.log
cat gpurun_out/ab5_pytest.log
bb() {
  echo "== build_bench lib=$1 scene=$2"
  FOUNDATION_PT_LIB=$PWD/ab_libs/$1.so timeout 300 python scripts/build_bench.py $2 2>&1 | tail -2
}
{
bb refit2 terrain
bb refit4 terrain
bb refit2 terrain
bb refit4 terrain
bb refit2 spheres
bb refit4 spheres
bb refit4 instanced
} 2>&1 | tee gpurun_out/ab5.log
FOUNDATION_PT_LIB=$PWD/ab_libs/refit4.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ab5_launches_build.csv python scripts/probe.py --scene terrain --rays 1024 --reps 1 > gpurun_out/ab5_ncu.log 2>&1
grep -E "k_refit|k_karras|k_write_tris" gpurun_out/ab5_launches_build.csv | cut -d, -f5,12- | head
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_build_is_byte_identical_to_oracle" 2>&1 | tail -5 | tee gpurun_out/ab5_racecheck.log
