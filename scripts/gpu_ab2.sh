#!/bin/bash
# one-box experiment: GPU parity tests on the tree's library, then same-box A/B of the variant libraries under ab_libs/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab2_pytest.log
cat gpurun_out/ab2_pytest.log
run() {  # lib scene
  echo "== lib=$1 scene=$2"
  FOUNDATION_PT_LIB=$PWD/ab_libs/$1.so timeout 300 python scripts/probe.py --scene $2 --rays 16777216 --reps 3 --spp 16 2>&1 | grep -E "commit|closest|any:|render|Error|error"
}
bb() {
  echo "== build_bench lib=$1 scene=$2"
  FOUNDATION_PT_LIB=$PWD/ab_libs/$1.so timeout 300 python scripts/build_bench.py $2 2>&1 | tail -3
}
{
bb base terrain
bb refit terrain
bb base terrain
bb refit terrain
bb base instanced
bb refit instanced
run base terrain
run wide terrain
run base terrain
run wide terrain
run base spheres
run wide spheres
run base instanced
run wide instanced
} 2>&1 | tee gpurun_out/ab2.log
