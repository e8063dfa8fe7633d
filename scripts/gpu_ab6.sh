#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab6_pytest.log
cat gpurun_out/ab6_pytest.log
run() {  # lib scene
  echo "== lib=$1 scene=$2"
  FOUNDATION_PT_LIB=$PWD/ab_libs/$1.so timeout 300 python scripts/probe.py --scene $2 --rays 16777216 --reps 3 --spp 16 2>&1 | grep -E "closest|any:|render|Error|error"
}
{
run cur terrain
run sstack2 terrain
run sstack4 terrain
run cur terrain
run sstack2 terrain
run sstack4 terrain
run cur spheres
run sstack2 spheres
run sstack4 spheres
run cur instanced
run sstack2 instanced
run sstack4 instanced
} 2>&1 | tee gpurun_out/ab6.log
