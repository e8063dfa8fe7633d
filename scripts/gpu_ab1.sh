#!/bin/bash
# one-box experiment: GPU parity tests on the tree's library, then same-box A/B of the variant libraries under ab_libs/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab1_pytest.log
cat gpurun_out/ab1_pytest.log
run() {  # name lib thresh scene extra
  echo "== $1 lib=$2 thresh=$3 scene=$4"
  FOUNDATION_PT_LIB=$PWD/ab_libs/$2.so FOUNDATION_PT_FETCH_THRESH=$3 timeout 300 python scripts/probe.py --scene $4 --rays 16777216 --reps 3 --spp 16 2>&1 | grep -E "commit|closest|any:|render|Error|error"
}
{
run base base 24 terrain
run pref pref 24 terrain
run pref28 pref 28 terrain
run pref32 pref 32 terrain
run pref20 pref 20 terrain
run c64 pref_c64 28 terrain
run ct06 pref_ct06 24 terrain
run ct015 pref_ct015 24 terrain
run base base 24 terrain
run pref pref 24 terrain
run pref28 pref 28 terrain
run base base 24 spheres
run pref pref 24 spheres
run base base 24 instanced
run pref pref 24 instanced
} 2>&1 | tee gpurun_out/ab1.log
