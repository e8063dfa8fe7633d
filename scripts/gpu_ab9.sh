#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/ab9_pytest.log
cat gpurun_out/ab9_pytest.log
{
for rep in 1 2; do
for o in 0 1; do
  for sc in terrain spheres instanced; do
    echo "== order=$o scene=$sc"
    FOUNDATION_PT_PIXEL_ORDER=$o timeout 300 python scripts/probe.py --scene $sc --rays 1024 --reps 1 --spp 32 2>&1 | grep -E "render" | tail -1
  done
done
done
} 2>&1 | tee gpurun_out/ab9.log
