#!/bin/bash
mkdir -p gpurun_out
{
for t in 12 16 20 24 28; do
  for sc in terrain spheres; do
    echo "== thresh=$t scene=$sc"
    FOUNDATION_PT_FETCH_THRESH=$t timeout 300 python scripts/probe.py --scene $sc --rays 16777216 --reps 2 --spp 16 2>&1 | grep -E "closest|any:|render" | tail -3
  done
done
} 2>&1 | tee gpurun_out/ab8.log
