#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for th in 0 8 16 20 24 28 32; do
  echo "== thresh $th"; FOUNDATION_PT_FETCH_THRESH=$th timeout 300 python scripts/probe.py --scene terrain --n 2236 --rays 16777216 --reps 2 --spp 2 2>&1 | grep -E "closest|any|render" | tail -3
done
for b in 4 6 8 12; do
  echo "== blocks/sm $b"; FOUNDATION_PT_TRACE_BLOCKS_PER_SM=$b timeout 300 python scripts/probe.py --scene terrain --n 2236 --rays 16777216 --reps 2 2>&1 | grep -E "closest" | tail -1
done
