#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pf in 0 1; do for th in 20 24 28; do
echo "== prefetch $pf thresh $th"; FOUNDATION_PT_PREFETCH=$pf FOUNDATION_PT_FETCH_THRESH=$th timeout 300 python scripts/probe.py --scene terrain --n 2236 --rays 16777216 --reps 3 --spp 16 2>&1 | grep -E "closest|any|render" | tail -3
done; done
