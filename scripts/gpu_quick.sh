#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for th in 16 24 28; do
echo "== thresh $th"; FOUNDATION_PT_FETCH_THRESH=$th timeout 300 python scripts/probe.py --scene terrain --n 2236 --rays 16777216 --reps 3 --spp 4 2>&1 | grep -E "closest|any|render|commit" | tail -4
done
timeout 300 python scripts/probe.py --scene spheres --rays 16777216 --reps 2 --spp 4 2>&1 | grep -E "closest|any|render|commit" | tail -4
timeout 300 python scripts/probe.py --scene instanced --rays 16777216 --reps 2 --spp 4 2>&1 | grep -E "closest|any|render|commit" | tail -4
