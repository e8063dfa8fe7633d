#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for np in 0 1; do
echo "== no_l2_pin $np"; if [ $np = 1 ]; then export FOUNDATION_PT_NO_L2_PIN=1; else unset FOUNDATION_PT_NO_L2_PIN; fi
timeout 300 python scripts/probe.py --scene terrain --n 2236 --rays 16777216 --reps 3 --spp 16 2>&1 | grep -E "closest|any|render" | tail -3
done
unset FOUNDATION_PT_NO_L2_PIN
timeout 300 python scripts/probe.py --scene terrain --n 2236 --rays 67108864 --reps 3 2>&1 | grep -E "closest|any|render" | tail -3
