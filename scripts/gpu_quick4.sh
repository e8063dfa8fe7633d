#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for b in 7 8; do echo "== blocks/sm $b"; FOUNDATION_PT_TRACE_BLOCKS_PER_SM=$b timeout 300 python scripts/probe.py --scene terrain --n 2236 --rays 16777216 --reps 3 --spp 16 2>&1 | grep -E "closest|any|render" | tail -3; done
