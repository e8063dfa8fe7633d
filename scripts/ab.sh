#!/bin/bash
# same-box A/B of variant libraries: scripts/ab.sh "<probe args>" libA.so libB.so ...   (each run twice, interleaved)
args="$1"; shift
for rep in 1 2; do
  for lib in "$@"; do
    echo "== $lib (pass $rep)"
    FOUNDATION_PT_LIB=$lib timeout 600 python scripts/probe.py $args 2>&1 | grep -E "commit|closest|any:|render"
  done
done
