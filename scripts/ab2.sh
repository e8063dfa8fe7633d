#!/bin/bash
# same-box A/B with scripts/probe2.py: scripts/ab2.sh "<probe2 args>" libA.so libB.so ...   (each twice, interleaved)
args="$1"; shift
for rep in 1 2; do
  for lib in "$@"; do
    echo -n "$lib: "
    FOUNDATION_PT_LIB=$lib timeout 600 python scripts/probe2.py $args 2>&1 | tail -1
  done
done
