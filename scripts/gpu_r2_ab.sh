#!/bin/bash
# same-box A/B of the traversal variants under ab_libs/ (scripts/mkvariant.sh): two interleaved passes of scripts/probe2.py
# usage: scripts/gpu_r2_ab.sh <log name> <variant> [<variant> ...]
mkdir -p gpurun_out
log=gpurun_out/$1; shift
: > $log
for rep in 1 2; do for v in "$@"; do echo -n "$v: "; FOUNDATION_PT_LIB=ab_libs/$v.so timeout ${PROBE_TIMEOUT:-150} python scripts/probe2.py ${PROBE_ARGS:---builds 1 --spp 8 --log2-rays 24} 2>&1 | tail -1; done; done | tee -a $log
