#!/bin/bash
# scripts/mkvariant.sh <name> [-DFLAG=..]... : compile an A/B variant of the library into ab_libs/<name>.so and print the
# ptxas resource lines of the traversal kernels (registers / spills)
name="$1"; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-mfma,-fvisibility=hidden -shared \
  -Xptxas -v "$@" -o ab_libs/$name.so foundation_b200/csrc/foundation_pt.cu 2> ab_libs/$name.ptxas.log
rc=$?
grep -A1 -E "Compiling entry function '_Z(12k_trace_raysILb[01]ELb0ELb0E|8k_extendILb0ELb0E|9k_connectILb0E)" ab_libs/$name.ptxas.log | grep -E "registers|spill|Compiling" | sed -e 's/ptxas info    : //' | paste - - | sed -e "s/Compiling entry function '\([^']*\)' for 'sm_100a'/\1/" | cut -c1-220
exit $rc
