#!/bin/bash
# compute-sanitizer over the one-sweep radix sort (shared-memory atomicOr ranking, decoupled look-back) through the build tests
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_build_is_byte_identical_to_oracle or refit_tiles" 2>&1 | tail -8 | tee gpurun_out/r02_sanitizer_racecheck_sort.log
echo "racecheck rc=${PIPESTATUS[0]}" | tee -a gpurun_out/r02_sanitizer_racecheck_sort.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_build_is_byte_identical_to_oracle or refit_tiles or chain" 2>&1 | tail -8 | tee gpurun_out/r02_sanitizer_memcheck_sort.log
echo "memcheck rc=${PIPESTATUS[0]}" | tee -a gpurun_out/r02_sanitizer_memcheck_sort.log
