#!/bin/bash
# compute-sanitizer over the last session's kernel changes: refit with 128 threads per 256-leaf tile, collapse greedy assignment (racecheck: shared-memory rounds / butterflies),
# two waves in flight and the one-sweep sort (memcheck)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_build_is_byte_identical_to_oracle or refit_tiles or chain" 2>&1 | tail -6 | tee gpurun_out/r02_sanitizer_racecheck_final.log
echo "racecheck rc=${PIPESTATUS[0]}" | tee -a gpurun_out/r02_sanitizer_racecheck_final.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_textures.py -m gpu -x -q -k "two_waves or test_build_is_byte_identical_to_oracle or refit_tiles or textured or (test_image_matches_oracle and cornell)" 2>&1 | tail -6 | tee gpurun_out/r02_sanitizer_memcheck_final.log
echo "memcheck rc=${PIPESTATUS[0]}" | tee -a gpurun_out/r02_sanitizer_memcheck_final.log
