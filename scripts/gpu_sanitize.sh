#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cornell or instanced" 2>&1 | tail -25 | tee gpurun_out/sanitizer_memcheck.log
echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_build_is_byte_identical_to_oracle and cornell or test_image_matches_oracle and cornell" 2>&1 | tail -12 | tee gpurun_out/sanitizer_racecheck.log
