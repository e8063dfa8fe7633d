#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (persistent collapse with grid barriers, gather pack / scatter, attribute shading, deforming-mesh rebuild)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_textures.py -m gpu -x -q -k "deforming or textured" 2>&1 | tail -3 | tee gpurun_out/r02_new_tests.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_textures.py -m gpu -x -q -k "cornell or instanced or gather_local or textured or deforming or refit_tiles" 2>&1 | tail -12 | tee gpurun_out/r02_sanitizer_memcheck.log
echo "memcheck rc=${PIPESTATUS[0]}" | tee -a gpurun_out/r02_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -x -q -k "test_build_is_byte_identical_to_oracle or gather_local or (test_image_matches_oracle and cornell)" 2>&1 | tail -10 | tee gpurun_out/r02_sanitizer_racecheck.log
