mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2k_pytest.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "triangle_counts or hits_match" 2>&1 | tail -3 | tee -a gpurun_out/r2k_pytest.log
for rep in 1 2; do for lib in ab_libs/head.so ab_libs/walk32.so; do echo -n "$lib: "; FOUNDATION_PT_LIB=$lib timeout 300 python scripts/probe2.py --hash --spp 0 --log2-rays 20 2>&1 | tail -1; done; done | tee gpurun_out/r2k_ab.log
