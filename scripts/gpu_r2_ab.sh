mkdir -p gpurun_out
bash scripts/ab2.sh "--hash" ab_libs/v_f1.so ab_libs/slim_fb8.so ab_libs/slim_fb9.so ab_libs/slim_late_fb9.so 2>&1 | tee gpurun_out/r2g_ab.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2g_pytest.log
