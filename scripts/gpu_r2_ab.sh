mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2e_pytest.log
bash scripts/ab2.sh "--hash" ab_libs/v_f1.so ab_libs/p_late.so ab_libs/p_early.so 2>&1 | tee gpurun_out/r2e_ab.log
