timeout 1200 python -m pytest tests/test_gpu_cossim.py tests/test_gpu_golden_synthetic.py tests/test_gpu_compat.py -q -m gpu -x > gpurun_out/r2q_tests.log 2>&1; echo "rctests=$?"; tail -5 gpurun_out/r2q_tests.log
timeout 600 python tests/gpu_k2_compare.py 663000 row 3 2>&1 | grep -E "phases|rep 2"
timeout 600 python tests/gpu_k2_compare.py 100000 row 3 2>&1 | grep -E "phases|rep 2" | tail -2
