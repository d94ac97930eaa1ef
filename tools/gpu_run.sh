timeout 1200 python -m pytest tests/test_gpu_golden_synthetic.py tests/test_gpu_tfidf.py tests/test_reference_suite.py tests/test_golden_api.py -q -m gpu > gpurun_out/r2j_tests.log 2>&1; echo "rctests=$?"
tail -40 gpurun_out/r2j_tests.log
