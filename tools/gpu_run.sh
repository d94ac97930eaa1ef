timeout 300 python tests/gpu_k2_compare.py 100000 both 2 > gpurun_out/r2d_cmp100k.log 2>&1; echo "rc100k=$?"
timeout 400 python tests/gpu_k2_compare.py 663000 both 3 > gpurun_out/r2d_cmp663k.log 2>&1; echo "rc663k=$?"
SG_B200_TILE_WARPS=16 timeout 400 python tests/gpu_k2_compare.py 663000 tiles 3 > gpurun_out/r2d_cmp663k_w16.log 2>&1; echo "rc663k16=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches.csv python tests/gpu_k2_compare.py 663000 tiles 2 > gpurun_out/r2d_ncu.log 2>&1; echo "rcncu=$?"
timeout 900 python -m pytest tests/test_gpu_cossim.py tests/test_gpu_tfidf.py tests/test_gpu_compat.py tests/test_reference_suite.py tests/test_golden_api.py -x -q -m gpu > gpurun_out/r2d_tests.log 2>&1; echo "rctests=$?"
tail -5 gpurun_out/r2d_tests.log
