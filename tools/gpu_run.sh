timeout 400 python tests/gpu_k2_compare.py 663000 both 3 > gpurun_out/r2l_cmp663k.log 2>&1; echo "rc663k=$?"
tail -8 gpurun_out/r2l_cmp663k.log
timeout 300 python tests/gpu_k2_compare.py 100000 both 2 > gpurun_out/r2l_cmp100k.log 2>&1; echo "rc100k=$?"
tail -3 gpurun_out/r2l_cmp100k.log
timeout 900 python -m pytest tests/test_gpu_cossim.py tests/test_gpu_compat.py tests/test_gpu_fullsize.py::test_full_size_kernels_and_pruning_levels_agree -q -m gpu -x > gpurun_out/r2l_tests.log 2>&1; echo "rctests=$?"
tail -6 gpurun_out/r2l_tests.log
