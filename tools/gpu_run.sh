timeout 400 python tests/gpu_k2_compare.py 663000 both 3 > gpurun_out/r2g_cmp663k.log 2>&1; echo "rc663k=$?"
tail -8 gpurun_out/r2g_cmp663k.log
timeout 1200 python -m pytest tests -q -m gpu -x --deselect tests/test_multi_gpu.py > gpurun_out/r2g_tests.log 2>&1; echo "rctests=$?"
tail -15 gpurun_out/r2g_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches.csv python tests/gpu_k2_compare.py 663000 row 2 > gpurun_out/r2g_ncu.log 2>&1; echo "rcncu=$?"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "rcbench=$?"
tail -c 600 gpurun_out/r2g_bench.err
