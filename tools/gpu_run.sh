# one GPU call: correctness first (the whole -m gpu suite), then timings, then the bench line
timeout 300 python tests/gpu_k2_compare.py 100000 both 2 > gpurun_out/r2d_cmp100k.log 2>&1; echo "rc100k=$?"
timeout 400 python tests/gpu_k2_compare.py 663000 both 3 > gpurun_out/r2d_cmp663k.log 2>&1; echo "rc663k=$?"
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_multi_gpu.py > gpurun_out/r2d_tests.log 2>&1; echo "rctests=$?"
tail -15 gpurun_out/r2d_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches.csv python tests/gpu_k2_compare.py 663000 tiles 2 > gpurun_out/r2d_ncu.log 2>&1; echo "rcncu=$?"
SG_B200_TILE_WARPS=16 timeout 400 python tests/gpu_k2_compare.py 663000 tiles 3 > gpurun_out/r2d_cmp663k_w16.log 2>&1; echo "rc663k16=$?"
timeout 900 python bench.py --steps 2 --warmup 2 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "rcbench=$?"
tail -c 600 gpurun_out/r2d_bench.err
