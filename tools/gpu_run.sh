timeout 300 python tests/gpu_k2_compare.py 100000 both 2 > gpurun_out/r2f_cmp100k.log 2>&1; echo "rc100k=$?"
timeout 400 python tests/gpu_k2_compare.py 663000 both 3 > gpurun_out/r2f_cmp663k.log 2>&1; echo "rc663k=$?"
cat gpurun_out/r2f_cmp663k.log | tail -8
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches.csv python tests/gpu_k2_compare.py 663000 tiles 2 > gpurun_out/r2f_ncu.log 2>&1; echo "rcncu=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_candidates -s 1 -c 1 -o gpurun_out/r2f_tilecand python tests/gpu_k2_compare.py 663000 tiles 1 > gpurun_out/r2f_ncu1.log 2>&1; echo "rc1=$?"
timeout 900 python -m pytest tests/test_gpu_cossim.py tests/test_gpu_compat.py -q -m gpu -x > gpurun_out/r2f_tests.log 2>&1; echo "rctests=$?"
tail -5 gpurun_out/r2f_tests.log
