timeout 400 python tests/gpu_k2_compare.py 663000 row 3 > gpurun_out/r2h_cmp663k.log 2>&1; echo "rc663k=$?"
tail -4 gpurun_out/r2h_cmp663k.log
timeout 1200 python -m pytest tests -q -m gpu -x --deselect tests/test_multi_gpu.py > gpurun_out/r2h_tests.log 2>&1; echo "rctests=$?"
tail -8 gpurun_out/r2h_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_launches.csv python tests/gpu_k2_compare.py 663000 row 2 > gpurun_out/r2h_ncu.log 2>&1; echo "rcncu=$?"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "rcbench=$?"
tail -c 600 gpurun_out/r2h_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "rcsmoke=$?"; tail -2 gpurun_out/r2h_smoke.log
