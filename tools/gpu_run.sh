timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "rcref=$?"
tail -c 300 gpurun_out/r2_bench_reference_arm.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "rcbench=$?"
tail -c 300 gpurun_out/r2_bench_1gpu.err
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_tests_final.log 2>&1; echo "rctests=$?"
tail -5 gpurun_out/r2_tests_final.log
