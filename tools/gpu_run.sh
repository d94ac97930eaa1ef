python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "rcsmoke=$?"; tail -1 gpurun_out/r2_final_smoke.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_tests_final2.log 2>&1; echo "rctests=$?"
tail -4 gpurun_out/r2_tests_final2.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_1gpu_final.json 2> gpurun_out/r2_bench_1gpu_final.err; echo "rcbench=$?"
tail -c 300 gpurun_out/r2_bench_1gpu_final.err
