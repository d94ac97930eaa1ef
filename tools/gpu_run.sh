python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "rcsmoke=$?"; tail -1 gpurun_out/r2_final_smoke.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_tests_final3.log 2>&1; echo "rctests=$?"
tail -4 gpurun_out/r2_tests_final3.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_1gpu_final3.json 2> gpurun_out/r2_bench_1gpu_final3.err; echo "rcbench=$?"
tail -c 300 gpurun_out/r2_bench_1gpu_final3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_final3_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_final3_bench_under_ncu.log 2>&1; echo "rcncu1=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rescore_refined -c 1 -f -o gpurun_out/r2_rescore_refined python tests/gpu_k2_compare.py 663000 row 1 > gpurun_out/r2_rescore_refined_run.log 2>&1; echo "rcncu2=$?"
