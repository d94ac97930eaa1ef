# Validation of a tree on one B200 (run as: gpurun --timeout 3000 -- 'bash tools/gpu_run.sh'; ~8 GPU-minutes):
# build + smoke, the whole -m gpu suite, the default bench line, the ncu launch list of the same bench command.
# Everything lands in gpurun_out/ (scratch); copy what should be judged into profiles/ with profiles/summarize.py.
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rcsmoke=$?"; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/tests_gpu.log 2>&1; echo "rctests=$?"
tail -4 gpurun_out/tests_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "rcbench=$?"
tail -c 300 gpurun_out/bench_1gpu.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "rcncu=$?"
