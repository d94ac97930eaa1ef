timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_final_bench_under_ncu.log 2>&1; echo "rcncu=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cossim_candidates -s 1 -c 1 -o gpurun_out/r2_final_k2row python tests/gpu_k2_compare.py 663000 row 1 > gpurun_out/r2_final_ncu1.log 2>&1; echo "rc1=$?"
timeout 300 python bench_configs.py --config 2 --cpu > gpurun_out/r2_config2.json 2> gpurun_out/r2_config2.err; echo "rc2=$?"
timeout 900 python bench_configs.py --config 5 --cpu > gpurun_out/r2_config5.json 2> gpurun_out/r2_config5.err; echo "rc5=$?"
timeout 900 python bench_configs.py --config 4 --cpu > gpurun_out/r2_config4.json 2> gpurun_out/r2_config4.err; echo "rc4=$?"
tail -c 400 gpurun_out/r2_config4.err gpurun_out/r2_config5.err
