timeout 1200 python -m pytest tests/test_gpu_cossim.py tests/test_gpu_golden_synthetic.py -q -m gpu -x > gpurun_out/r2p_tests.log 2>&1; echo "rctests=$?"; tail -15 gpurun_out/r2p_tests.log
timeout 600 python tests/gpu_k2_compare.py 663000 row 3 2>&1 | grep -E "phases|rep 2"
timeout 600 python tests/gpu_k2_compare.py 100000 row 3 2>&1 | grep -E "phases|rep 2" | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --no-aux > gpurun_out/r2p_bench_1gpu.json 2> gpurun_out/r2p_bench_1gpu.err; echo "rcbench=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2p_bench_1gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']['s_each_step'], d['e2e']['value']); print(d.get('phases_ms')); print(d.get('parity'))
P
