timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 --no-aux > gpurun_out/r2s_bench_8gpu.json 2> gpurun_out/r2s_bench_8gpu.err; echo "rcbench=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2s_bench_8gpu.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','shard_check')}); print(d['e2e']['s_each_step'], d['e2e']['value']); print(d.get('phases_ms',{}).get('rank0'))
P
