nvidia-smi -L | wc -l
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/r2_bench_${n}gpu.json 2> gpurun_out/r2_bench_${n}gpu.err; echo "rcbench$n=$?"
tail -c 300 gpurun_out/r2_bench_${n}gpu.err | tail -3
done
