nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_fullsize.py::test_config5_shape_groups_equal_cpu_port -q -m gpu -x > gpurun_out/r2i_tests_mgpu.log 2>&1; echo "rctests=$?"
tail -12 gpurun_out/r2i_tests_mgpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2i_bench2.json 2> gpurun_out/r2i_bench2.err; echo "rcbench2=$?"
tail -c 1500 gpurun_out/r2i_bench2.err
