for tw in 128 256 192; do
SG_B200_TILE_W=$tw timeout 400 python tests/gpu_k2_compare.py 663000 row 3 > gpurun_out/r2k_cmp663k_tw$tw.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r2k_cmp663k_tw$tw.log
done
for pr in 0.8 0.95; do
SG_B200_PRUNE=$pr timeout 400 python tests/gpu_k2_compare.py 663000 row 2 > gpurun_out/r2k_cmp663k_pr$pr.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r2k_cmp663k_pr$pr.log
done
SG_B200_TILE_W=128 timeout 300 python tests/gpu_k2_compare.py 100000 row 3 > gpurun_out/r2k_cmp100k_tw128.log 2>&1; tail -1 gpurun_out/r2k_cmp100k_tw128.log
timeout 300 python tests/gpu_k2_compare.py 100000 row 3 > gpurun_out/r2k_cmp100k_tw256.log 2>&1; tail -1 gpurun_out/r2k_cmp100k_tw256.log
