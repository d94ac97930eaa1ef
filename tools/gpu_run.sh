cd string_grouper_b200/csrc
for v in 0 1; do
touch sg_cossim.cu; make -s EXTRA="-DSG_RESCORE_SEARCH=$v" 2>&1 | grep -E "error"
cd ../..
echo "== rescore search=$v"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:rescore --csv python tests/gpu_k2_compare.py 663000 both 1 2>&1 | grep -E "rescore_kernel|identical" | cut -c1-200 | tail -4
cd string_grouper_b200/csrc
done
cd ../..
timeout 600 python -m pytest tests/test_gpu_cossim.py tests/test_gpu_compat.py -q -m gpu -x 2>&1 | tail -3
