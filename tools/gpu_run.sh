cd string_grouper_b200/csrc
for v in 0 1; do
touch sg_cossim.cu; make -s EXTRA="-DSG_RESCORE_SEARCH=$v" 2>&1 | grep -E "error"
cd ../..
echo "== rescore search=$v"
timeout 300 python tests/gpu_k2_compare.py 663000 row 3 2>&1 | grep phases | tail -2
timeout 300 python tests/gpu_k2_compare.py 100000 row 3 2>&1 | grep phases | tail -1
cd string_grouper_b200/csrc
done
