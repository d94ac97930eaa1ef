cd string_grouper_b200/csrc
for cfg in "1 1" "2 4" "4 8" "1 8" "4 1" "2 8"; do
set -- $cfg
touch sg_cossim.cu; make -s EXTRA="-DSG_WALK_MLP=$1 -DSG_FILTER_MLP=$2" 2>&1 | grep -E "error" 
cd ../..
echo "== walk=$1 filter=$2"
timeout 300 python tests/gpu_k2_compare.py 663000 row 3 2>&1 | tail -2 | cut -c1-110
cd string_grouper_b200/csrc
done
