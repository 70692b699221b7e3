# usage: bash profiles/scripts/variants.sh name1 name2 ...   (libs esdg_cns_b200/libesdg_b200_<name>.so)
for v in "$@"; do
  ESDG_B200_LIB=esdg_cns_b200/libesdg_b200_$v.so python profiles/quick_stage_time.py 2>&1 | tail -1
done
