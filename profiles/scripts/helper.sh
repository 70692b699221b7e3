for v in h00 h11 h1m3; do
  ESDG_B200_LIB=esdg_cns_b200/libesdg_b200_$v.so timeout 200 python profiles/stage_time_n.py 4 128 128 16 2>&1 | tail -1
  ESDG_B200_LIB=esdg_cns_b200/libesdg_b200_$v.so timeout 200 python profiles/stage_time_n.py 3 64 64 64 2>&1 | tail -1
done
