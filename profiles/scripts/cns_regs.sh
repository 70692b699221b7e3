for v in c128 c104 c96 c80; do
  echo "variant $v"
  ESDG_B200_LIB=esdg_cns_b200/libesdg_b200_$v.so python profiles/measure_2d.py 2>&1 | grep -A3 "cns quad"
done
