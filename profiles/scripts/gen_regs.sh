for v in g96 g80 g64 g64c64; do
  echo "variant $v"
  ESDG_B200_LIB=esdg_cns_b200/libesdg_b200_$v.so python profiles/measure_2d.py 2>&1 | grep -E "euler quad|generic_rhs|cns quad|cns_rhs"
  ESDG_B200_LIB=esdg_cns_b200/libesdg_b200_$v.so timeout 300 python -m pytest tests/test_gpu_perf_report.py -x -q -s 2>&1 | grep -E "tri N|failed|Error"
done
