timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cns.py -x -q 2>&1 | tail -4
for v in 0 1; do
  echo "ESDG_SYM_PAIRS=$v"
  ESDG_SYM_PAIRS=$v python profiles/measure_2d.py 2>&1 | grep -v "^$"
  ESDG_SYM_PAIRS=$v timeout 300 python -m pytest tests/test_gpu_perf_report.py -x -q -s 2>&1 | grep -E "tri N|passed|failed|Error"
done
