python profiles/measure_2d.py 2>&1 | grep -A2 "euler quad"
for eb in 2 4 6 8 10; do
  echo "ESDG_GENERIC_EB=$eb"
  ESDG_GENERIC_EB=$eb timeout 300 python -m pytest tests/test_gpu_perf_report.py -x -q -s 2>&1 | grep -E "tri N|failed|Error"
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -x -q 2>&1 | tail -2
