for eb in 16 20 24 32; do
  echo "ESDG_GENERIC_EB=$eb"
  ESDG_GENERIC_EB=$eb python profiles/measure_2d.py 2>&1 | grep -A2 "euler quad"
done
