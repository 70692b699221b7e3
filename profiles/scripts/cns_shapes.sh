# CNS launch-shape sweep (quad N=3, 512^2): elements per CTA x threads per CTA
for cfg in "4 128" "4 64" "2 64"; do
  set -- $cfg
  echo "EB=$1 threads=$2"
  ESDG_CNS_EB=$1 ESDG_CNS_THREADS=$2 python profiles/measure_2d.py 2>&1 | grep -A3 "cns quad"
done
