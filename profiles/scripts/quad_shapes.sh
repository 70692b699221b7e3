run() { echo "$@"; env "$@" python profiles/measure_2d.py 2>&1 | grep -A2 "euler quad" | grep -E "euler|generic_rhs"; }
run A=0
run ESDG_B200_LIB=esdg_cns_b200/libesdg_b200_q80.so ESDG_GENERIC_EB=16 ESDG_GENERIC_THREADS=256
run ESDG_B200_LIB=esdg_cns_b200/libesdg_b200_q64.so ESDG_GENERIC_EB=16 ESDG_GENERIC_THREADS=256
run ESDG_B200_LIB=esdg_cns_b200/libesdg_b200_q64.so ESDG_GENERIC_EB=12 ESDG_GENERIC_THREADS=192
run ESDG_B200_LIB=esdg_cns_b200/libesdg_b200_q80.so ESDG_GENERIC_EB=12 ESDG_GENERIC_THREADS=192
run ESDG_GENERIC_EB=16 ESDG_GENERIC_THREADS=256
