MODE=cns K1D=256 ncu --set full --clock-control none --import-source on -k regex:k_cns_rhs -s 4 -c 1 -o gpurun_out/prof_cns_rhs4 -f python profiles/run_2d_once.py > gpurun_out/ncu_c4.log 2>&1
MODE=cns K1D=256 ncu --set full --clock-control none --import-source on -k regex:k_cns_sigma -s 4 -c 1 -o gpurun_out/prof_cns_sigma4 -f python profiles/run_2d_once.py > gpurun_out/ncu_s4.log 2>&1
MODE=euler K1D=256 ncu --set full --clock-control none --import-source on -k regex:k_generic_rhs -s 4 -c 1 -o gpurun_out/prof_generic_rhs4 -f python profiles/run_2d_once.py > gpurun_out/ncu_g4.log 2>&1
ls -la gpurun_out/*4.ncu-rep
