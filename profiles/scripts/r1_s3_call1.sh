set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err; tail -c 600 gpurun_out/bench_s3.json
ESDG_B200_LIB=esdg_cns_b200/libesdg_b200_timing.so python profiles/phase_timing.py 2>&1 | tee gpurun_out/phase_timing.log
python profiles/measure_2d.py 2>&1 | tee gpurun_out/measure_2d_s3.log
python profiles/compare_kernels.py 2>&1 | tee gpurun_out/compare_s3.log
MODE=cns K1D=256 ncu --set full --clock-control none --import-source on -k regex:k_cns_rhs -s 4 -c 1 -o gpurun_out/prof_cns_rhs3 -f python profiles/run_2d_once.py > gpurun_out/ncu_c3.log 2>&1
MODE=cns K1D=256 ncu --set full --clock-control none --import-source on -k regex:k_cns_sigma -s 4 -c 1 -o gpurun_out/prof_cns_sigma3 -f python profiles/run_2d_once.py > gpurun_out/ncu_s3.log 2>&1
ls -la gpurun_out | tail -5
