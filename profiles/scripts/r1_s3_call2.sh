set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_s3b.json 2> gpurun_out/bench_s3b.err; tail -c 1500 gpurun_out/bench_s3b.json
python bench.py --steps 2 --warmup 1 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_s3.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch_s3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_hex_stage_line -s 6 -c 1 -o gpurun_out/prof_s3_line -f python profiles/quick_stage_time.py > gpurun_out/ncu_s3_line.log 2>&1
ls -la gpurun_out | tail -4
