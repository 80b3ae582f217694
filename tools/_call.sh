timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/r02_launches_bench_final_s4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02_launches_bench_final_s4.log 2>&1
wc -l gpurun_out/r02_launches_bench_final_s4.csv
