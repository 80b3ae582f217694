timeout 600 python bench.py --no-cpu-baseline > gpurun_out/s4f_bench.log 2>&1; tail -c 300 gpurun_out/s4f_bench.log
