set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s4_pytest_gpu_final.log 2>&1; echo "rc=$?" >> gpurun_out/s4_pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4_smoke.log 2>&1; tail -n 3 gpurun_out/s4_smoke.log
timeout 600 python bench.py > gpurun_out/s4_bench_final.log 2>&1
timeout 400 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/s4_refarm_final.log 2>&1
tail -n 4 gpurun_out/s4_pytest_gpu_final.log; tail -c 800 gpurun_out/s4_bench_final.log; tail -c 300 gpurun_out/s4_refarm_final.log
