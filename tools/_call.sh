set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s4_pytest_gpu_final2.log 2>&1; echo "rc=$?" >> gpurun_out/s4_pytest_gpu_final2.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4_smoke2.log 2>&1; tail -n 2 gpurun_out/s4_smoke2.log
timeout 600 python bench.py > gpurun_out/s4_bench_final2.log 2>&1
tail -n 3 gpurun_out/s4_pytest_gpu_final2.log; tail -c 400 gpurun_out/s4_bench_final2.log
