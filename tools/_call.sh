set -x
ls -la oracle/_ref > gpurun_out/s4_ref_ls.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/s4_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/s4_pytest_gpu.log
timeout 300 python bench.py --impl reference --variant 16_224 --steps 1 --warmup 1 > gpurun_out/s4_refarm.log 2>&1
timeout 600 python bench.py > gpurun_out/s4_bench.log 2>&1
tail -n 4 gpurun_out/s4_pytest_gpu.log; head -c 300 gpurun_out/s4_refarm.log; tail -c 1500 gpurun_out/s4_bench.log
