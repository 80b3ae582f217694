set -x
timeout 600 python -m pytest tests/test_decode_kernels_gpu.py -x -q -k "half" > gpurun_out/s4_t1.log 2>&1; echo "rc=$?" >> gpurun_out/s4_t1.log
timeout 600 python -m pytest tests/test_e2e_gpu.py -x -q -k "fp16" > gpurun_out/s4_t2.log 2>&1; echo "rc=$?" >> gpurun_out/s4_t2.log
timeout 600 python -m pytest tests/test_fullsize_gpu.py -x -q -s -k "fp16" > gpurun_out/s4_t3.log 2>&1; echo "rc=$?" >> gpurun_out/s4_t3.log
timeout 400 python tools/decode_probe.py 512 prec bf16x3 fp16 bf16 > gpurun_out/s4_prec_probe.log 2>&1
timeout 300 python bench.py --impl reference --variant 16_224 --steps 1 --warmup 1 > gpurun_out/s4_refarm.log 2>&1
tail -3 gpurun_out/s4_t1.log gpurun_out/s4_t2.log gpurun_out/s4_t3.log; cat gpurun_out/s4_prec_probe.log; tail -c 600 gpurun_out/s4_refarm.log
