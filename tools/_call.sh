timeout 400 python tools/decode_probe.py 512 prec bf16 bf16+fold fp16 > gpurun_out/s4_fold_probe.log 2>&1
cat gpurun_out/s4_fold_probe.log | tail -20
