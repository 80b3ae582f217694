timeout 300 python -m pytest tests/test_e2e_gpu.py -x -q -k "early or tiny_greedy or beam or sampling" > gpurun_out/s4e_t1.log 2>&1; tail -n 3 gpurun_out/s4e_t1.log
for rep in 1 2; do
VITCAP_EARLY_EXIT_EVERY=1 timeout 200 python tools/decode_probe.py 512 prec fp16 2>&1 | grep round | sed 's/^/every 1: /'
timeout 200 python tools/decode_probe.py 512 prec fp16 2>&1 | grep round | sed 's/^/every 4 (default at 512): /'
done > gpurun_out/s4e_exit_every_ab.log 2>&1
cat gpurun_out/s4e_exit_every_ab.log
