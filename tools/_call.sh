timeout 300 python -m pytest tests/test_decode_kernels_gpu.py -x -q -k "vocab_argmax" > gpurun_out/s4b_t1.log 2>&1; tail -n 3 gpurun_out/s4b_t1.log
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_decode_kernels_gpu.py -x -q -k "vocab_argmax" > gpurun_out/s4b_memcheck.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/s4b_memcheck.log
timeout 200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 5 python -m pytest tests/test_decode_kernels_gpu.py -x -q -k "vocab_argmax_half" > gpurun_out/s4b_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed|Race reported|and " gpurun_out/s4b_racecheck.log | sort | uniq -c | head
for rep in 1 2; do
VITCAP_LIB=$PWD/vitcap_b200/lib/libvitcap_b200_old.so timeout 200 python tools/decode_probe.py 512 prec fp16 bf16x3 2>&1 | grep round | sed 's/^/old lib: /'
timeout 200 python tools/decode_probe.py 512 prec fp16 bf16x3 2>&1 | grep round | sed 's/^/new lib: /'
done > gpurun_out/s4b_vocab_bias_ab.log 2>&1
cat gpurun_out/s4b_vocab_bias_ab.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_dec" -c 40 --csv --log-file gpurun_out/s4b_dec_launches.csv python tools/decode_probe.py 512 eager > /dev/null 2>&1
grep "4, 208" gpurun_out/s4b_dec_launches.csv | head -3
