timeout 300 python -m pytest tests/test_decode_kernels_gpu.py tests/test_e2e_gpu.py -x -q -k "finish_ln or fp16 or tiny_greedy or early" > gpurun_out/s4c_t1.log 2>&1; tail -n 3 gpurun_out/s4c_t1.log
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_decode_kernels_gpu.py -x -q -k "finish_ln" > gpurun_out/s4c_memcheck.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/s4c_memcheck.log
timeout 200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 5 python -m pytest tests/test_decode_kernels_gpu.py -x -q -k "finish_ln" > gpurun_out/s4c_racecheck.log 2>&1; grep -E "RACECHECK SUMMARY|passed|failed|Race reported" gpurun_out/s4c_racecheck.log | sort | uniq -c | head
for rep in 1 2; do
VITCAP_LIB=$PWD/vitcap_b200/lib/libvitcap_b200_old.so timeout 200 python tools/decode_probe.py 512 prec fp16 2>&1 | grep round | sed 's/^/old lib: /'
timeout 200 python tools/decode_probe.py 512 prec fp16 2>&1 | grep round | sed 's/^/new lib: /'
done > gpurun_out/s4c_finish_ab.log 2>&1
cat gpurun_out/s4c_finish_ab.log
timeout 120 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg --clock-control none -k regex:"finish_ln" -c 12 --csv --log-file gpurun_out/s4c_fin_launches.csv python tools/decode_probe.py 512 eager > /dev/null 2>&1
grep "finish_ln" gpurun_out/s4c_fin_launches.csv | awk -F'","' '{print $5, $13, $15}' | cut -c1-40,160-260 | head -8
