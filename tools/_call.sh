timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/s4_bench_n2.log 2>&1
tail -c 600 gpurun_out/s4_bench_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --variant 16_224 --gpus 2 --steps 1 --warmup 1 > gpurun_out/s4_ref_n2.log 2>&1
grep -c '"impl": "reference"' gpurun_out/s4_ref_n2.log
