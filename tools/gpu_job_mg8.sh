#!/bin/bash
# 8 GPUs of one box: the strong-scaling bench at N=8 and N=4 (the 2-rank bit-exactness test: tools/gpu_job_mg.sh)
set -x
mkdir -p gpurun_out
python -m mcell_b200.build > gpurun_out/mg8_build.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/mg8_bench_8gpu.json 2> gpurun_out/mg8_bench_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/mg8_bench_4gpu.json 2> gpurun_out/mg8_bench_4gpu.err
tail -1 gpurun_out/mg8_bench_8gpu.json | cut -c1-1800; tail -3 gpurun_out/mg8_bench_8gpu.err; tail -1 gpurun_out/mg8_bench_4gpu.json | cut -c1-1800; tail -3 gpurun_out/mg8_bench_4gpu.err
