#!/bin/bash
# 4 GPUs: 2-rank bit-exactness tests (incl. device release), then the strong-scaling bench at N=4 and N=2 with balanced slabs
set -x
mkdir -p gpurun_out
python -m mcell_b200.build > gpurun_out/mg4b_build.log 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/mg4b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/mg4b_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/mg4b_bench_4gpu.json 2> gpurun_out/mg4b_bench_4gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/mg4b_bench_2gpu.json 2> gpurun_out/mg4b_bench_2gpu.err
tail -5 gpurun_out/mg4b_pytest.log
tail -1 gpurun_out/mg4b_bench_4gpu.json | cut -c1-1500; tail -3 gpurun_out/mg4b_bench_4gpu.err; tail -1 gpurun_out/mg4b_bench_2gpu.json | cut -c1-400
