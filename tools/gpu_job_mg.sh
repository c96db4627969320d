#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/mg_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/mg_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/mg_bench_2gpu.json 2> gpurun_out/mg_bench_2gpu.err
tail -5 gpurun_out/mg_pytest.log; tail -2 gpurun_out/mg_bench_2gpu.json | cut -c1-600; tail -5 gpurun_out/mg_bench_2gpu.err
