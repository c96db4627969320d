#!/bin/bash
# One GPU call: parity tests, bench at 1e7, ncu --set full of both diffuse kernels at 1e7
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --molecules 10000000 --no-cpu > gpurun_out/a_bench_1e7.json 2> gpurun_out/a_bench_1e7.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_diffuse -s 6 -c 2 -o gpurun_out/a_prof -f \
   python bench.py --molecules 10000000 --no-cpu --steps 2 --warmup 1 --e2e-calls 1 > gpurun_out/a_ncu.log 2>&1
tail -3 gpurun_out/a_pytest.log; cat gpurun_out/a_bench_1e7.json
