#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
timeout 600 python bench.py --molecules 10000000 --no-cpu > gpurun_out/b_bench_1e7.json 2> gpurun_out/b_bench_1e7.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_diffuse -s 2 -c 2 -o gpurun_out/b_prof -f \
   python bench.py --molecules 10000000 --no-cpu --steps 2 --warmup 1 --e2e-calls 1 > gpurun_out/b_ncu.log 2>&1
timeout 600 python bench.py --molecules 100000000 --no-cpu > gpurun_out/b_bench_1e8.json 2> gpurun_out/b_bench_1e8.err
# variant: 3 resident blocks (80 registers)
MCX_NVCC_EXTRA="-DMCX_FAST_MINBLOCKS=3" python -m mcell_b200.build --force > gpurun_out/b_build3.log 2>&1
timeout 600 python bench.py --molecules 10000000 --no-cpu > gpurun_out/b_bench_1e7_mb3.json 2> gpurun_out/b_bench_1e7_mb3.err
tail -3 gpurun_out/b_pytest.log; cat gpurun_out/b_bench_1e7.json gpurun_out/b_bench_1e7_mb3.json
