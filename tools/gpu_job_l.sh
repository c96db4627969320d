#!/bin/bash
# Full GPU tier + default bench (1e8, with cpu_baseline) + ncu launch list + ncu --set full of the pipeline at 1e8
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
( time timeout 900 python bench.py ) > gpurun_out/l_bench_1e8.json 2> gpurun_out/l_bench_1e8.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/l_launches.csv \
   python bench.py --steps 2 --warmup 1 --e2e-calls 1 --no-cpu > gpurun_out/l_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_diffuse|k_scatter|k_resolve' -s 24 -c 12 -o gpurun_out/l_prof -f \
   python bench.py --steps 2 --warmup 1 --e2e-calls 1 --no-cpu > gpurun_out/l_ncu.log 2>&1
cp mcell_b200/libmcx.so gpurun_out/l_libmcx.so
tail -12 gpurun_out/l_pytest.log; cat gpurun_out/l_bench_1e8.json | cut -c1-600
