#!/bin/bash
# Round-end rehearsal on one GPU: full GPU tier, smoke(), default bench (1e8 + cpu_baseline), reference arm,
# ncu launch list and ncu --set full of the same bench command
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest.log
( time timeout 300 python __graft_entry__.py --smoke ) > gpurun_out/x_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/x_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/x_bench_1e8.json 2> gpurun_out/x_bench_1e8.err
( time timeout 900 python bench.py --impl reference --steps 10 --warmup 3 ) > gpurun_out/x_bench_reference.json 2> gpurun_out/x_bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/x_launches.csv \
   python bench.py --steps 2 --warmup 1 --e2e-calls 1 --no-cpu > gpurun_out/x_launches.log 2>&1
# capture window: the upload launches 1 matching kernel (k_scatter of the initial sort), every iteration 15 (pass 0,
# compaction, pass 1, 3 generic launches, 8 k_resolve, k_scatter): skip 1 + 3 x 15 and take the whole 4th iteration — a
# steady-state one (the first iteration after an upload sends every molecule to pass 1, profiles/r01_x_* fell on it)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_diffuse|k_scatter|k_resolve|k_compact' -s 46 -c 15 -o gpurun_out/x_prof -f \
   python bench.py --steps 3 --warmup 2 --e2e-calls 1 --no-cpu > gpurun_out/x_ncu.log 2>&1
timeout 900 python tools/bench_configs.py > gpurun_out/x_configs.jsonl 2> gpurun_out/x_configs.err
cp mcell_b200/libmcx.so gpurun_out/x_libmcx.so
tail -4 gpurun_out/x_pytest.log; tail -4 gpurun_out/x_smoke.log; cut -c1-400 gpurun_out/x_bench_1e8.json; tail -3 gpurun_out/x_bench_1e8.err; cut -c1-300 gpurun_out/x_bench_reference.json; tail -3 gpurun_out/x_bench_reference.err
