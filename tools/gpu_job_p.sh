#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu > gpurun_out/p_bench_1e8.json 2> gpurun_out/p_bench_1e8.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/p_launches_cfg4.csv \
   python tools/bench_configs.py --only 4 --iters 4 > gpurun_out/p_cfg4.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_diffuse' -s 24 -c 5 -o gpurun_out/p_prof_cfg4 -f \
   python tools/bench_configs.py --only 4 --iters 4 > gpurun_out/p_ncu.log 2>&1
cp mcell_b200/libmcx.so gpurun_out/p_libmcx.so
python - <<'PY'
import json
d=json.load(open("gpurun_out/p_bench_1e8.json")); r=d["roofline"]
print("1e8 ms/step %.3f fast %.3f slow %.3f resolve %.3f sort %.3f deferred %.4f e2e %.3g"%(d["ms_per_step"], r["ms_diffuse_fast"], r["ms_diffuse_slow"], r["ms_resolve"], r["ms_sort"], r["deferred_fraction"], d["e2e"]["value"]))
PY
python tools/ncu_summary.py launches gpurun_out/p_launches_cfg4.csv
