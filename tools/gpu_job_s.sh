#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/s_bench_1e8.json 2> gpurun_out/s_bench_1e8.err
timeout 900 python tools/bench_configs.py > gpurun_out/s_configs.jsonl 2> gpurun_out/s_configs.err
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_diffuse|k_scatter|k_resolve|k_compact' -s 30 -c 15 -o gpurun_out/s_prof -f \
   python bench.py --steps 2 --warmup 1 --e2e-calls 1 --no-cpu > gpurun_out/s_ncu.log 2>&1
cp mcell_b200/libmcx.so gpurun_out/s_libmcx.so
tail -4 gpurun_out/s_pytest.log
python - <<'PY'
import json
for l in open("gpurun_out/s_configs.jsonl"):
    d=json.loads(l); print("%-55s %9.3f ms/it  %.3g mol-steps/s  fast0 %.3f  pass1+generic %.3f resolve %.3f sort %.3f deferred %.4f"%(d["config"][:55], d["ms_per_iteration"], d["molecule_steps_per_sec"], d["ms_fast_pass0"], d["ms_pass1_and_generic"], d["ms_resolve"], d["ms_sort"], d["deferred_fraction"]))
d=json.load(open("gpurun_out/s_bench_1e8.json")); r=d["roofline"]
print("1e8 ms/step %.3f fast %.3f slow %.3f resolve %.3f sort %.3f deferred %.4f e2e %.3g"%(d["ms_per_step"], r["ms_diffuse_fast"], r["ms_diffuse_slow"], r["ms_resolve"], r["ms_sort"], r["deferred_fraction"], d["e2e"]["value"]))
PY
