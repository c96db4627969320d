#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log
timeout 600 python bench.py --molecules 10000000 --no-cpu > gpurun_out/g_bench_1e7.json 2> gpurun_out/g_bench_1e7.err
tail -30 gpurun_out/g_pytest.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/g_bench_1e7.json")); r=d["roofline"]
print("ms/step %.3f fast %.3f slow %.3f resolve %.3f sort %.3f deferred %.4f"%(d["ms_per_step"], r["ms_diffuse_fast"], r["ms_diffuse_slow"], r["ms_resolve"], r["ms_sort"], r["deferred_fraction"]), r["deferred_by_reason"])
PY
