#!/bin/bash
# device release + viz dump tests, then timing of the current build at 1e8 and configs 1-4
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_release.py tests/test_host_adapter.py -m gpu -x -q > gpurun_out/v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest.log
timeout 400 python bench.py --no-cpu --e2e-calls 1 --steps 6 --warmup 3 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
timeout 600 python tools/bench_configs.py > gpurun_out/v_configs.jsonl 2> gpurun_out/v_configs.err
tail -30 gpurun_out/v_pytest.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/v_bench.json")); r=d["roofline"]
print("1e8 ms/step %.3f fast %.3f slow %.3f resolve %.3f sort %.3f"%(d["ms_per_step"], r["ms_diffuse_fast"], r["ms_diffuse_slow"], r["ms_resolve"], r["ms_sort"]))
for l in open("gpurun_out/v_configs.jsonl"):
    d=json.loads(l); print("%-55s %9.3f ms/it fast0 %.3f  pass1+generic %.3f"%(d["config"][:55], d["ms_per_iteration"], d["ms_fast_pass0"], d["ms_pass1_and_generic"]))
PY
