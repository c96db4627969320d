#!/bin/bash
# Two-pass fast kernel: full GPU tier, bench at 1e7 and 1e8
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/m_pytest.log
timeout 600 python bench.py --molecules 10000000 --no-cpu > gpurun_out/m_bench_1e7.json 2> gpurun_out/m_bench_1e7.err
timeout 600 python bench.py --no-cpu > gpurun_out/m_bench_1e8.json 2> gpurun_out/m_bench_1e8.err
tail -5 gpurun_out/m_pytest.log
for f in gpurun_out/m_bench_1e7.json gpurun_out/m_bench_1e8.json; do python - $f <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]
print(sys.argv[1], "ms/step %.3f fast %.3f slow %.3f resolve %.3f sort %.3f deferred %.4f e2e %.3g"%(d["ms_per_step"], r["ms_diffuse_fast"], r["ms_diffuse_slow"], r["ms_resolve"], r["ms_sort"], r["deferred_fraction"], d["e2e"]["value"]), r["deferred_by_reason"])
PY
done
