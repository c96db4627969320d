#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q ) > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
timeout 600 python bench.py --no-cpu > gpurun_out/n_bench_1e8.json 2> gpurun_out/n_bench_1e8.err
timeout 900 python tools/bench_configs.py > gpurun_out/n_configs.jsonl 2> gpurun_out/n_configs.err
tail -3 gpurun_out/n_pytest.log
python - gpurun_out/n_bench_1e8.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]
print(sys.argv[1], "ms/step %.3f fast %.3f slow %.3f resolve %.3f sort %.3f deferred %.4f e2e %.3g"%(d["ms_per_step"], r["ms_diffuse_fast"], r["ms_diffuse_slow"], r["ms_resolve"], r["ms_sort"], r["deferred_fraction"], d["e2e"]["value"]), r["deferred_by_reason"])
PY
cat gpurun_out/n_configs.jsonl; tail -3 gpurun_out/n_configs.err
