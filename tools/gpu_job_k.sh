#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.log
timeout 600 python bench.py --molecules 10000000 --no-cpu > gpurun_out/k_bench_1e7.json 2> gpurun_out/k_bench_1e7.err
MCX_NVCC_EXTRA="-DMCX_FAST_MINBLOCKS=5" python -m mcell_b200.build --force > gpurun_out/k_build5.log 2>&1
timeout 600 python bench.py --molecules 10000000 --no-cpu > gpurun_out/k_bench_1e7_f5.json 2> gpurun_out/k_bench_1e7_f5.err
tail -2 gpurun_out/k_pytest.log
for f in gpurun_out/k_bench_1e7.json gpurun_out/k_bench_1e7_f5.json; do python - $f <<'PY'
import json,sys
d=json.load(open(sys.argv[1])); r=d["roofline"]
print(sys.argv[1], "ms/step %.3f fast %.3f slow %.3f resolve %.3f sort %.3f deferred %.4f"%(d["ms_per_step"], r["ms_diffuse_fast"], r["ms_diffuse_slow"], r["ms_resolve"], r["ms_sort"], r["deferred_fraction"]))
PY
done
