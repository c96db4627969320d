#!/bin/bash
# same-box A/B: Ziggurat rejection branch out of line (gslow), three resident blocks per SM for the fast passes (mb3)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_release.py -m gpu -x -q > gpurun_out/y_pytest_base.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest_base.log
MCX_LIB=$PWD/mcell_b200/libmcx_gslow.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/y_pytest_gslow.log 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest_gslow.log
for v in base gslow mb3 gslowmb3 base; do
  MCX_LIB=$PWD/mcell_b200/libmcx_$v.so timeout 300 python bench.py --no-cpu --e2e-calls 1 --steps 6 --warmup 3 > gpurun_out/y_${v}_$RANDOM.json 2> gpurun_out/y_$v.err
done
tail -2 gpurun_out/y_pytest_base.log; tail -2 gpurun_out/y_pytest_gslow.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/y_*_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print("%-34s ms/step %.3f fast %.3f slow %.3f resolve %.3f sort %.3f"%(f.split('/')[-1],d["ms_per_step"], r["ms_diffuse_fast"], r["ms_diffuse_slow"], r["ms_resolve"], r["ms_sort"]))
    except Exception as e: print(f,"failed",e)
PY
