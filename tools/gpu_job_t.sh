#!/bin/bash
# fast-pass variants at 1e8: streaming stores of the result records, two gathers in flight per probe trip, L1 size
set -x
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 400 python bench.py --no-cpu --e2e-calls 1 --steps 6 --warmup 3 > gpurun_out/t_$name.json 2> gpurun_out/t_$name.err; }
run base A=1
run ss MCX_LIB=$PWD/mcell_b200/libmcx_ss.so
run u2 MCX_LIB=$PWD/mcell_b200/libmcx_u2.so
run ssu2 MCX_LIB=$PWD/mcell_b200/libmcx_ssu2.so
run base_smem100 MCX_FAST_CARVEOUT=100
python - <<'PY'
import json
for n in ["base","ss","u2","ssu2","base_smem100"]:
    try:
        d=json.load(open("gpurun_out/t_%s.json"%n)); r=d["roofline"]
        print("%-14s ms/step %.3f fast %.3f slow %.3f resolve %.3f sort %.3f"%(n,d["ms_per_step"], r["ms_diffuse_fast"], r["ms_diffuse_slow"], r["ms_resolve"], r["ms_sort"]))
    except Exception as e: print(n, "failed", e)
PY
