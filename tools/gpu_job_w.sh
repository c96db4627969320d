#!/bin/bash
# same-box A/B of the rolled wall loops (w) and rolled Gaussian draws (g) of the fast pass: 1e8 twice each, configs 1-4 once
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_release.py -m gpu -x -q > gpurun_out/w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/w_pytest.log
for rep in 1 2; do
  for v in w0g0 w1g0 w1g1 w0g1; do
    MCX_LIB=$PWD/mcell_b200/libmcx_$v.so timeout 400 python bench.py --no-cpu --e2e-calls 1 --steps 6 --warmup 3 > gpurun_out/w_${v}_$rep.json 2> gpurun_out/w_${v}_$rep.err
  done
done
for v in w0g0 w1g0 w1g1 w0g1; do
  MCX_LIB=$PWD/mcell_b200/libmcx_$v.so timeout 600 python tools/bench_configs.py > gpurun_out/w_configs_$v.jsonl 2> gpurun_out/w_configs_$v.err
done
tail -3 gpurun_out/w_pytest.log
python - <<'PY'
import json
for v in ["w0g0","w1g0","w1g1","w0g1"]:
    for rep in (1,2):
        try:
            d=json.load(open("gpurun_out/w_%s_%d.json"%(v,rep))); r=d["roofline"]
            print("%s #%d 1e8 ms/step %.3f fast %.3f slow %.3f resolve %.3f sort %.3f"%(v,rep,d["ms_per_step"], r["ms_diffuse_fast"], r["ms_diffuse_slow"], r["ms_resolve"], r["ms_sort"]))
        except Exception as e: print(v,rep,"failed",e)
    try:
        for l in open("gpurun_out/w_configs_%s.jsonl"%v):
            d=json.loads(l); print("   %-50s %9.3f ms/it fast0 %.3f  pass1+generic %.3f"%(d["config"][:50], d["ms_per_iteration"], d["ms_fast_pass0"], d["ms_pass1_and_generic"]))
    except Exception as e: print(v,"configs failed",e)
PY
