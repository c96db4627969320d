#!/bin/bash
# cell-shape sweep of the fast pass at 1e8: CLAMP_R AX AYZ CELL_EDGE
mkdir -p gpurun_out
for cfg in "4.0 2.25 1.5 0" "3.2 2.25 1.5 1.8" "3.2 1.8 1.34 1.8" "3.5 2.25 1.5 2.0" "3.2 2.25 1.7 1.8" "2.8 2.25 1.5 1.6"; do
  set -- $cfg
  MCX_CELL_CLAMP_R=$1 MCX_CELL_AX=$2 MCX_CELL_AYZ=$3 timeout 300 python bench.py --no-cpu --e2e-calls 1 --steps 6 --warmup 3 --cell-edge $4 > gpurun_out/r2r_sweep.json 2>> gpurun_out/r2r_sweep.err
  python - "$cfg" <<PY
import json,sys
d=json.load(open("gpurun_out/r2r_sweep.json")); r=d["roofline"]
print("SWEEP %s: ms/step %.3f fast %.3f slow %.3f resolve %.3f sort %.3f deferred %.4f" % (sys.argv[1], d["ms_per_step"], r["ms_diffuse_fast"], r["ms_diffuse_slow"], r["ms_resolve"], r["ms_sort"], r["deferred_fraction"]))
PY
done
