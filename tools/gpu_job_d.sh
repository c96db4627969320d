#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_diffuse -s 2 -c 2 -o gpurun_out/d_prof -f \
   python bench.py --molecules 10000000 --no-cpu --steps 2 --warmup 1 --e2e-calls 1 > gpurun_out/d_ncu.log 2>&1
