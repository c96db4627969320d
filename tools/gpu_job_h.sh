#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_diffuse -s 3 -c 3 -o gpurun_out/h_prof -f \
   python bench.py --molecules 10000000 --no-cpu --steps 2 --warmup 1 --e2e-calls 1 > gpurun_out/h_ncu.log 2>&1
cp mcell_b200/libmcx.so gpurun_out/h_libmcx.so
