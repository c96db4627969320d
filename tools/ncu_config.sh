#!/bin/bash
# ncu --set full of selected kernels of one BASELINE config: tools/ncu_config.sh TAG CONFIG KERNEL_REGEX SKIP COUNT [lib]
TAG=$1; CFG=$2; KREGEX=$3; SKIP=$4; COUNT=$5; LIB=${6:-mcell_b200/libmcx.so}
mkdir -p gpurun_out
MCX_LIB=$PWD/$LIB timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -s $SKIP -c $COUNT \
  -o gpurun_out/${TAG}_prof -f python bench.py --config $CFG --steps 3 --warmup 3 --e2e-calls 1 --no-cpu > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
