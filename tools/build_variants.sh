#!/bin/bash
# Builds tuning variants of libmcx.so side by side (mcell_b200/libmcx_<name>.so); select one with MCX_LIB=<path>.
# usage: tools/build_variants.sh name1="-DFLAG ..." name2="..."
set -e
cd "$(dirname "$0")/.."
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  MCX_NVCC_EXTRA="$flags" python -m mcell_b200.build --force > /dev/null
  cp mcell_b200/libmcx.so "mcell_b200/libmcx_${name}.so"
  echo "built libmcx_${name}.so with: $flags"
done
python -m mcell_b200.build --force > /dev/null
