#!/bin/bash
# final confirmation of HEAD on one GPU: the whole GPU tier and smoke()
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_pytest.log
( timeout 120 python __graft_entry__.py --smoke ) > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/z_smoke.log
tail -6 gpurun_out/z_pytest.log; tail -2 gpurun_out/z_smoke.log
