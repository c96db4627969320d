#!/bin/bash
# 2 GPUs: the 2-rank bit-exactness tests (volume, surface molecules, counted volumes), peer-memory and NCCL halo paths
set -x
mkdir -p gpurun_out
python -m mcell_b200.build > gpurun_out/mg2_build.log 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/mg2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/mg2_pytest.log
MCX_HALO_NCCL=1 timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/mg2_pytest_nccl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/mg2_pytest_nccl.log
tail -30 gpurun_out/mg2_pytest.log; tail -5 gpurun_out/mg2_pytest_nccl.log
