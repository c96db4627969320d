#!/bin/bash
# 2 GPUs: the 2-rank bit-exactness test and the N=2 bench, halo refresh over peer memory vs NCCL send/recv
set -x
mkdir -p gpurun_out
python -m mcell_b200.build > gpurun_out/mg_build.log 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/mg_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/mg_pytest.log

timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/mg_bench_2gpu.json 2> gpurun_out/mg_bench_2gpu.err
MCX_HALO_NCCL=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/mg_bench_2gpu_nccl.json 2> gpurun_out/mg_bench_2gpu_nccl.err
tail -4 gpurun_out/mg_pytest.log; tail -2 gpurun_out/mg_pytest_nccl.log
for f in gpurun_out/mg_bench_2gpu.json gpurun_out/mg_bench_2gpu_nccl.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1]); r=d["roofline"]
    print(sys.argv[1], "value %.4g ms/step %.3f fast %.3f slow %.3f resolve %.3f sort+halo %.3f e2e %.3g"%(d["value"], d["ms_per_step"], r["ms_diffuse_fast"], r["ms_diffuse_slow"], r["ms_resolve"], r["ms_sort"], d["e2e"]["value"]))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
done
tail -4 gpurun_out/mg_bench_2gpu.err
