#!/bin/bash
# One parameterised GPU job (replaces the per-letter gpu_job_*.sh scripts of round 1).
#   usage: tools/gpu_job.sh TAG stage [stage ...]
# Output goes to gpurun_out/TAG_*.  Stages:
#   pytest      the whole GPU tier                       smoke      __graft_entry__.smoke()
#   bench       default bench.py (1e8, cpu_baseline)     benchq     bench.py --no-cpu (no CPU leg)
#   reference   bench.py --impl reference                configs    bench.py --config 1..4
#   launches    ncu launch list of a short bench         ncu        ncu --set full of a steady-state iteration
#   ab:LIBS     A/B of library variants (comma-separated names under mcell_b200/libmcx_<name>.so) on bench --no-cpu
#   mg:N        torchrun bench with N ranks              mgtest:N   multi-GPU pytest tier
set -x
TAG=$1; shift
mkdir -p gpurun_out
O=gpurun_out/$TAG
for stage in "$@"; do
  case $stage in
    pytest)    ( time timeout 1500 python -m pytest tests -m gpu -x -q ) > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log; tail -5 ${O}_pytest.log ;;
    smoke)     ( time timeout 300 python __graft_entry__.py --smoke ) > ${O}_smoke.log 2>&1; echo "smoke rc=$?" >> ${O}_smoke.log; tail -3 ${O}_smoke.log ;;
    bench)     ( time timeout 900 python bench.py ) > ${O}_bench_1e8.json 2> ${O}_bench_1e8.err; cut -c1-1500 ${O}_bench_1e8.json; tail -3 ${O}_bench_1e8.err ;;
    benchq)    ( time timeout 600 python bench.py --no-cpu ) > ${O}_benchq_1e8.json 2> ${O}_benchq_1e8.err; cut -c1-1500 ${O}_benchq_1e8.json; tail -3 ${O}_benchq_1e8.err ;;
    reference) ( time timeout 900 python bench.py --impl reference --steps 10 --warmup 3 ) > ${O}_bench_reference.json 2> ${O}_bench_reference.err; cut -c1-400 ${O}_bench_reference.json ;;
    configs)   for c in 1 2 3 4; do timeout 600 python bench.py --config $c --no-cpu --steps 20 --warmup 5 --e2e-calls 1 >> ${O}_configs.jsonl 2>> ${O}_configs.err; done; cut -c1-600 ${O}_configs.jsonl ;;
    launches)  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file ${O}_launches.csv \
                 python bench.py --steps 2 --warmup 3 --e2e-calls 1 --no-cpu > ${O}_launches.log 2>&1 ;;
    ncu)       # skip the release/upload launches and three warm-up iterations, capture one steady-state iteration
               timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-k_diffuse|k_scatter|k_round|k_resolve|k_compact|k_probe}" \
                 -s ${NCU_SKIP:-60} -c ${NCU_COUNT:-16} -o ${O}_prof -f python bench.py --steps 3 --warmup 3 --e2e-calls 1 --no-cpu > ${O}_ncu.log 2>&1; tail -3 ${O}_ncu.log ;;
    ab:*)      for v in $(echo ${stage#ab:} | tr ',' ' '); do
                 lib=mcell_b200/libmcx_$v.so; [ "$v" = head ] && lib=mcell_b200/libmcx.so
                 extra_env=""; case $v in env_*) lib=mcell_b200/libmcx.so; extra_env="$(echo ${v#env_} | tr '+' ' ')";; esac   # env_VAR=1+VAR2=x: HEAD with tuning variables
                 for rep in 1 2; do env $extra_env MCX_LIB=$PWD/$lib timeout 600 python bench.py --no-cpu --e2e-calls 1 ${AB_ARGS} > ${O}_ab_${v}_$rep.json 2>> ${O}_ab.err
                   python - <<EOF
import json; d=json.load(open("${O}_ab_${v}_$rep.json")); r=d["roofline"]
print("AB $v rep $rep: ms/step %.3f fast %.3f slow %.3f resolve %.3f sort %.3f value %.4g" % (d["ms_per_step"], r["ms_diffuse_fast"], r["ms_diffuse_slow"], r["ms_resolve"], r["ms_sort"], d["value"]))
EOF
                 done; done ;;
    mg:*)      n=${stage#mg:}; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
                 bench.py --gpus $n --steps 10 --warmup 3 > ${O}_bench_${n}gpu.json 2> ${O}_bench_${n}gpu.err; cut -c1-1500 ${O}_bench_${n}gpu.json; tail -3 ${O}_bench_${n}gpu.err ;;
    mgtest:*)  ( time timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > ${O}_mgtest.log 2>&1; echo "rc=$?" >> ${O}_mgtest.log; tail -5 ${O}_mgtest.log ;;
    *)         echo "unknown stage $stage" ;;
  esac
done
