#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log; }
TMO=400 run t_insitu_golden python -m pytest tests/test_insitu_gpu.py -q -k golden_unets -s
TMO=400 run t_insitu_full python -m pytest tests/test_insitu_gpu.py -q -k "fullsize and not sd_v1" -s
TMO=300 run sweep_bn python tools/sweep_bn.py
TMO=200 run gemm_small_time python tools/prof_gemm_small.py --time
TMO=200 run gn_time python tools/prof_gn.py
TMO=300 run ncu_gemm_small ncu --set full --clock-control none --import-source on -k regex:gemm_i8 -c 5 -f -o gpurun_out/ncu_gemm_small python tools/prof_gemm_small.py
TMO=300 run ncu_gn ncu --set full --clock-control none --import-source on -k regex:gn_apply -c 2 -f -o gpurun_out/ncu_gn python tools/prof_gn.py --once
ls -la gpurun_out/*.ncu-rep
