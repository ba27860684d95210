#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log; }
TMO=400 run t_ops python -m pytest tests/test_ops_gpu.py -q -x
TMO=400 run t_unet python -m pytest tests/test_unet_gpu.py -q -k "golden or fresh"
TMO=100 run attn_time python tools/prof_attn.py 4096 4096
TMO=100 run attn_time_1024 python tools/prof_attn.py 1024 1024
TMO=200 run gemm_small_time python tools/prof_gemm_small.py --time
TMO=200 run gn_time python tools/prof_gn.py
TMO=300 run sweep_bn python tools/sweep_bn.py
TMO=300 run bench_b python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TMO=200 run op_profile python tools/op_profile.py sd_v1 16
