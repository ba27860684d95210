#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log; }
TMO=400 run t_ops python -m pytest tests/test_ops_gpu.py -q
TMO=400 run t_unet python -m pytest tests/test_unet_gpu.py -q -k "golden or fresh"
TMO=400 run t_insitu_golden python -m pytest tests/test_insitu_gpu.py -q -k golden_unets
TMO=100 run attn_2cta python tools/prof_attn.py 4096 4096
QDIFF_ATTN_2CTA=0 TMO=100 run attn_1cta python tools/prof_attn.py 4096 4096
TMO=100 run attn_2cta_1024 python tools/prof_attn.py 1024 1024
TMO=200 run gemm_small_time python tools/prof_gemm_small.py --time
TMO=300 run sweep_bn python tools/sweep_bn.py
TMO=300 run bench_c python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TMO=200 run op_profile python tools/op_profile.py sd_v1 16
TMO=300 run ncu_attn ncu --set full --clock-control none --import-source on -k regex:qattention_tc -c 1 -f -o gpurun_out/ncu_attn python tools/prof_attn.py 4096 4096
