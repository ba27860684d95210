#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log; }
TMO=300 run t_attn python -m pytest tests/test_ops_gpu.py -q -k "attention"
TMO=100 run attn_g2 python tools/prof_attn.py 4096 4096
QDIFF_ATTN_GROUPS=1 TMO=100 run attn_g1 python tools/prof_attn.py 4096 4096
TMO=100 run attn_g2_1024 python tools/prof_attn.py 1024 1024
QDIFF_ATTN_2CTA=0 TMO=100 run attn_g2_1024_1cta python tools/prof_attn.py 1024 1024
TMO=400 run t_unet python -m pytest tests/test_unet_gpu.py tests/test_insitu_gpu.py -q -k "golden"
TMO=300 run bench_g python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TMO=300 run ncu_attn ncu --set full --clock-control none --import-source on -k regex:qattention_tc -c 1 -f -o gpurun_out/ncu_attn_g2 python tools/prof_attn.py 4096 4096
