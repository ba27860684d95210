#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log; }
TMO=400 run t_cfg python -m pytest tests/test_unet_gpu.py -q -k "cfg_prefix" -s
TMO=600 run t_w8 python -m pytest tests/test_unet_gpu.py tests/test_insitu_gpu.py -q -k "w8a8 or church"
TMO=300 run bench_h python bench.py --steps 10 --warmup 3 --no-cpu-baseline
TMO=300 run bench_church python bench.py --workload lsun_church --steps 10 --warmup 3 --no-cpu-baseline
QDIFF_W8_KDUP=0 TMO=300 run bench_church_2gemm python bench.py --workload lsun_church --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
