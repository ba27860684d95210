#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log; }
TMO=400 run t_ops python -m pytest tests/test_ops_gpu.py -q
TMO=400 run t_unet python -m pytest tests/test_unet_gpu.py -q -k "golden or fresh" -s
TMO=400 run t_insitu_golden python -m pytest tests/test_insitu_gpu.py -q -k golden_unets -s
TMO=300 run bench_f python bench.py --steps 10 --warmup 3 --no-cpu-baseline
QDIFF_GN_STATS=0 TMO=300 run bench_f_nostats python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
TMO=200 run op_profile python tools/op_profile.py sd_v1 16
