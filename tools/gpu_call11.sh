#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log; }
TMO=400 run t_cfg python -m pytest tests/test_unet_gpu.py -q -k "cfg_prefix or packed" -s
TMO=300 run t_attn python -m pytest tests/test_ops_gpu.py -q -k "attention"
TMO=300 run t_samplers python -m pytest tests/test_samplers_gpu.py -q
TMO=300 run bench_h python bench.py --steps 10 --warmup 3 --no-cpu-baseline
QDIFF_CFG_DEDUP=0 TMO=300 run bench_h_nodedup python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
