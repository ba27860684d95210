#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log; }
TMO=300 run t_attn python -m pytest tests/test_ops_gpu.py -q -k "attention or out_q_f16"
{ timeout 120 python tools/prof_attn.py 2>&1 | tail -1; ATTN_F16=0 timeout 120 python tools/prof_attn.py 2>&1 | tail -1; timeout 120 python tools/prof_attn.py 1024 2>&1 | tail -1; } > gpurun_out/attn_time.log 2>&1; cat gpurun_out/attn_time.log
TMO=600 run t_insitu python -m pytest tests/test_insitu_gpu.py -q -x
TMO=300 run bench_h python bench.py --steps 10 --warmup 3 --no-cpu-baseline
