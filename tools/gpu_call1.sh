#!/bin/bash
# GPU call 1 of round 2: new parity tests, the existing suite, a bench line, per-op profile, smoke.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log; }
TMO=400 run t_insitu_golden python -m pytest tests/test_insitu_gpu.py -q -k golden_unets -s
TMO=400 run t_samplers python -m pytest tests/test_samplers_gpu.py -q -s
TMO=400 run t_insitu_full python -m pytest tests/test_insitu_gpu.py -q -k "fullsize and not sd_v1" -s
TMO=500 run t_rest python -m pytest tests -q -m gpu --ignore=tests/test_insitu_gpu.py --ignore=tests/test_samplers_gpu.py -k "not sd_v1-16" -s
TMO=300 run bench_a python bench.py --steps 10 --warmup 3
TMO=200 run op_profile python tools/op_profile.py sd_v1 16
TMO=200 run smoke python -c "import __graft_entry__ as g; g.smoke()"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
