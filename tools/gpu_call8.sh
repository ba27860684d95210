#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log; }
TMO=400 run t_new python -m pytest tests/test_unet_gpu.py tests/test_samplers_gpu.py -q -k "packed or dpm" -s
TMO=900 run t_all python -m pytest tests -q -m gpu -x --deselect "tests/test_insitu_gpu.py::test_every_op_matches_oracle_fullsize[sd_v1-1]"
TMO=200 run smoke python -c "import __graft_entry__ as g; g.smoke()"
TMO=300 run bench_cifar python bench.py --workload cifar10 --steps 10 --warmup 3
TMO=400 run bench_bedroom python bench.py --workload lsun_bedroom --steps 10 --warmup 3
TMO=300 run bench_church python bench.py --workload lsun_church --steps 10 --warmup 3
TMO=300 run bench_sd python bench.py --steps 20 --warmup 5
TMO=300 run bench_ref python bench.py --impl reference --steps 3 --warmup 1
