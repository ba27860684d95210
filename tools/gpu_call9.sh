#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log; }
TMO=900 run t_all python -m pytest tests -q -m gpu --deselect "tests/test_insitu_gpu.py::test_every_op_matches_oracle_fullsize[sd_v1-1]"
TMO=300 run bench_cifar python bench.py --workload cifar10 --steps 10 --warmup 3
TMO=400 run bench_bedroom python bench.py --workload lsun_bedroom --steps 10 --warmup 3
TMO=600 run ncu_launches ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-graph
TMO=900 run t_insitu_sd python -m pytest "tests/test_insitu_gpu.py::test_every_op_matches_oracle_fullsize[sd_v1-1]" -q -s
