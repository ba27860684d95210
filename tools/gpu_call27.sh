#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log; }
TMO=300 run t_ops python -m pytest tests/test_ops_gpu.py -q -x
TMO=400 run t_unet python -m pytest tests/test_unet_gpu.py -q -x -k "golden or cfg_prefix"
TMO=300 run bench_pdl python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
QDIFF_PDL=0 TMO=300 run bench_nopdl python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
TMO=300 run bench_church_pdl python bench.py --workload lsun_church --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
QDIFF_PDL=0 TMO=300 run bench_church_nopdl python bench.py --workload lsun_church --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
TMO=300 run bench_cifar_pdl python bench.py --workload cifar10 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
QDIFF_PDL=0 TMO=300 run bench_cifar_nopdl python bench.py --workload cifar10 --steps 10 --warmup 3 --no-cpu-baseline --no-roofline
for f in bench_pdl bench_nopdl bench_church_pdl bench_church_nopdl bench_cifar_pdl bench_cifar_nopdl; do echo -n "$f: "; grep '^{' gpurun_out/$f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"; done
