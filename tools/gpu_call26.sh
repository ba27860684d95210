#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log; }
TMO=1500 run t_all python -m pytest tests -m gpu -q -x
TMO=200 run smoke python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')"
TMO=300 run bench_h python bench.py --steps 10 --warmup 3 --no-cpu-baseline
