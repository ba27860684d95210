#!/bin/bash
# Round-end evidence run on one B200 (under gpurun): full GPU suite, the four BASELINE bench lines, the ncu launch list of
# one SD bench step (+ summary and GEMM traffic), decode timings.  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
run() { name=$1; shift; ( time timeout "$TMO" "$@" ) > gpurun_out/$name.log 2>&1; echo "rc=$?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log; }
TMO=900 run final_gputests python -m pytest tests -m gpu -x -q
# the ncu launch list first: bench.py reads roofline.traffic from the summary of THIS build's launches
TMO=600 run final_ncu_launches ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches_step.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --no-graph
python tools/launch_summary.py gpurun_out/final_launches_step.csv --traffic gpurun_out/final_roofline_traffic_sd_v1.json > gpurun_out/final_launches_step.summary.txt 2>&1
cp gpurun_out/final_roofline_traffic_sd_v1.json profiles/r02_roofline_traffic_sd_v1.json
head -24 gpurun_out/final_launches_step.summary.txt
for w in sd_v1 cifar10 lsun_bedroom lsun_church; do
  TMO=400 run final_bench_$w python bench.py --workload $w
  grep '^{' gpurun_out/final_bench_$w.log > gpurun_out/final_bench_$w.json
done
TMO=300 run final_bench_ref python bench.py --impl reference --steps 2 --warmup 1
for a in "sd_v1 1" "sd_v1 4" "lsun_bedroom 4"; do timeout 300 python tools/bench_decode.py $a 2>&1 | tail -1; done > gpurun_out/final_decode.txt
cat gpurun_out/final_decode.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
