#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
V=$PWD/q-diffusion_b200/csrc/experimental/variants
for v in pb3 pb2 pb3_noxu pb3_p1 pb3_p2 pb3_p1_noxu pb3_p2_noxu; do
  echo -n "$v: "; QDIFF_B200_LIB=$V/lib_$v.so timeout 120 python tools/prof_attn.py 2>&1 | tail -1
  echo -n "$v int: "; ATTN_F16=0 QDIFF_B200_LIB=$V/lib_$v.so timeout 120 python tools/prof_attn.py 2>&1 | tail -1
done > gpurun_out/attn_pb.log 2>&1
QDIFF_B200_LIB=$V/lib_pb3.so timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "attention" 2>&1 | tail -2 >> gpurun_out/attn_pb.log
cat gpurun_out/attn_pb.log
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:qattention_tc -c 1 -f -o gpurun_out/ncu_attn_f16 python tools/prof_attn.py 4096 4096 ) > gpurun_out/ncu_attn_f16.log 2>&1
ls -la gpurun_out/*.ncu-rep
