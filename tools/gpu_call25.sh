#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
V=$PWD/q-diffusion_b200/csrc/experimental/variants
for v in p0 p3 p2 p4 p3_q2 p3_q1; do
  echo -n "$v: "; QDIFF_B200_LIB=$V/lib_$v.so timeout 120 python tools/prof_attn.py 2>&1 | tail -1
  echo -n "$v int: "; ATTN_F16=0 QDIFF_B200_LIB=$V/lib_$v.so timeout 120 python tools/prof_attn.py 2>&1 | tail -1
done > gpurun_out/attn_poly.log 2>&1
for v in p3 p3_q1; do QDIFF_B200_LIB=$V/lib_$v.so timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "attention" 2>&1 | tail -2 >> gpurun_out/attn_poly.log; done
cat gpurun_out/attn_poly.log
