#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
V=$PWD/q-diffusion_b200/csrc/experimental/variants
for v in w32 w16; do
  for T in 4096 1024; do echo -n "$v: "; QDIFF_B200_LIB=$V/lib_$v.so timeout 120 python tools/prof_attn.py $T 2>&1 | tail -1; done
  QDIFF_B200_LIB=$V/lib_$v.so timeout 300 python -m pytest tests/test_ops_gpu.py -q -k "attention" 2>&1 | tail -2
done > gpurun_out/attn_pf.log 2>&1
cat gpurun_out/attn_pf.log
