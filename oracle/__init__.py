"""ORACLE package: CPU restatement of the reference's hot path. TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never from the product package (q-diffusion_b200/), which must fail loudly without its
CUDA library instead of falling back to anything in here.
"""
