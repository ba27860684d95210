"""GroupNorm(+SiLU)+quantise launches at the SD UNet's shapes (B=16): engine-timed, or one launch each for ncu.
usage: prof_gn.py [--once]"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-diffusion_b200"))
from qdiff_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
B = 16
SHAPES = [(320, 4096), (640, 4096), (960, 4096), (640, 1024), (1920, 1024), (1280, 256), (2560, 256), (1280, 64)]
if "--shape" in sys.argv:          # --shape B C HW: one shape (e.g. CIFAR top level: 256 128 1024)
    i = sys.argv.index("--shape")
    B, SHAPES = int(sys.argv[i + 1]), [(int(sys.argv[i + 2]), int(sys.argv[i + 3]))]
L = _lib.lib()
for Cc, HW in SHAPES:
    x = torch.randn(B * HW, Cc, device=dev)
    gamma, beta = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    q = ops.act_qparams(0.03, 120, 8, False)
    out = torch.empty(B * HW, Cc, dtype=torch.uint8, device=dev)
    ws = torch.empty(ops.gn_workspace_floats(B, HW, Cc), device=dev)
    d = ops.groupnorm_desc(x, gamma, beta, ws, B=B, HW=HW, C_=Cc, ld_x=Cc, eps=1e-5, silu=True, outs=[(out, Cc, q)])
    if "--once" in sys.argv:
        ops.groupnorm_quant(d)
        torch.cuda.synchronize()
        continue
    # a producer stand-in (writes x) precedes every GroupNorm so that the L2 state resembles the UNet program's
    eng = C.c_void_p()
    _lib.check(L.qd_engine_create(0, C.byref(eng)), "create")
    REPS = 10
    for _ in range(REPS):
        _lib.check(L.qd_engine_add_op(eng, _lib.QD_OP_GROUPNORM, C.byref(d)), "add")
    _lib.check(L.qd_engine_finalize(eng), "finalize")
    _lib.check(L.qd_engine_run(eng, _lib.stream_ptr()), "run")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(L.qd_engine_run(eng, _lib.stream_ptr()), "run")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    nbytes = B * HW * Cc * 9
    print(f"groupnorm C={Cc:5d} HW={HW:5d}: {ms * 1e3:7.1f} us  {nbytes / ms / 1e6:7.0f} GB/s (2 reads + code write)", flush=True)
    L.qd_engine_destroy(eng)
