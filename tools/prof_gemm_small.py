"""Launch the small-K GEMM variants of the SD transformer blocks once each (for `ncu --set full`), or time them
with CUDA events (`--time`): to_q (requantised, padded per-head layout), to_v (transposed V^T codes), ff.net.0
(GEGLU fused), ff.net.2 / proj_out (fp32 + residual), proj_in (requantised plain)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-diffusion_b200"))
from qdiff_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M, T = 16 * 4096, 4096


def case(name):
    g = torch.Generator(device=dev).manual_seed(1)
    rnd = lambda *s: torch.rand(*s, device=dev, generator=g)
    oq = ops.qparams(0.05, 128, 0, 255)
    if name == "to_q":
        C_, N = 320, 320
        out_q = torch.empty(M, 8 * 64, dtype=torch.uint8, device=dev)
        kw = dict(out_q=out_q, ldq=512, oq=oq, out_q_head=(40, 64))
    elif name == "to_v":
        C_, N = 320, 320
        out_q = torch.empty(16, 320, T + 16, dtype=torch.uint8, device=dev)
        kw = dict(out_q=out_q, ldq=T + 16, oq=oq, out_q_transposed=True, rows_per_batch=T)
    elif name == "geglu":
        C_, N = 320, 2560
        out_q = torch.empty(M, 1280, dtype=torch.uint8, device=dev)
        kw = dict(out_q=out_q, ldq=1280, oq=oq, geglu=True)
    elif name == "ff_out":
        C_, N = 1280, 320
        out = torch.randn(M, 320, device=dev)
        kw = dict(out=out, ldo=320, residual=out, ldr=320)
    elif name == "proj_in":
        C_, N = 320, 320
        out = torch.empty(M, 320, device=dev)
        kw = dict(out=out, ldo=320)
    else:
        raise KeyError(name)
    a = torch.randint(0, 256, (M, C_), dtype=torch.uint8, device=dev)
    w = torch.randint(-7, 8, (N, C_), dtype=torch.int8, device=dev)
    scale = rnd(N) * 1e-3
    bias = rnd(N)
    corr = torch.randint(-1000, 1000, (N,), dtype=torch.int32, device=dev)
    d = ops.gemm_desc(a, w, scale, M=M, N=N, C=C_, a_signed=False, bias=bias, corr=corr, **kw)
    d._keep = (a, w, scale, bias, corr, kw)
    return d, 2.0 * M * N * C_


names = ["to_q", "to_v", "geglu", "ff_out", "proj_in"]
if "--time" in sys.argv:
    for n in names:
        d, ops_ = case(n)
        for _ in range(3):
            ops.qgemm(d)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.qgemm(d)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{n:8s} {ms * 1e3:8.1f} us  {ops_ / ms / 1e9:8.1f} TOPS", flush=True)
else:
    for n in names:
        d, _ = case(n)
        ops.qgemm(d)
        torch.cuda.synchronize()
        print(n, "done")
