"""Launch a few representative GEMM/conv shapes once each (for `ncu --set full`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-diffusion_b200"))
from qdiff_b200 import ops  # noqa: E402

SHAPES = [("conv 960->320 @64", 16, 64, 64, 960, 320, 9), ("linear 320->320 @4096", 16, 64, 64, 320, 320, 1),
          ("conv 1280->1280 @16", 16, 16, 16, 1280, 1280, 9)]
dev = torch.device("cuda:0")
for name, B, H, W, C, N, taps in SHAPES:
    M = B * H * W
    a = torch.randint(0, 256, (M, C), dtype=torch.uint8, device=dev)
    w = torch.randint(-7, 8, (N, taps * C), dtype=torch.int8, device=dev)
    scale = torch.rand(N, device=dev) * 1e-3
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev)
    d = ops.gemm_desc(a, w, scale, M=M, N=N, C=C, taps=taps, conv_bhw=(B, H, W) if taps == 9 else None,
                      a_signed=False, bias=bias, out=out, ldo=N)
    for _ in range(3):
        ops.qgemm(d)
    torch.cuda.synchronize()
    print(name, "done")
