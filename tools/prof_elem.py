"""Representative memory-bound launches (GroupNorm+SiLU+quant, LayerNorm+quant, GEGLU quant) for ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-diffusion_b200"))
from qdiff_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, HW, C = 16, 4096, 960
x = torch.randn(B * HW, C, device=dev)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
q = ops.act_qparams(0.03, 120, 8, False)
out = torch.empty(B * HW, C, dtype=torch.uint8, device=dev)
ws = torch.empty(ops.gn_workspace_floats(B, HW, C), device=dev)
d = ops.groupnorm_desc(x, gamma, beta, ws, B=B, HW=HW, C_=C, ld_x=C, eps=1e-5, silu=True, outs=[(out, C, q)])


def timed(fn, name, nbytes):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name}: {ms:.3f} ms, {nbytes / ms / 1e6:.0f} GB/s algorithmic")


timed(lambda: ops.groupnorm_quant(d), "groupnorm+silu+quant C=960 @64x64 B=16", B * HW * C * 9)
C2 = 320
x2 = torch.randn(B * HW, C2, device=dev)
outs = [(torch.empty(B * HW, C2, dtype=torch.uint8, device=dev), C2, q) for _ in range(3)]
dl = ops.layernorm_desc(x2, torch.ones(C2, device=dev), torch.zeros(C2, device=dev), M=B * HW, C_=C2, ld_x=C2, eps=1e-5, outs=outs)
timed(lambda: ops.layernorm_quant(dl), "layernorm+3 quant C=320 65536 tokens", B * HW * C2 * 7)
x3 = torch.randn(B * HW, 2560, device=dev)
o3 = torch.empty(B * HW, 1280, dtype=torch.uint8, device=dev)
dq = ops.quantize_desc(x3, o3, M=B * HW, C_=1280, ld_src=2560, ld_dst=1280, q0=q, act=2)
timed(lambda: ops.quantize(dq), "geglu+quant 65536 x 1280", B * HW * (2560 * 4 + 1280))
