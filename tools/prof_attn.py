"""One SD self-attention launch at the 64x64 level (B=16, 8 heads, d=40, 4096 tokens, asymmetric,
sm_abit 16) for `ncu --set full`, plus a CUDA-event timing."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "q-diffusion_b200"))
from qdiff_b200 import ops  # noqa: E402
from qdiff_b200._lib import AttentionDesc, ptr  # noqa: E402

dev = torch.device("cuda:0")
B, heads, d, T = 16, 8, 40, 4096
P = 64
q = torch.zeros(B, T, heads, P, dtype=torch.uint8, device=dev)
q[..., :d] = torch.randint(0, 256, (B, T, heads, d), dtype=torch.uint8, device=dev)
k = torch.zeros(B, T, heads, P, dtype=torch.uint8, device=dev)
k[..., :d] = torch.randint(0, 256, (B, T, heads, d), dtype=torch.uint8, device=dev)
vt = torch.randint(0, 256, (B, heads * d, T), dtype=torch.uint8, device=dev)
out = torch.empty(B, T, heads * d, device=dev)
a = AttentionDesc()
a.q, a.k, a.vt = ptr(q), ptr(k), ptr(vt)
a.ld_q = a.ld_k = heads * P
a.ld_vt, a.v_batch_stride = T, heads * d * T
a.B, a.heads, a.d, a.Tq, a.Tk = B, heads, d, T, T
a.head_stride_q = a.head_stride_k = P
a.head_stride_v = d
a.zq, a.zk, a.zv, a.zw = 120, 131, 127, 0
a.p_qmin, a.p_qmax, a.sm_bits = 0, 65535, 16
a.sim_scale = 0.04 * 0.04 * d ** -0.5 * 0.05
a.delta_w = 1.0 / 65535
a.out_scale = a.delta_w * 0.03
a.out, a.ld_out = ptr(out), heads * d
ws = torch.zeros(B * heads * T, dtype=torch.int32, device=dev)
a.ws = ptr(ws)
for _ in range(2):
    ops.attention(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.attention(a)
e1.record()
torch.cuda.synchronize()
scores = B * heads * T * T
print(f"self-attention 64x64: {e0.elapsed_time(e1):.3f} ms, {scores / e0.elapsed_time(e1) / 1e6:.1f} Gscore/s")
