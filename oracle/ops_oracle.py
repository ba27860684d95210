"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement, in plain torch/numpy float + integer arithmetic, of the reference's per-op
semantics on the UNet hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.

Each function cites the reference lines it restates (paths relative to the reference repo).
Pinned against the imported reference by tests/test_oracle_golden.py via the fixtures that
tools/make_golden.py generated from /root/reference (tests/golden/*.pt).
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- quantizers
def uaq_clamp_range(n_bits, symmetric):
    """qdiff/quant_layer.py:54,83-87 -- sym: n_lv = 2^(n-1)-1, clamp [-n_lv-1, n_lv]; asym [0, 2^n-1]."""
    if symmetric:
        n_lv = 2 ** (n_bits - 1) - 1
        return -n_lv - 1, n_lv
    return 0, 2 ** n_bits - 1


def uaq_codes(x, delta, zero_point, n_bits, symmetric):
    """Integer codes of UniformAffineQuantizer.forward (qdiff/quant_layer.py:82-87)."""
    lo, hi = uaq_clamp_range(n_bits, symmetric)
    x_int = torch.round(x / delta) + zero_point
    return torch.clamp(x_int, lo, hi)


def uaq_fake_quant(x, delta, zero_point, n_bits, symmetric):
    """qdiff/quant_layer.py:82-89: quantise then de-quantise."""
    return (uaq_codes(x, delta, zero_point, n_bits, symmetric) - zero_point) * delta


def uaq_init_max(x, n_bits, symmetric, always_zero=False):
    """'max' scale init for a per-tensor quantizer (qdiff/quant_layer.py:142-160)."""
    x_min = min(x.min().item(), 0)
    x_max = max(x.max().item(), 0)
    x_absmax = max(abs(x_min), x_max)
    if symmetric:
        n_levels = 2 ** (n_bits - 1) - 1
        delta = x_absmax / n_levels
    else:
        n_levels = 2 ** n_bits
        delta = float(x.max().item() - x.min().item()) / (n_levels - 1)
    if delta < 1e-8:
        delta = 1e-8
    zero_point = round(-x_min / delta) if not (symmetric or always_zero) else 0
    return torch.tensor(delta, dtype=torch.float32), zero_point


def weight_init_max(w, n_bits):
    """Channel-wise 'max' init (qdiff/quant_layer.py:114-136 looping :142-160 per out channel)."""
    deltas, zps = [], []
    for c in range(w.shape[0]):
        d, z = uaq_init_max(w[c], n_bits, False)
        deltas.append(d)
        zps.append(float(z))
    return torch.stack(deltas), torch.tensor(zps, dtype=torch.float32)


def adaround_hard_fake_quant(w, delta, zero_point, alpha, n_bits):
    """AdaRoundQuantizer.forward, hard branch (qdiff/adaptive_rounding.py:49-59)."""
    shape = (-1,) + (1,) * (w.dim() - 1)
    delta = delta.reshape(shape)
    zero_point = zero_point.reshape(shape)
    x_floor = torch.floor(w / delta)
    x_int = x_floor + (alpha >= 0).float()
    x_quant = torch.clamp(x_int + zero_point, 0, 2 ** n_bits - 1)
    return (x_quant - zero_point) * delta


def uaq_weight_fake_quant(w, delta, zero_point, n_bits):
    """Weight path before convert_adaround: rne (qdiff/quant_layer.py:82-88, channel-wise)."""
    shape = (-1,) + (1,) * (w.dim() - 1)
    delta = delta.reshape(shape)
    zero_point = zero_point.reshape(shape)
    x_quant = torch.clamp(torch.round(w / delta) + zero_point, 0, 2 ** n_bits - 1)
    return (x_quant - zero_point) * delta


# ----------------------------------------------------------------------------- integer GEMM / conv
def int_linear(a_codes, zx, ws, scale, bias=None):
    """Exact integer restatement of QuantModule.forward for Linear/1x1:
    y = scale[n] * sum_k (a-zx) * ws[n,k] + b.  a_codes [M,K], ws [N,K] (zero-point-free)."""
    acc = (a_codes.double() - zx) @ ws.double().t()
    y = acc * scale.double()[None, :]
    if bias is not None:
        y = y + bias.double()[None, :]
    return y


def int_conv3x3(a_codes_nchw, zx, ws4, scale, bias=None):
    """Same for a 3x3/pad-1 conv; padded positions are REAL zeros (code == zx)."""
    acc = F.conv2d(a_codes_nchw.double() - zx, ws4.double(), None, stride=1, padding=1)
    y = acc * scale.double()[None, :, None, None]
    if bias is not None:
        y = y + bias.double()[None, :, None, None]
    return y


# ----------------------------------------------------------------------------- elementwise
def silu(x):
    return x * torch.sigmoid(x)


def geglu(x2):
    """ldm/modules/attention.py:42-44."""
    x, gate = x2.chunk(2, dim=-1)
    return x * F.gelu(gate)


def timestep_embedding_ldm(t, dim, max_period=10000):
    """ldm/modules/diffusionmodules/util.py:151-171."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def timestep_embedding_ddim(t, dim):
    """ddim/models/diffusion.py:6-24."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float32) * -e)
    e = t.float()[:, None] * e[None, :]
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1)
    if dim % 2 == 1:
        e = F.pad(e, (0, 1, 0, 0))
    return e


# ----------------------------------------------------------------------------- attention
def attention_fake_quant(q, k, v, qp_q, qp_k, qp_v, qp_w, scale_after):
    """Shared core of the three attention flavours once q,k,v are [BH, T, d] float:
    quant(q), quant(k) -> QK^T (* scale_after) -> softmax -> quant(P), quant(v) -> PV.
    qp_* = (delta, zp, n_bits, symmetric).  cross_attn_forward qdiff/quant_block.py:200-219;
    QuantAttnBlock :366-380; QuantQKMatMul/QuantSMVMatMul :124-129,153-154 (pre-scaled q,k -> scale_after=1).
    """
    qq = uaq_fake_quant(q, *qp_q)
    kq = uaq_fake_quant(k, *qp_k)
    sim = torch.einsum('bid,bjd->bij', qq, kq) * scale_after
    p = sim.softmax(dim=-1)
    pq = uaq_fake_quant(p, *qp_w)
    vq = uaq_fake_quant(v, *qp_v)
    return torch.einsum('bij,bjd->bid', pq, vq)
