"""ORACLE (test infrastructure only -- see oracle/__init__.py).

Functional CPU restatement (torch fp32 / fp64, NCHW) of the first-stage DECODE step that follows the sampling loop
(SURVEY section 8 row f2), driven directly by a first-stage state dict (keys `decoder.*`, `post_quant_conv.*`,
`quantize.embedding.weight` - the names the reference's checkpoints use):

  decoder_forward        Decoder.forward            ldm/modules/diffusionmodules/model.py:538-571
    resnet_block         ResnetBlock.forward        model.py:122-144 (temb is None in the decoder: no temb_proj)
    attn_block           AttnBlock.forward          model.py:179-205
    upsample             Upsample.forward           model.py:57-61 (nearest 2x, then conv3x3)
  kl_decode              AutoencoderKL.decode       ldm/models/autoencoder.py:330-333
  vq_decode              VQModelInterface.decode    ldm/models/autoencoder.py:274-283
  vq_quantize            VectorQuantizer2.forward   taming-transformers (taming/modules/vqvae/quantize.py, `legacy` branch,
                         no remap), a third-party dependency the reference installs from git master
                         (environment.yaml: `-e git+https://github.com/CompVis/taming-transformers.git@master`) and that
                         is ABSENT from /root/reference and from this image.  Its published algorithm is restated here:
                         z -> 'b h w c' -> flat; d = sum(z^2,1,keepdim) + sum(e^2,1) - 2 * z @ e^T; index = argmin(d,1);
                         z_q = e[index]; straight-through z + (z_q - z); back to 'b c h w'.  PARITY UNPINNED for this
                         one function (no reference-side output can be produced here); everything else in this file is
                         pinned by tests/golden/decoder_*.pt (tools/make_decoder_golden.py runs the reference's own
                         Decoder class and, through import stubs for the absent pytorch_lightning / taming packages, the
                         unmodified AutoencoderKL.decode and VQModelInterface.decode).
  decode_first_stage     LatentDiffusion.decode_first_stage  ldm/models/diffusion/ddpm.py:710-767 (the plain branch:
                         z = 1/scale_factor * z, then first_stage_model.decode(z))
"""
import torch
import torch.nn.functional as F


def _gn(x, sd, key, eps=1e-6, groups=32):
    return F.group_norm(x, groups, sd[key + ".weight"].to(x.dtype), sd[key + ".bias"].to(x.dtype), eps)


def _conv(x, sd, key, padding=0):
    return F.conv2d(x, sd[key + ".weight"].to(x.dtype), sd[key + ".bias"].to(x.dtype), padding=padding)


def _swish(x):
    return x * torch.sigmoid(x)            # nonlinearity(), model.py:36-38


def resnet_block(x, sd, key):
    """model.py:122-144 with temb None and dropout in eval mode."""
    h = _conv(_swish(_gn(x, sd, key + ".norm1")), sd, key + ".conv1", 1)
    h = _conv(_swish(_gn(h, sd, key + ".norm2")), sd, key + ".conv2", 1)
    if (key + ".nin_shortcut.weight") in sd:
        x = _conv(x, sd, key + ".nin_shortcut")
    elif (key + ".conv_shortcut.weight") in sd:
        x = _conv(x, sd, key + ".conv_shortcut", 1)
    return x + h


def attn_block(x, sd, key):
    """model.py:179-205: single head, d = C, scores scaled by C^-1/2 after the product."""
    h = _gn(x, sd, key + ".norm")
    q, k, v = _conv(h, sd, key + ".q"), _conv(h, sd, key + ".k"), _conv(h, sd, key + ".v")
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(h, sd, key + ".proj_out")


def decoder_forward(sd, z, prefix="decoder.", trace=None):
    """Decoder.forward (model.py:538-571).  The architecture is read off the state dict: number of levels, blocks per
    level, presence of attention / upsampling convs."""
    p = prefix
    nlev = 1 + max(int(k[len(p) + 3:].split(".")[0]) for k in sd if k.startswith(p + "up."))
    h = _conv(z, sd, p + "conv_in", 1)
    h = resnet_block(h, sd, p + "mid.block_1")
    h = attn_block(h, sd, p + "mid.attn_1")
    h = resnet_block(h, sd, p + "mid.block_2")
    if trace is not None:
        trace["mid"] = h
    for lv in reversed(range(nlev)):
        ib = 0
        while f"{p}up.{lv}.block.{ib}.norm1.weight" in sd:
            h = resnet_block(h, sd, f"{p}up.{lv}.block.{ib}")
            if f"{p}up.{lv}.attn.{ib}.norm.weight" in sd:
                h = attn_block(h, sd, f"{p}up.{lv}.attn.{ib}")
            ib += 1
        if lv != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            if f"{p}up.{lv}.upsample.conv.weight" in sd:
                h = _conv(h, sd, f"{p}up.{lv}.upsample.conv", 1)
        if trace is not None:
            trace[f"up.{lv}"] = h
    h = _swish(_gn(h, sd, p + "norm_out"))
    return _conv(h, sd, p + "conv_out", 1)


def kl_decode(sd, z, trace=None):
    """AutoencoderKL.decode (autoencoder.py:330-333)."""
    return decoder_forward(sd, _conv(z, sd, "post_quant_conv"), trace=trace)


def vq_quantize(sd, z, return_indices=False):
    """VectorQuantizer2.forward (taming, legacy branch, no remap): nearest codebook entry per latent pixel."""
    e = sd["quantize.embedding.weight"].to(z.dtype)
    zp = z.permute(0, 2, 3, 1).contiguous()
    flat = zp.view(-1, e.shape[1])
    d = torch.sum(flat ** 2, dim=1, keepdim=True) + torch.sum(e ** 2, dim=1) - 2 * torch.einsum('bd,dn->bn', flat, e.t())
    idx = torch.argmin(d, dim=1)
    z_q = e[idx].view(zp.shape)
    z_q = zp + (z_q - zp)                  # straight-through estimator: kept, it rounds
    z_q = z_q.permute(0, 3, 1, 2).contiguous()
    return (z_q, idx) if return_indices else z_q


def vq_decode(sd, z, force_not_quantize=False, trace=None):
    """VQModelInterface.decode (autoencoder.py:274-283)."""
    quant = z if force_not_quantize else vq_quantize(sd, z)
    return decoder_forward(sd, _conv(quant, sd, "post_quant_conv"), trace=trace)


def decode_first_stage(sd, z, scale_factor=1.0, kind="kl", force_not_quantize=False):
    """LatentDiffusion.decode_first_stage (ddpm.py:710-767), plain branch."""
    z = 1. / scale_factor * z
    return kl_decode(sd, z) if kind == "kl" else vq_decode(sd, z, force_not_quantize)
