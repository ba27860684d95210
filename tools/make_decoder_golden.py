"""Golden fixtures for the first-stage decode step, produced by RUNNING THE REFERENCE
(imported read-only from /root/reference) on tiny seeded autoencoders.  Build container only.

    python tools/make_decoder_golden.py

ldm/models/autoencoder.py imports two packages that are not installed here: pytorch_lightning (only as the base class
`pl.LightningModule`; stubbed with torch.nn.Module) and taming (VectorQuantizer2; stubbed with the restatement of its
published algorithm from oracle/decoder_oracle.py - so the VQ nearest-neighbour step itself is NOT pinned by these
fixtures, everything after it is).  With the stubs the reference's UNMODIFIED AutoencoderKL.decode and
VQModelInterface.decode run, on top of the reference's own Decoder class.
"""
import os
import sys
import tempfile
import types

import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def _import_reference():
    stub = tempfile.mkdtemp()
    os.makedirs(os.path.join(stub, "omegaconf"))
    open(os.path.join(stub, "omegaconf", "__init__.py"), "w").write("")
    open(os.path.join(stub, "omegaconf", "listconfig.py"), "w").write("class ListConfig(list):\n    pass\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = torch.nn.Module
    sys.modules["pytorch_lightning"] = pl
    from oracle import decoder_oracle as D

    class VectorQuantizer2(torch.nn.Module):
        """Stand-in for taming's class (absent): same constructor and forward signature, algorithm restated in the oracle."""

        def __init__(self, n_e, e_dim, beta, remap=None, unknown_index="random", sane_index_shape=False, legacy=True):
            super().__init__()
            assert remap is None
            self.n_e, self.e_dim, self.beta = n_e, e_dim, beta
            self.embedding = torch.nn.Embedding(n_e, e_dim)
            self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)

        def forward(self, z, temp=None, rescale_logits=False, return_logits=False):
            z_q, idx = D.vq_quantize({"quantize.embedding.weight": self.embedding.weight.detach()}, z, return_indices=True)
            return z_q, torch.zeros(()), (None, None, idx)

    for name in ("taming", "taming.modules", "taming.modules.vqvae", "taming.modules.vqvae.quantize"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["taming.modules.vqvae.quantize"].VectorQuantizer2 = VectorQuantizer2


def _sd(model):
    return {k: v.detach().clone() for k, v in model.state_dict().items()
            if k.startswith(("decoder.", "post_quant_conv.", "quantize."))}


def main():
    _import_reference()
    from ldm.models.autoencoder import AutoencoderKL, VQModelInterface
    from ldm.modules.diffusionmodules.model import Decoder
    loss = dict(target="torch.nn.Identity")
    with torch.no_grad():
        # ---- KL-f4-style toy (the SD / church first stage is AutoencoderKL; configs/stable-diffusion/v1-inference.yaml:46-67)
        torch.manual_seed(0)
        dd = dict(double_z=True, z_channels=4, resolution=32, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 2],
                  num_res_blocks=1, attn_resolutions=[], dropout=0.0)
        m = AutoencoderKL(ddconfig=dd, lossconfig=loss, embed_dim=4).eval()
        z = torch.randn(2, 4, 8, 8) * 3.0
        sf = 0.18215
        out = m.decode(1. / sf * z)                                  # ddpm.py:731 + autoencoder.py:330-333
        dec_only = Decoder(**dd)
        dec_only.load_state_dict(m.decoder.state_dict())
        z_dec = torch.randn(2, 4, 8, 8)
        torch.save(dict(kind="kl", ddconfig=dd, embed_dim=4, scale_factor=sf, sd=_sd(m), z=z, out=out,
                        z_dec=z_dec, out_dec=dec_only(z_dec)), os.path.join(OUT, "decoder_kl_tiny.pt"))
        print("decoder_kl_tiny", tuple(out.shape), float(out.abs().mean()))
        # ---- VQ-f2-style toy with attention at the 8x8 level (the LSUN-bedroom first stage is VQModelInterface;
        # configs/latent-diffusion/lsun_bedrooms-ldm-vq-4.yaml)
        torch.manual_seed(1)
        dd = dict(double_z=False, z_channels=3, resolution=16, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2],
                  num_res_blocks=1, attn_resolutions=[8], dropout=0.0)
        m = VQModelInterface(embed_dim=3, ddconfig=dd, lossconfig=loss, n_embed=64).eval()
        m.quantize.embedding.weight.data.normal_(0, 1.0)
        z = torch.randn(2, 3, 8, 8)
        out = m.decode(z)
        out_nq = m.decode(z, force_not_quantize=True)
        _, _, (_, _, idx) = m.quantize(z)
        torch.save(dict(kind="vq", ddconfig=dd, embed_dim=3, n_embed=64, scale_factor=1.0, sd=_sd(m), z=z, out=out,
                        out_not_quantized=out_nq, indices=idx), os.path.join(OUT, "decoder_vq_tiny.pt"))
        print("decoder_vq_tiny", tuple(out.shape), float(out.abs().mean()))


if __name__ == "__main__":
    main()
