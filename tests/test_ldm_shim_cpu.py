"""The LatentDiffusion surface (qdiff_b200/ldm_shim.py) is what the reference's sampler classes need: when
/root/reference is importable (build container) its UNMODIFIED PLMSSampler and DDIMSampler are run on top of the shim
around a toy eps-model and must reproduce the committed reference results (tests/golden/samplers.pt); everywhere, the
shim's apply_model / DiffusionWrapper dispatch is checked against ddpm.py:895-905,1426-1445 semantics."""
import os
import sys

import pytest
import torch

from qdiff_b200.ldm_shim import LatentDiffusionShim
from tools.make_sampler_golden import toy_eps

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class ToyUNet:
    """Stands in for qdiff_b200.QuantModel: same call signature (x, timesteps, context=None)."""

    def __init__(self):
        self.calls = []

    def __call__(self, x, timesteps=None, context=None):
        self.calls.append(context)
        return toy_eps(x, timesteps, context)


def test_apply_model_dispatch():
    u = ToyUNet()
    m = LatentDiffusionShim(u, "crossattn", 1000, 0.00085, 0.012, device="cpu")
    x, t, c = torch.randn(2, 4, 8, 8), torch.tensor([5, 7]), torch.randn(2, 5, 16)
    ref = toy_eps(x, t, c)
    assert torch.equal(m.apply_model(x, t, c), ref)                       # tensor -> {'c_crossattn': [c]}
    assert torch.equal(m.apply_model(x, t, [c]), ref)
    assert torch.equal(m.apply_model(x, t, {"c_crossattn": [c]}), ref)
    assert u.calls[0] is c                                                # single context: the caller's tensor object is kept
    two = m.apply_model(x, t, {"c_crossattn": [c[:, :2], c[:, 2:]]})      # several contexts are concatenated on dim 1
    assert torch.allclose(two, ref)
    m0 = LatentDiffusionShim(ToyUNet(), None, device="cpu")
    assert torch.equal(m0.apply_model(x, t, None), toy_eps(x, t))
    assert m.model.diffusion_model is u and m.num_timesteps == 1000
    assert m.alphas_cumprod.dtype == torch.float32 and m.alphas_cumprod_prev[0] == 1.0


def test_reference_samplers_run_unchanged_on_the_shim():
    if not os.path.isdir("/root/reference/ldm"):
        pytest.skip("reference sources not present on this machine")
    from tools.make_golden import _import_reference
    _import_reference()
    try:
        from ldm.models.diffusion import ddim as ref_ddim
        from ldm.models.diffusion.plms import PLMSSampler
    except Exception as e:           # pragma: no cover
        pytest.skip(f"reference not importable here: {e}")
    g = torch.load(os.path.join(GOLD, "samplers.pt"), map_location="cpu", weights_only=False)

    class CpuPLMS(PLMSSampler):       # the reference's register_buffer moves everything to "cuda" (plms.py:19-23)
        def register_buffer(self, name, attr):
            setattr(self, name, attr)

    class CpuDDIM(ref_ddim.DDIMSampler):
        def register_buffer(self, name, attr):
            setattr(self, name, attr)

    p = g["plms"]
    shim = LatentDiffusionShim(ToyUNet(), "crossattn", 1000, p["linear_start"], p["linear_end"], device="cpu")
    with torch.no_grad():
        out, _ = CpuPLMS(shim).sample(S=p["S"], batch_size=p["x_T"].shape[0], shape=tuple(p["x_T"].shape[1:]),
                                      conditioning=p["cond"], verbose=False, unconditional_guidance_scale=p["scale"],
                                      unconditional_conditioning=p["uc"], eta=0.0, x_T=p["x_T"])
    assert (out - p["out"]).abs().max().item() <= 2e-5 * max(1.0, p["out"].abs().max().item())
    d = g["ddim"]
    noises = list(d["noises"])
    real = ref_ddim.noise_like
    ref_ddim.noise_like = lambda shape, device, repeat=False: noises.pop(0)
    try:
        with torch.no_grad():
            out2, _ = CpuDDIM(shim).sample(S=d["S"], batch_size=d["x_T"].shape[0], shape=tuple(d["x_T"].shape[1:]),
                                           conditioning=d["cond"], verbose=False,
                                           unconditional_guidance_scale=d["scale"], unconditional_conditioning=d["uc"],
                                           eta=d["eta"], x_T=d["x_T"])
    finally:
        ref_ddim.noise_like = real
    assert (out2 - d["out"]).abs().max().item() <= 2e-5 * max(1.0, d["out"].abs().max().item())
