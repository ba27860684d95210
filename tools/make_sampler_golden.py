"""Golden fixtures for the sampler loops, produced by RUNNING THE REFERENCE's sampler code (imported read-only from
/root/reference) around a small deterministic eps-model.  Build container only; the fixture is committed.

    python tools/make_sampler_golden.py      -> tests/golden/samplers.pt

PLMS: ldm/models/diffusion/plms.py (PLMSSampler.sample), DDIM-style generalized steps: ddim/functions/denoising.py.
The reference's register_buffer() moves every schedule tensor to "cuda" unconditionally (plms.py:19-23); the subclass
below only overrides that one helper so the loop itself runs unmodified on the CPU.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.make_golden import OUT, _import_reference  # noqa: E402


def toy_eps(x, t, context=None):
    """Smooth, deterministic, batch-wise eps-model: the same callable drives the reference and the oracle."""
    tt = (t.float() / 1000.0).reshape(-1, 1, 1, 1)
    e = torch.tanh(0.8 * x + 0.3 * torch.sin(6.0 * tt) + 0.1 * torch.roll(x, 1, dims=-1))
    if context is not None:
        e = e + 0.05 * context.float().mean(dim=(1, 2)).reshape(-1, 1, 1, 1)
    return e


def main():
    _import_reference()
    from ldm.models.diffusion.plms import PLMSSampler
    from ddim.functions.denoising import generalized_steps

    from oracle.sampler_oracle import ldm_schedule
    g = torch.Generator().manual_seed(7)
    ac = ldm_schedule(1000, 0.00085, 0.0120)          # configs/stable-diffusion/v1-inference.yaml
    betas = 1.0 - torch.cat([ac[:1], ac[1:] / ac[:-1]])

    class ToyLDM:                                       # the attributes PLMSSampler reads from LatentDiffusion
        num_timesteps = 1000
        device = torch.device("cpu")
        alphas_cumprod = ac
        alphas_cumprod_prev = torch.cat([torch.ones(1), ac[:-1]])
        betas_ = betas

        def __init__(self):
            self.betas = betas

        def apply_model(self, x, t, c):
            return toy_eps(x, t, c)

    class CpuPLMS(PLMSSampler):
        def register_buffer(self, name, attr):
            setattr(self, name, attr)

    B, shape, S, scale = 2, (4, 8, 8), 10, 3.0
    x_T = torch.randn(B, *shape, generator=g)
    cond = torch.randn(B, 5, 16, generator=g)
    uc = torch.randn(1, 5, 16, generator=g).expand(B, 5, 16).contiguous()
    sampler = CpuPLMS(ToyLDM())
    with torch.no_grad():
        plms_out, _ = sampler.sample(S=S, batch_size=B, shape=shape, conditioning=cond, verbose=False,
                                     unconditional_guidance_scale=scale, unconditional_conditioning=uc, eta=0.0,
                                     x_T=x_T)

    # ddim (CIFAR path): linear beta schedule of the cifar10 config, 20 uniform steps, eta = 0
    betas_c = torch.linspace(0.0001, 0.02, 1000, dtype=torch.float64).float()
    seq = list(range(0, 1000, 50))
    x0 = torch.randn(B, 3, 8, 8, generator=g)
    # denoising.py:21 hard-codes xs[-1].to('cuda'): shim Tensor.to so that 'cuda' means the CPU for this one call
    real_to = torch.Tensor.to

    def to_cpu_shim(self, *a, **k):
        a = tuple("cpu" if (isinstance(v, str) and v.startswith("cuda")) else v for v in a)
        return real_to(self, *a, **k)

    torch.Tensor.to = to_cpu_shim
    try:
        with torch.no_grad():
            xs, _ = generalized_steps(x0, seq, lambda x, t: toy_eps(x, t), betas_c, eta=0.0)
    finally:
        torch.Tensor.to = real_to
    os.makedirs(OUT, exist_ok=True)
    torch.save(dict(plms=dict(x_T=x_T, cond=cond, uc=uc, scale=scale, S=S, out=plms_out,
                              linear_start=0.00085, linear_end=0.0120),
                    generalized=dict(x=x0, seq=seq, betas=betas_c, out=xs[-1])),
               os.path.join(OUT, "samplers.pt"))
    print("samplers.pt:", tuple(plms_out.shape), float(plms_out.std()), tuple(xs[-1].shape), float(xs[-1].std()))


if __name__ == "__main__":
    main()
