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
    from ldm.models.diffusion import ddim as ref_ddim
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

    # DDIM with eta = 1 and classifier-free guidance (cfg 3 runs `-e 1.0`): the per-step noise draws are recorded by
    # wrapping the reference's noise_like (util.py:264-267) so the oracle / engine can be fed the same noise
    class CpuDDIM(ref_ddim.DDIMSampler):
        def register_buffer(self, name, attr):
            setattr(self, name, attr)

    ddim_noises = []
    real_noise_like = ref_ddim.noise_like

    def recording_noise_like(shape, device, repeat=False):
        n = real_noise_like(shape, device, repeat)
        ddim_noises.append(n.clone())
        return n

    ref_ddim.noise_like = recording_noise_like
    x_T2 = torch.randn(B, *shape, generator=g)
    torch.manual_seed(11)
    try:
        with torch.no_grad():
            ddim_out, _ = CpuDDIM(ToyLDM()).sample(S=8, batch_size=B, shape=shape, conditioning=cond, verbose=False,
                                                   unconditional_guidance_scale=2.0, unconditional_conditioning=uc,
                                                   eta=1.0, x_T=x_T2)
    finally:
        ref_ddim.noise_like = real_noise_like

    # DPM-Solver++ (2M) through the reference's DPMSolverSampler (--dpm); its register_buffer also forces "cuda"
    from ldm.models.diffusion.dpm_solver import DPMSolverSampler

    class CpuDPM(DPMSolverSampler):
        def register_buffer(self, name, attr):
            setattr(self, name, attr)

    x_T3 = torch.randn(B, *shape, generator=g)
    dpm_out = {}
    for S_dpm in (6, 20):          # 6: lower_order_final active (steps < 15); 20: plain second order
        with torch.no_grad():
            o, _ = CpuDPM(ToyLDM()).sample(S=S_dpm, batch_size=B, shape=shape, conditioning=cond, verbose=False,
                                           unconditional_guidance_scale=2.5, unconditional_conditioning=uc, x_T=x_T3)
        dpm_out[S_dpm] = o

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
        # cfg 2 style: quadratic timestep schedule (sample_diffusion_ddim.py:294-301) with eta = 1; the per-step
        # torch.randn_like draws (denoising.py:28) are recorded
        import numpy as np
        seq_q = [int(s) for s in list(np.linspace(0, np.sqrt(1000 * 0.8), 12) ** 2)]
        gs_noises = []
        real_randn_like = torch.randn_like

        def recording_randn_like(t, *a, **k):
            n = real_randn_like(t, *a, **k)
            gs_noises.append(n.clone())
            return n

        x1 = torch.randn(B, 3, 8, 8, generator=g)
        torch.manual_seed(13)
        torch.randn_like = recording_randn_like
        try:
            with torch.no_grad():
                xs_q, _ = generalized_steps(x1, seq_q, lambda x, t: toy_eps(x, t), betas_c, eta=1.0)
        finally:
            torch.randn_like = real_randn_like
    finally:
        torch.Tensor.to = real_to
    os.makedirs(OUT, exist_ok=True)
    torch.save(dict(plms=dict(x_T=x_T, cond=cond, uc=uc, scale=scale, S=S, out=plms_out,
                              linear_start=0.00085, linear_end=0.0120),
                    generalized=dict(x=x0, seq=seq, betas=betas_c, out=xs[-1]),
                    ddim=dict(x_T=x_T2, cond=cond, uc=uc, scale=2.0, S=8, eta=1.0, noises=ddim_noises, out=ddim_out,
                              linear_start=0.00085, linear_end=0.0120),
                    generalized_quad=dict(x=x1, seq=seq_q, betas=betas_c, eta=1.0, noises=gs_noises, out=xs_q[-1]),
                    dpm=dict(x_T=x_T3, cond=cond, uc=uc, scale=2.5, out=dpm_out, linear_start=0.00085, linear_end=0.0120)),
               os.path.join(OUT, "samplers.pt"))
    print("samplers.pt:", tuple(plms_out.shape), float(plms_out.std()), tuple(xs[-1].shape), float(xs[-1].std()))


if __name__ == "__main__":
    main()
