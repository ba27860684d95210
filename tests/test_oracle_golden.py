"""Pin the CPU oracle against outputs of the reference itself (fixtures from tools/make_golden.py,
which imports /root/reference).  CPU-only: runs in the build container and on the GPU box."""
import os

import pytest
import torch

from oracle import ops_oracle as O
from oracle import unet_oracle as U

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["ddim_w4a8_split", "ldm_legacy_w4a8", "ldm_updown_w4a8", "sd_tiny_w4a8_sm16", "ldm_updown_w8a8"]
# BASELINE configs[0] (weight-only W8, the reference's own CPU-runnable case): kept in a separate list because the name
# says what it pins first - the oracle against the reference; the GPU tests run the engine's weight-only lowering on it too
ORACLE_ONLY = ["ddim_w8_weightonly"]
# weight-only LDM / SD fixtures; they also carry the reference's full-precision output (set_quant_state(False, False))
WEIGHT_ONLY_LDM = ["sd_tiny_w4_weightonly", "ldm_updown_w8_weightonly", "ldm_legacy_w4_weightonly"]


def load_case(name):
    g = torch.load(os.path.join(GOLD, name + ".pt"), map_location="cpu", weights_only=False)
    g["ckpt"] = {k: (v.float() if k.endswith(".alpha") else v) for k, v in g["ckpt"].items()}
    return g


def oracle_forward(g, trace=None, dtype=torch.float32, weight_quant=True):
    """dtype=float64 evaluates the SAME algorithm with (almost) no rounding noise: the distance between that
    and the fp32 result is the reference's intrinsic noise band (fake-quant networks amplify ulp-level
    perturbations up to quantisation-noise level within a few layers; see DESIGN.md, Parity)."""
    q = g["qcfg"]
    qc = U.QuantCfg(q["weight_bit"], q["act_bit"], q["a_sym"], q["sm_abit"], q["quant_act"] and weight_quant, adaround=True,
                    weight_quant=weight_quant)
    te = (O.timestep_embedding_ldm, O.timestep_embedding_ddim)
    U.set_dtype(dtype)
    if dtype != torch.float32:   # keep the fp32 sinusoid table, widen afterwards
        O.timestep_embedding_ldm = lambda t, dim, **k: te[0](t, dim).to(dtype)
        O.timestep_embedding_ddim = lambda t, dim: te[1](t, dim).to(dtype)
    try:
        x = g["x"].to(dtype)
        ctx = g["context"].to(dtype) if g["context"] is not None else None
        if g["family"] == "ddim":
            p = g["params"]
            cfg = dict(ch=p["ch"], ch_mult=p["ch_mult"], num_res_blocks=p["num_res_blocks"],
                       attn_resolutions=p["attn_resolutions"], resolution=p["resolution"], resamp_with_conv=True,
                       split_shortcut=p["split_shortcut"])
            return U.ddim_unet_forward(g["ckpt"], cfg, qc, x, g["t"], trace=trace)
        arch = U.ldm_arch_from_params(split=g["params"].get("split", False), **g["params"]["unet"])
        return U.ldm_unet_forward(g["ckpt"], arch, qc, x, g["t"], ctx, trace=trace)
    finally:
        U.set_dtype(torch.float32)
        O.timestep_embedding_ldm, O.timestep_embedding_ddim = te


def noise_band_mse(g, ref):
    """MSE between the fp64 evaluation of the reference algorithm and its fp32 result `ref`."""
    return ((oracle_forward(g, dtype=torch.float64).double() - ref.double()) ** 2).mean().item()


def test_quantizer_known_answers():
    k = torch.load(os.path.join(GOLD, "quantizer_kats.pt"), map_location="cpu", weights_only=False)
    for kat in k["act"]:
        y = O.uaq_fake_quant(kat["x"], kat["delta"], kat["zero_point"], kat["n_bits"], kat["symmetric"])
        assert torch.equal(y, kat["y"]), kat["n_bits"]
    # init: the reference derived (delta, zp) from the tensor; so must the oracle
    for kat in k["act"][:4]:
        d, z = O.uaq_init_max(kat["x"], kat["n_bits"], kat["symmetric"], kat["always_zero"])
        assert torch.equal(d, kat["delta"]) and float(z) == kat["zero_point"]
    # ties and symmetric clamp
    ties = k["act"][4]
    assert ties["y"].tolist() == [0.0, 2.0, 2.0, 4.0, -128.0, -128.0, 126.0, 127.0, 127.0]
    w = k["weight"]
    d, z = O.weight_init_max(w["w"], 4)
    assert torch.equal(d, w["delta"].flatten()) and torch.equal(z, w["zero_point"].flatten())
    assert torch.equal(O.uaq_weight_fake_quant(w["w"], d, z, 4), w["y"])


@pytest.mark.parametrize("name", WEIGHT_ONLY_LDM)
def test_oracle_full_precision_state_matches_reference(name):
    """set_quant_state(False, False) of the reference (FP baselines / calibration data) vs the oracle without quantizers."""
    g = load_case(name)
    out = oracle_forward(g, weight_quant=False)
    ref = g["out_fp"]
    assert (out - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
    assert (ref - g["out"]).abs().max().item() > 1e-3        # and it is a different function than the weight-only state


@pytest.mark.parametrize("name", CASES + ORACLE_ONLY + WEIGHT_ONLY_LDM)
def test_oracle_matches_reference(name):
    g = load_case(name)
    trace = {}
    out = oracle_forward(g, trace)
    ref = g["out"]
    err = (out - ref).abs().max().item()
    mse = ((out - ref) ** 2).mean().item()
    # same fp32 torch ops in the same order -> expect (near) bit equality; allow reduction-order noise only
    assert err <= 1e-5 * max(1.0, ref.abs().max().item()), (name, err, mse)
    for key, t in g["traces"].items():
        if key in trace and key != "layers":
            e = (trace[key] - t).abs().max().item()
            assert e <= 1e-5 * max(1.0, t.abs().max().item()), (name, key, e)


def test_sampler_oracle_matches_reference_samplers():
    """oracle/sampler_oracle.py vs the reference's own PLMSSampler.sample and generalized_steps run around the same toy
    eps-model (fixture: tools/make_sampler_golden.py).  The GPU sampler tests compare the engine with this oracle."""
    from oracle import sampler_oracle as S
    from tools.make_sampler_golden import toy_eps
    g = torch.load(os.path.join(GOLD, "samplers.pt"), map_location="cpu", weights_only=False)
    p = g["plms"]
    ac = S.ldm_schedule(1000, p["linear_start"], p["linear_end"])
    out = S.plms_sample(lambda x, t, c: toy_eps(x, t, c), p["x_T"], p["cond"], p["uc"], p["scale"], ac, p["S"])
    err = (out - p["out"]).abs().max().item()
    assert err <= 2e-5 * max(1.0, p["out"].abs().max().item()), err
    q = g["generalized"]
    out2 = S.generalized_steps(lambda x, t: toy_eps(x, t), q["x"], q["seq"], q["betas"], eta=0.0)
    err2 = (out2 - q["out"]).abs().max().item()
    assert err2 <= 2e-5 * max(1.0, q["out"].abs().max().item()), err2
    # DDIMSampler.sample with eta = 1 + classifier-free guidance, fed the noise the reference drew (ddim.py:170-220)
    d = g["ddim"]
    ac = S.ldm_schedule(1000, d["linear_start"], d["linear_end"])
    out3 = S.ddim_sample(lambda x, t, c: toy_eps(x, t, c), d["x_T"], d["cond"], d["uc"], d["scale"], ac, d["S"],
                         eta=d["eta"], noises=d["noises"])
    err3 = (out3 - d["out"]).abs().max().item()
    assert len(d["noises"]) == d["S"] and err3 <= 2e-5 * max(1.0, d["out"].abs().max().item()), err3
    # generalized_steps on the quadratic schedule (cfg 2) with eta = 1 (denoising.py:25-29)
    q2 = g["generalized_quad"]
    out4 = S.generalized_steps(lambda x, t: toy_eps(x, t), q2["x"], q2["seq"], q2["betas"], eta=q2["eta"],
                               noises=q2["noises"])
    err4 = (out4 - q2["out"]).abs().max().item()
    assert err4 <= 2e-5 * max(1.0, q2["out"].abs().max().item()), err4
    # DPM-Solver++ (2M) as DPMSolverSampler drives it (--dpm): 6 steps (lower_order_final active) and 20 steps
    dp = g["dpm"]
    ac = S.ldm_schedule(1000, dp["linear_start"], dp["linear_end"])
    for steps, ref in dp["out"].items():
        out5 = S.dpm_solver_sample(lambda x, t, c: toy_eps(x, t, c), dp["x_T"], dp["cond"], dp["uc"], dp["scale"], ac, steps)
        err5 = (out5 - ref).abs().max().item()
        assert err5 <= 5e-5 * max(1.0, ref.abs().max().item()), (steps, err5)
