"""Generate golden fixtures by RUNNING THE UNMODIFIED REFERENCE (imported read-only from
/root/reference) on tiny seeded UNets.  Run in the build container only (the GPU box has no
/root/reference); the resulting tests/golden/*.pt files are committed and travel.

    python tools/make_golden.py

For each case: build the reference model, wrap in qdiff.QuantModel, reproduce the --resume flow of
qdiff/utils.py:382-457 on CPU (weight-quantizer init by dummy forward, convert_adaround, act-quantizer
init), perturb alpha/delta so the fixture behaves like a calibrated checkpoint (hard AdaRound decisions
differ from nearest, activations clip), save the state dict in `ckpt.pth` format plus inputs, outputs and
per-block traces of a second forward.
"""
import os
import sys
import tempfile
import types

import torch
import torch.nn as nn

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")


def _import_reference():
    stub = tempfile.mkdtemp()
    os.makedirs(os.path.join(stub, "omegaconf"))
    open(os.path.join(stub, "omegaconf", "__init__.py"), "w").write("")
    open(os.path.join(stub, "omegaconf", "listconfig.py"), "w").write("class ListConfig(list):\n    pass\n")
    sys.path.insert(0, stub)
    sys.path.insert(0, REF)


def _rerandomize_zero_params(model, gen):
    with torch.no_grad():
        for p in model.parameters():
            if p.numel() > 1 and float(p.abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.05)


def _to_ckpt(qnn):
    """What the scripts do before torch.save(qnn.state_dict()) (sample_diffusion_ddim.py:223-234)."""
    from qdiff.adaptive_rounding import AdaRoundQuantizer
    from qdiff.quant_layer import UniformAffineQuantizer
    for m in qnn.model.modules():
        if isinstance(m, AdaRoundQuantizer):
            m.zero_point = nn.Parameter(m.zero_point)
            m.delta = nn.Parameter(m.delta)
        elif isinstance(m, UniformAffineQuantizer) and m.zero_point is not None:
            if not torch.is_tensor(m.zero_point):
                m.zero_point = nn.Parameter(torch.tensor(float(m.zero_point)))
            elif not isinstance(m.zero_point, nn.Parameter):
                m.zero_point = nn.Parameter(m.zero_point.float())
    return {k: v.detach().clone() for k, v in qnn.state_dict().items()}


def _from_ckpt_state(qnn):
    """Undo: plain tensors / python ints again, as after resume_cali_model (qdiff/utils.py:443-457)."""
    from qdiff.adaptive_rounding import AdaRoundQuantizer
    from qdiff.quant_layer import UniformAffineQuantizer
    for m in qnn.model.modules():
        if isinstance(m, AdaRoundQuantizer):
            z, d = m.zero_point.data, m.delta.data
            delattr(m, "zero_point"); delattr(m, "delta")
            m.zero_point, m.delta = z, d
        elif isinstance(m, UniformAffineQuantizer) and isinstance(m.zero_point, nn.Parameter):
            z = m.zero_point.item()
            delattr(m, "zero_point")
            m.zero_point = int(z)


def make_case(name, family, params, qcfg, batch, ctx_dim=None, seed=0):
    from qdiff import QuantModel
    from qdiff.adaptive_rounding import AdaRoundQuantizer
    from qdiff.quant_layer import UniformAffineQuantizer
    from qdiff.utils import convert_adaround
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    if family == "ddim":
        from ddim.models.diffusion import Model
        ns = types.SimpleNamespace
        cfg = ns(model=ns(type="simple", in_channels=params["in_channels"], out_ch=params["out_ch"], ch=params["ch"],
                          ch_mult=params["ch_mult"], num_res_blocks=params["num_res_blocks"],
                          attn_resolutions=params["attn_resolutions"], dropout=0.0, resamp_with_conv=True),
                 data=ns(image_size=params["resolution"]), diffusion=ns(num_diffusion_timesteps=1000),
                 split_shortcut=params["split_shortcut"])
        model = Model(cfg)
        in_ch, res = params["in_channels"], params["resolution"]
    else:
        from ldm.modules.diffusionmodules.openaimodel import UNetModel
        model = UNetModel(**params["unet"])
        model.split = params.get("split", False)
        in_ch, res = params["unet"]["in_channels"], params["res"]
    model.eval()
    _rerandomize_zero_params(model, gen)

    wq = {'n_bits': qcfg["weight_bit"], 'channel_wise': True, 'scale_method': 'max'}
    aq = {'n_bits': qcfg["act_bit"], 'symmetric': qcfg["a_sym"], 'channel_wise': False, 'scale_method': 'max',
          'leaf_param': qcfg["quant_act"]}
    qnn = QuantModel(model=model, weight_quant_params=wq, act_quant_params=aq, sm_abit=qcfg["sm_abit"])
    qnn.eval()
    qnn.set_grad_ckpt(False)

    def inputs(g):
        x = torch.randn(batch, in_ch, res, res, generator=g)
        t = torch.randint(0, 1000, (batch,), generator=g)
        c = torch.randn(batch, 7, ctx_dim, generator=g) if ctx_dim else None
        return x, t, c

    def fwd(x, t, c):
        with torch.no_grad():
            return qnn(x, t, c) if c is not None else qnn(x, t)

    cal = inputs(gen)
    qnn.set_quant_state(True, False)
    fwd(*cal)                                   # init weight quantizers (+ creates split quantizers)
    convert_adaround(qnn)
    with torch.no_grad():
        for m in qnn.model.modules():
            if isinstance(m, AdaRoundQuantizer):
                m.alpha.copy_(torch.rand(m.alpha.shape, generator=gen) * 2 - 1)
                m.delta = m.delta * (0.9 + 0.1 * torch.rand(m.delta.shape, generator=gen))
    if qcfg["quant_act"]:
        qnn.set_quant_state(True, True)
        fwd(*cal)                               # init act quantizers ('max' on the calibration batch)
        with torch.no_grad():
            for m in qnn.model.modules():
                if isinstance(m, UniformAffineQuantizer) and m.delta is not None:
                    m.delta.data.mul_(0.8 + 0.2 * torch.rand((), generator=gen))
    ckpt = _to_ckpt(qnn)
    # only the sign of alpha matters on the inference path (adaptive_rounding.py:55): store it as int8
    ckpt = {k: (torch.where(v >= 0, 1, -1).to(torch.int8) if k.endswith(".alpha") else v) for k, v in ckpt.items()}
    _from_ckpt_state(qnn)
    qnn.set_quant_state(True, qcfg["quant_act"])

    x, t, c = inputs(gen)
    traces = {}
    hooks = []

    def add_hook(mod, key):
        hooks.append(mod.register_forward_hook(lambda m, i, o, key=key: traces.__setitem__(key, o.detach().clone())))

    if family == "ldm":
        for i, b in enumerate(model.input_blocks):
            add_hook(b, f"input_blocks.{i}")
        add_hook(model.middle_block, "middle_block")
        for i, b in enumerate(model.output_blocks):
            add_hook(b, f"output_blocks.{i}")
    else:
        add_hook(model.mid.block_2, "mid")
    out = fwd(x, t, c)
    for h in hooks:
        h.remove()
    extra = {}
    if qcfg.get("save_fp"):                     # the reference's full-precision state on the same inputs
        qnn.set_quant_state(False, False)
        extra["out_fp"] = fwd(x, t, c)
        qnn.set_quant_state(True, qcfg["quant_act"])
    os.makedirs(OUT, exist_ok=True)
    # fp16-exact storage is not acceptable for a parity oracle: keep fp32
    torch.save(dict(name=name, family=family, params=params, qcfg=qcfg, ckpt=ckpt, x=x, t=t, context=c, out=out,
                    traces=traces, torch_version=torch.__version__, **extra), os.path.join(OUT, name + ".pt"))
    nq = sum(1 for k in ckpt if k.endswith("act_quantizer.delta"))
    print(f"{name}: {len(ckpt)} ckpt keys, {nq} act quantizers, out std {out.std().item():.4f}, "
          f"file {os.path.getsize(os.path.join(OUT, name + '.pt')) / 1e6:.2f} MB")


CASES = [
    # cfg 1: CIFAR-style DDIM UNet, 8-bit weights only (quant_act off): the reference's own CPU-runnable case
    ("ddim_w8_weightonly", "ddim",
     dict(in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2], num_res_blocks=1, attn_resolutions=[8], resolution=16,
          split_shortcut=False),
     dict(weight_bit=8, act_bit=8, a_sym=True, sm_abit=8, quant_act=False), None),
    # cfg 2: CIFAR-style DDIM UNet, W4A8 symmetric, split shortcut
    ("ddim_w4a8_split", "ddim",
     dict(in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2], num_res_blocks=1, attn_resolutions=[8], resolution=16,
          split_shortcut=True),
     dict(weight_bit=4, act_bit=8, a_sym=True, sm_abit=8, quant_act=True), None),
    # cfg 3: LDM with legacy multi-head attention (num_head_channels=32), W4A8 symmetric, no split
    ("ldm_legacy_w4a8", "ldm",
     dict(unet=dict(image_size=16, in_channels=3, out_channels=3, model_channels=32, attention_resolutions=[2, 1],
                    num_res_blocks=1, channel_mult=[1, 2], num_head_channels=32), res=16),
     dict(weight_bit=4, act_bit=8, a_sym=True, sm_abit=8, quant_act=True), None),
    # cfg 5: church-style scale-shift norm + resblock up/down, 8 bit asymmetric  (W4 here: W8 storage is a later row)
    ("ldm_updown_w4a8", "ldm",
     dict(unet=dict(image_size=16, in_channels=4, out_channels=4, model_channels=32, attention_resolutions=[1, 2],
                    num_res_blocks=1, channel_mult=[1, 2], num_heads=2, use_scale_shift_norm=True,
                    resblock_updown=True), res=16),
     dict(weight_bit=4, act_bit=8, a_sym=False, sm_abit=8, quant_act=True), None),
    # cfg 5 proper: the same church-style UNet with 8-bit weights (wq - zw spans [-255,255]: exercises the W8 operand split)
    ("ldm_updown_w8a8", "ldm",
     dict(unet=dict(image_size=16, in_channels=4, out_channels=4, model_channels=32, attention_resolutions=[1, 2],
                    num_res_blocks=1, channel_mult=[1, 2], num_heads=2, use_scale_shift_norm=True,
                    resblock_updown=True), res=16),
     dict(weight_bit=8, act_bit=8, a_sym=False, sm_abit=8, quant_act=True), None),
    # cfg 4: SD-style spatial transformer with cross attention, asymmetric W4A8, sm_abit 16, split shortcut
    ("sd_tiny_w4a8_sm16", "ldm",
     dict(unet=dict(image_size=16, in_channels=4, out_channels=4, model_channels=32, attention_resolutions=[2, 1],
                    num_res_blocks=1, channel_mult=[1, 2], num_heads=2, use_spatial_transformer=True,
                    transformer_depth=1, context_dim=64, legacy=False), res=16, split=True),
     dict(weight_bit=4, act_bit=8, a_sym=False, sm_abit=16, quant_act=True), 64),
    # weight-only (quant_act off) and full-precision states of the LDM / SD families: what `--ptq` without `--quant_act`
    # runs (qdiff/utils.py:407), plus set_quant_state(False, False) on the same inputs (out_fp)
    ("sd_tiny_w4_weightonly", "ldm",
     dict(unet=dict(image_size=16, in_channels=4, out_channels=4, model_channels=32, attention_resolutions=[2, 1],
                    num_res_blocks=1, channel_mult=[1, 2], num_heads=2, use_spatial_transformer=True,
                    transformer_depth=1, context_dim=64, legacy=False), res=16, split=True),
     dict(weight_bit=4, act_bit=8, a_sym=False, sm_abit=16, quant_act=False, save_fp=True), 64),
    ("ldm_updown_w8_weightonly", "ldm",
     dict(unet=dict(image_size=16, in_channels=4, out_channels=4, model_channels=32, attention_resolutions=[1, 2],
                    num_res_blocks=1, channel_mult=[1, 2], num_heads=2, use_scale_shift_norm=True,
                    resblock_updown=True), res=16),
     dict(weight_bit=8, act_bit=8, a_sym=False, sm_abit=8, quant_act=False, save_fp=True), None),
    ("ldm_legacy_w4_weightonly", "ldm",
     dict(unet=dict(image_size=16, in_channels=3, out_channels=3, model_channels=32, attention_resolutions=[2, 1],
                    num_res_blocks=1, channel_mult=[1, 2], num_head_channels=32), res=16),
     dict(weight_bit=4, act_bit=8, a_sym=True, sm_abit=8, quant_act=False, save_fp=True), None),
]


def make_quantizer_kats():
    """Known-answer vectors straight from the reference quantizer classes."""
    from qdiff.quant_layer import UniformAffineQuantizer
    g = torch.Generator().manual_seed(5)
    kats = []
    for n_bits, sym, az in [(8, False, False), (8, True, False), (4, False, False), (16, False, True)]:
        x = torch.randn(4, 37, generator=g) * 3
        if az:
            x = x.abs() / 10
        q = UniformAffineQuantizer(n_bits=n_bits, symmetric=sym, channel_wise=False, scale_method='max', always_zero=az)
        y = q(x)
        kats.append(dict(n_bits=n_bits, symmetric=sym, always_zero=az, x=x, y=y, delta=q.delta.clone(),
                         zero_point=float(q.zero_point)))
    # ties: round-half-even, and the symmetric clamp [-128, 127]
    q = UniformAffineQuantizer(n_bits=8, symmetric=True)
    q.delta, q.zero_point, q.inited = torch.tensor(1.0), 0, True
    xt = torch.tensor([0.5, 1.5, 2.5, 3.5, -200.0, -128.4, 126.5, 127.5, 300.0])
    kats.append(dict(n_bits=8, symmetric=True, always_zero=False, x=xt, y=q(xt), delta=torch.tensor(1.0), zero_point=0.0))
    # channel-wise weight init
    w = torch.randn(6, 5, 3, 3, generator=g)
    qw = UniformAffineQuantizer(n_bits=4, symmetric=False, channel_wise=True, scale_method='max')
    yw = qw(w)
    torch.save(dict(act=kats, weight=dict(w=w, y=yw, delta=qw.delta.clone(), zero_point=qw.zero_point.clone())),
               os.path.join(OUT, "quantizer_kats.pt"))
    print("quantizer_kats: ok")


if __name__ == "__main__":
    _import_reference()
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])
    if not only:
        make_quantizer_kats()
    seeds = {"ddim_w4a8_split": 100, "ldm_legacy_w4a8": 101, "ldm_updown_w4a8": 102, "sd_tiny_w4a8_sm16": 103,
             "ldm_updown_w8a8": 104, "ddim_w8_weightonly": 105, "sd_tiny_w4_weightonly": 106,
             "ldm_updown_w8_weightonly": 107, "ldm_legacy_w4_weightonly": 108}
    for (name, family, params, qcfg, ctx) in CASES:
        if only and name not in only:
            continue
        make_case(name, family, params, qcfg, batch=2, ctx_dim=ctx, seed=seeds[name])
