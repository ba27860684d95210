"""ORACLE side helper (test infrastructure): architecture dicts for the CPU oracle matching the
synthetic full-size workloads of qdiff_b200.synth (BASELINE.json configs)."""
from qdiff_b200 import synth, unet

from . import unet_oracle as U


def oracle_cfg(name):
    spec = synth.SPECS[name]
    if spec["family"] == "ddim":
        c = unet.ddim_config(split_shortcut=spec["split"])
        return dict(ch=c.model.ch, ch_mult=c.model.ch_mult, num_res_blocks=c.model.num_res_blocks,
                    attn_resolutions=c.model.attn_resolutions, resolution=c.data.image_size, resamp_with_conv=True,
                    split_shortcut=spec["split"])
    return U.ldm_arch_from_params(split=spec["split"], **dict(unet.CONFIGS[name]["params"]))


def quant_cfg(name):
    s = synth.SPECS[name]
    return U.QuantCfg(s["weight_bit"], s["act_bit"], s["a_sym"], s["sm_abit"], True, adaround=True)


def oracle_forward(name, ckpt, x, t, ctx=None, **kw):
    if synth.SPECS[name]["family"] == "ddim":
        return U.ddim_unet_forward(ckpt, oracle_cfg(name), quant_cfg(name), x, t, **kw)
    return U.ldm_unet_forward(ckpt, oracle_cfg(name), quant_cfg(name), x, t, ctx, **kw)
