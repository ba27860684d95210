"""Host-side mirror of qdiff/quant_block.py: block wrappers that own the extra activation
quantizers of the attention matmuls (reference quant_block.py:20-401).

Each wrapper adopts the children of the block it replaces (same attribute names, hence the same
checkpoint keys).  `forward` of a block is never executed: qdiff_b200.graph lowers the whole tree
to one engine program.  Dispatch is by class NAME so both qdiff_b200.unet containers and the
reference's own ldm/ddim module objects are accepted.
"""
import torch.nn as nn

from .quant_layer import QuantModule, StraightThrough, UniformAffineQuantizer


def _engine_only(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__}: blocks execute inside the CUDA engine program "
                       "(QuantModel.forward); no eager block forward exists.")


class BaseQuantBlock(nn.Module):
    def __init__(self, act_quant_params: dict = {}):
        super().__init__()
        self.use_weight_quant = False
        self.use_act_quant = False
        self.act_quantizer = UniformAffineQuantizer(**act_quant_params)  # constructed, never applied (as upstream)
        self.activation_function = StraightThrough()
        self.ignore_reconstruction = False

    forward = _engine_only

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_weight_quant, self.use_act_quant = weight_quant, act_quant
        for m in self.modules():
            if isinstance(m, QuantModule):
                m.set_quant_state(weight_quant, act_quant)


def _adopt(dst, src, names):
    for n in names:
        if hasattr(src, n):
            setattr(dst, n, getattr(src, n))


class QuantResBlock(BaseQuantBlock):
    def __init__(self, res, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        _adopt(self, res, ["channels", "emb_channels", "dropout", "out_channels", "use_conv", "use_checkpoint",
                           "use_scale_shift_norm", "in_layers", "updown", "up", "down", "h_upd", "x_upd", "emb_layers",
                           "out_layers", "skip_connection"])


class QuantQKMatMul(BaseQuantBlock):
    def __init__(self, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        self.scale = None
        self.act_quantizer_q = UniformAffineQuantizer(**act_quant_params)
        self.act_quantizer_k = UniformAffineQuantizer(**act_quant_params)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_act_quant = act_quant


class QuantSMVMatMul(BaseQuantBlock):
    def __init__(self, act_quant_params: dict = {}, sm_abit=8):
        super().__init__(act_quant_params)
        self.act_quantizer_v = UniformAffineQuantizer(**act_quant_params)
        w = dict(act_quant_params, n_bits=sm_abit, symmetric=False, always_zero=True)
        self.act_quantizer_w = UniformAffineQuantizer(**w)

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.use_act_quant = act_quant


class QuantAttentionBlock(BaseQuantBlock):
    def __init__(self, attn, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        _adopt(self, attn, ["channels", "num_heads", "use_checkpoint", "norm", "qkv", "attention", "proj_out"])


class QuantBasicTransformerBlock(BaseQuantBlock):
    def __init__(self, tran, act_quant_params: dict = {}, sm_abit: int = 8):
        super().__init__(act_quant_params)
        _adopt(self, tran, ["attn1", "ff", "attn2", "norm1", "norm2", "norm3", "checkpoint"])
        w = dict(act_quant_params, n_bits=sm_abit, always_zero=True)
        for attn in (self.attn1, self.attn2):
            attn.act_quantizer_q = UniformAffineQuantizer(**act_quant_params)
            attn.act_quantizer_k = UniformAffineQuantizer(**act_quant_params)
            attn.act_quantizer_v = UniformAffineQuantizer(**act_quant_params)
            attn.act_quantizer_w = UniformAffineQuantizer(**w)
            attn.use_act_quant = False

    def set_quant_state(self, weight_quant: bool = False, act_quant: bool = False):
        self.attn1.use_act_quant = act_quant
        self.attn2.use_act_quant = act_quant
        super().set_quant_state(weight_quant, act_quant)


class QuantResnetBlock(BaseQuantBlock):
    def __init__(self, res, act_quant_params: dict = {}):
        super().__init__(act_quant_params)
        _adopt(self, res, ["in_channels", "out_channels", "use_conv_shortcut", "norm1", "conv1", "temb_proj", "norm2",
                           "dropout", "conv2", "conv_shortcut", "nin_shortcut"])


class QuantAttnBlock(BaseQuantBlock):
    def __init__(self, attn, act_quant_params: dict = {}, sm_abit=8):
        super().__init__(act_quant_params)
        _adopt(self, attn, ["in_channels", "norm", "q", "k", "v", "proj_out"])
        self.act_quantizer_q = UniformAffineQuantizer(**act_quant_params)
        self.act_quantizer_k = UniformAffineQuantizer(**act_quant_params)
        self.act_quantizer_v = UniformAffineQuantizer(**act_quant_params)
        self.act_quantizer_w = UniformAffineQuantizer(**dict(act_quant_params, n_bits=sm_abit))


def get_specials(quant_act=False):
    """class NAME -> wrapper (reference quant_block.py:389-401 keys on the classes themselves)."""
    specials = {
        "ResBlock": QuantResBlock,
        "BasicTransformerBlock": QuantBasicTransformerBlock,
        "ResnetBlock": QuantResnetBlock,
        "AttnBlock": QuantAttnBlock,
    }
    if quant_act:
        specials["QKMatMul"] = QuantQKMatMul
        specials["SMVMatMul"] = QuantSMVMatMul
    else:
        specials["AttentionBlock"] = QuantAttentionBlock
    return specials
