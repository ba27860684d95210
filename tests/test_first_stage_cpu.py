"""First-stage decode, CPU side: the oracle is pinned to outputs of the reference's own classes
(fixtures from tools/make_decoder_golden.py), and the product containers take the reference's checkpoints key for key."""
import os

import pytest
import torch

from oracle import decoder_oracle as D

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["decoder_kl_tiny", "decoder_vq_tiny"]


def load(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), map_location="cpu", weights_only=False)


@pytest.mark.parametrize("name", CASES)
def test_decoder_oracle_matches_reference(name):
    """AutoencoderKL.decode / VQModelInterface.decode of the reference (run unmodified through import stubs) vs the oracle."""
    g = load(name)
    out = D.decode_first_stage(g["sd"], g["z"], g["scale_factor"], g["kind"])
    assert out.shape == g["out"].shape
    assert float((out - g["out"]).abs().max()) <= 1e-6 * float(g["out"].abs().max())
    if g["kind"] == "kl":      # the bare Decoder class (no import stubs involved at all)
        o2 = D.decoder_forward(g["sd"], g["z_dec"])
        assert float((o2 - g["out_dec"]).abs().max()) <= 1e-6 * float(g["out_dec"].abs().max())
    else:
        o2 = D.vq_decode(g["sd"], g["z"], force_not_quantize=True)
        assert float((o2 - g["out_not_quantized"]).abs().max()) <= 1e-6 * float(g["out_not_quantized"].abs().max())
        _, idx = D.vq_quantize(g["sd"], g["z"], return_indices=True)
        assert torch.equal(idx, g["indices"])


@pytest.mark.parametrize("name", CASES)
def test_containers_take_reference_checkpoints(name):
    """State-dict keys and shapes of the product containers == what the reference's modules saved."""
    from qdiff_b200 import first_stage as FS
    g = load(name)
    cfg = dict(kind=g["kind"], embed_dim=g["embed_dim"], ddconfig=g["ddconfig"], n_embed=g.get("n_embed"))
    fs = FS.build_first_stage(cfg)
    own = {k: tuple(v.shape) for k, v in fs.state_dict().items()}
    ref = {k: tuple(v.shape) for k, v in g["sd"].items()}
    assert own == ref
    fs.load_state_dict(g["sd"], strict=True)
    with pytest.raises(RuntimeError):           # no torch / CPU path
        fs.decode(g["z"])
    with pytest.raises(RuntimeError):
        fs.decoder(g["z"])


def test_named_first_stage_configs_build():
    """The three first stages of BASELINE's LDM configs: parameter counts of the decode side."""
    from qdiff_b200 import first_stage as FS
    with torch.device("meta"):
        n = {k: sum(p.numel() for p in FS.build_first_stage(k).parameters()) for k in FS.CONFIGS}
    assert n["sd_v1"] == n["lsun_church"] and n["sd_v1"] > 45e6        # kl-f8 decoder: 49.5 M parameters
    assert 10e6 < n["lsun_bedroom"] < n["sd_v1"]
