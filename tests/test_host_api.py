"""CPU-side checks of the host mirror: state-dict key compatibility with the reference's ckpt.pth
format, the C-ABI library loads and exports every symbol the header declares, and the product path
refuses to run without CUDA (no CPU fallback)."""
import os
import re

import pytest
import torch

from tests.test_oracle_golden import CASES, load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from qdiff_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "qdiff_b200.h")).read()
    declared = set(re.findall(r"\b(qd_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = _lib.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/qdiff_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)


@pytest.mark.parametrize("name", CASES)
def test_checkpoint_keys_roundtrip(name):
    """Wrapping our containers + resume_cali_model must consume exactly the reference's checkpoint keys."""
    from tests.test_unet_gpu import build_qnn
    g = load_case(name)
    qnn = build_qnn(g, torch.device("cpu"))
    mods = dict(qnn.named_modules())
    n_split = sum(1 for k in g["ckpt"] if k.endswith("weight_quantizer_0.delta"))
    assert sum(1 for m in mods.values() if type(m).__name__ == "QuantModule" and m.split) == n_split
    # quantizer values landed where the graph builder reads them
    for k, v in g["ckpt"].items():
        if k.endswith(".act_quantizer.delta"):
            q = mods[k[:-len(".delta")]]
            assert float(q.delta) == float(v) and q.inited


def test_forward_without_cuda_fails_loudly():
    from tests.test_unet_gpu import build_qnn
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = load_case("ddim_w4a8_split")
    qnn = build_qnn(g, torch.device("cpu"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        qnn(g["x"], g["t"])
    with pytest.raises(RuntimeError):
        qnn.model.conv_in(g["x"])
