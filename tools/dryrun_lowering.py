"""Dry run of the host-side graph builders WITHOUT a GPU: the C library is replaced by a recorder whose entry points all
succeed, tensors live on the CPU.  Catches Python-level mistakes (wrong arguments, shapes, missing attributes) of a
lowering before GPU time is spent on it.  Development aid only - nothing is computed.
usage: python tools/dryrun_lowering.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "q-diffusion_b200")):
    sys.path.insert(0, p)

from qdiff_b200 import _lib, first_stage, graph, ops  # noqa: E402


class _FakeLib:
    def __init__(self):
        self.ops = []

    def __getattr__(self, name):
        if name == "qd_engine_create":
            def create(dev, out):
                out._obj.value = 1
                return 0
            return create
        if name == "qd_engine_add_op":
            def add(e, kind, desc):
                self.ops.append(kind)
                return 0
            return add
        if name == "qd_groupnorm_workspace_floats":
            return lambda B, HW, C_, G: B * (HW // 64 + 1) * G * 4 + B * G * 2
        if name == "qd_launch_count":
            return lambda: 0
        return lambda *a, **k: 0


def main():
    fake = _FakeLib()
    for mod in (_lib, graph, ops, first_stage):
        if hasattr(mod, "lib"):
            mod.lib = lambda: fake
    ops._require_cuda = lambda *ts: None
    torch.cuda.is_available = lambda: True
    dev = torch.device("cpu")
    # ---- first stage
    for name, res in (("sd_v1", 16), ("lsun_bedroom", 16)):
        fs = first_stage.build_first_stage(name)
        cfg = first_stage.CONFIGS[name]
        zc = cfg["ddconfig"]["z_channels"]
        for prec in (1, 3, 6):
            n0 = len(fake.ops)
            b = first_stage.FirstStageBuilder(fs, dev, 2, prec)
            with torch.no_grad():
                x_in, t_in, out = b.lower(fs, (2, zc, res, res), cfg["kind"] == "vq")
            b.flush()
            print(f"first stage {name} precision {prec}: {len(fake.ops) - n0} ops, out {tuple(out.shape)}")
    # ---- weight-only / full-precision UNet states
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests.test_oracle_golden import WEIGHT_ONLY_LDM, load_case
    from tests.test_unet_gpu import build_qnn
    for name in WEIGHT_ONLY_LDM + ["ddim_w8_weightonly"]:
        g = load_case(name)
        qnn = build_qnn(g, dev)
        for state in ((True, False), (False, False)):
            qnn.set_quant_state(*state)
            n0 = len(fake.ops)
            b = graph.WeightOnlyBuilder(qnn, dev, g["x"].shape[0])
            with torch.no_grad():
                if g["family"] == "ddim":
                    b.lower_ddim(qnn.model, tuple(g["x"].shape))
                else:
                    b.lower_ldm(qnn.model, tuple(g["x"].shape), None if g["context"] is None else tuple(g["context"].shape))
            b.flush()
            print(f"{name} state {state}: {len(fake.ops) - n0} ops")
    # ---- INT8 lowering of the DDIM family (concat-free decoder): op counts with and without the in-place concat
    g = load_case("ddim_w4a8_split")
    qnn = build_qnn(g, dev)
    for mode in ("inplace", "copy"):
        os.environ["QDIFF_DDIM_CAT"] = mode
        b = graph.Builder(qnn, dev, g["x"].shape[0])
        with torch.no_grad():
            b.lower_ddim(qnn.model, tuple(g["x"].shape))
        b.flush()
        ncopy = sum(1 for k in b.op_kinds if k == _lib.QD_OP_COPY2D)
        print(f"ddim_w4a8_split INT8, QDIFF_DDIM_CAT={mode}: {b.nops} ops, {ncopy} copy2d")
    os.environ.pop("QDIFF_DDIM_CAT", None)


if __name__ == "__main__":
    main()
