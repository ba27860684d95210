"""Build the activation-quantizer fixture for the FULL-SIZE UNets of BASELINE.json (no pretrained or
calibrated checkpoints exist offline).

    python tools/make_calib.py sd_v1 cifar10 ...      -> tests/golden/calib_<name>.json

Procedure: seeded synthetic weights (qdiff_b200.unet.randomize_), channel-wise
'max' weight quantizers, seeded AdaRound alpha signs, then ONE forward of the CPU oracle in quick-init
mode on a seeded batch, which gives every activation quantizer the reference's 'max' initialisation
on that batch.  Only the ~1e3 activation scalars are stored; weights are regenerated from the seed.

This is calibration-side tooling: it imports the oracle, the product never does.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "q-diffusion_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from qdiff_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main(names):
    from oracle import synth_cfg
    for name in names:
        spec = synth.SPECS[name]
        t0 = time.time()
        model = synth.build_model(name)
        ckpt = synth.weight_ckpt(name, model)
        x, t, ctx = synth.calib_inputs(name)
        before = set(ckpt)
        with torch.no_grad():
            synth_cfg.oracle_forward(name, ckpt, x, t, ctx, init_missing=True)
        act = {k: float(v) for k, v in ckpt.items() if k not in before}
        os.makedirs(OUT, exist_ok=True)
        path = os.path.join(OUT, f"calib_{name}.json")
        json.dump(dict(name=name, spec=spec, act=act, torch_version=torch.__version__), open(path, "w"), indent=0)
        print(f"{name}: {len(act) // 2} activation quantizers -> {path} ({time.time() - t0:.1f}s)")


if __name__ == "__main__":
    main(sys.argv[1:] or ["cifar10", "sd_v1"])
