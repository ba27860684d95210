"""Oracle outputs for the FULL-SIZE UNets of BASELINE.json, computed once in the build container so that the GPU
parity tests do not spend GPU-box minutes on CPU oracle evaluations (SD v1-4 at UNet batch 16 is ~10 min of fp32 + fp64
oracle time).

    python tools/make_fullsize_golden.py [name:batch ...]     -> tests/golden/fullsize_<name>_b<batch>.pt

Stored per case: the seed of the inputs (regenerated with qdiff_b200.synth.calib_inputs), the CPU oracle's eps in fp32
(`ref`: the reference algorithm, pinned by tests/test_oracle_golden.py) and the same algorithm evaluated in float64
(`hi`: its distance from `ref` is the reference's own fp32 noise band).  Samples are evaluated one at
a time (the UNets have no cross-sample op), which bounds memory for the 4096^2 attention maps.
Weights / quantizers: seeded synthetic + tests/golden/calib_<name>.json, exactly what synth.build_qnn gives the engine.
"""
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
SEED = 777
DEFAULT = ["cifar10:2", "lsun_church:2", "lsun_bedroom:1", "sd_v1:1", "sd_v1:16"]


def main(items):
    from oracle import synth_cfg
    from oracle import unet_oracle as U
    for it in items:
        name, batch = it.split(":")
        batch = int(batch)
        t0 = time.time()
        _, ckpt = synth.full_ckpt(name)
        ckpt = {k: (v.float() if k.endswith(".alpha") else v) for k, v in ckpt.items()}
        x, t, ctx = synth.calib_inputs(name, batch=batch, seed=SEED)
        refs, his = [], []
        with torch.no_grad():
            for b in range(batch):
                xb, tb = x[b:b + 1], t[b:b + 1]
                cb = ctx[b:b + 1] if ctx is not None else None
                refs.append(synth_cfg.oracle_forward(name, ckpt, xb, tb, cb))
                U.set_dtype(torch.float64)
                try:
                    his.append(synth_cfg.oracle_forward(name, ckpt, xb.double(), tb, cb.double() if cb is not None else None).float())
                finally:
                    U.set_dtype(torch.float32)
                print(f"  {name} sample {b + 1}/{batch}  ({time.time() - t0:.0f}s)", flush=True)
        ref, hi = torch.cat(refs), torch.cat(his)
        band = ((hi.double() - ref.double()) ** 2).mean().item()
        path = os.path.join(OUT, f"fullsize_{name}_b{batch}.pt")
        torch.save(dict(name=name, batch=batch, seed=SEED, spec=synth.SPECS[name], ref=ref, hi=hi, band_mse=band,
                        torch_version=torch.__version__), path)
        print(f"{name} B={batch}: ref std {ref.std():.4f}, fp32 noise band mse {band:.3e} -> {path} ({time.time() - t0:.0f}s)")


if __name__ == "__main__":
    main(sys.argv[1:] or DEFAULT)
