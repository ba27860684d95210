"""Per-op timing of one recorded UNet program (CUDA events around every engine op, serialised).
usage: python tools/op_profile.py [workload] [batch]   -> gpurun_out/op_profile_<workload>.json + table"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "q-diffusion_b200")):
    sys.path.insert(0, p)
from qdiff_b200 import _lib, synth  # noqa: E402

KIND = {1: "gemm", 2: "quantize", 3: "groupnorm", 4: "layernorm", 5: "im2col", 6: "attention", 7: "temb", 8: "copy2d",
        9: "nchw2nhwc", 10: "nhwc2nchw", 11: "avgpool", 12: "upsample"}


def main(name="sd_v1", batch=16):
    dev = torch.device("cuda:0")
    qnn, _ = synth.build_qnn(name, cuda_graph=False)
    spec = synth.SPECS[name]
    x = torch.randn(batch, *spec["in_shape"], device=dev)
    t = torch.randint(0, 1000, (batch,), device=dev)
    ctx = torch.randn(batch, *spec["ctx"], device=dev) if spec["ctx"] else None
    qnn(x, t, ctx)
    prog = qnn.program(x, ctx)
    torch.cuda.synchronize()
    reps = 3
    times = [[] for _ in range(prog.nops)]
    for _ in range(reps):
        evs = []
        for i in range(prog.nops):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            prog.run_range(i, i + 1)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        for i, (e0, e1) in enumerate(evs):
            times[i].append(e0.elapsed_time(e1))
    rows = []
    for i in range(prog.nops):
        ms = sorted(times[i])[reps // 2]
        rows.append(dict(i=i, kind=KIND.get(prog.op_kinds[i], "?"), label=prog.op_names[i], ms=ms, gop=prog.op_flops[i] / 1e9))
    tot = sum(r["ms"] for r in rows)
    fam = {}
    for r in rows:
        fam[r["kind"]] = fam.get(r["kind"], 0.0) + r["ms"]
    print(f"{name} B={batch}: {prog.nops} ops, serialised sum {tot:.2f} ms")
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1]):
        print(f"  {k:10s} {v:8.3f} ms {100 * v / tot:5.1f}%  n={sum(1 for r in rows if r['kind'] == k)}")
    print("top ops:")
    for r in sorted(rows, key=lambda r: -r["ms"])[:45]:
        tops = f"{r['gop'] / r['ms']:8.1f} TOP/s" if r["gop"] else ""
        print(f"  {r['ms']:7.3f} ms  {r['kind']:9s} {r['label'][:70]:70s} {tops}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"op_profile_{name}.json"), "w"))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "sd_v1", int(sys.argv[2]) if len(sys.argv) > 2 else 16)
