"""First-stage decode timing on one GPU: ms per image, algorithmic TFLOP/s and executed bfloat16
tensor TFLOP/s (precision = plane products per MAC), per-kind share of the decode program.
usage: python tools/bench_decode.py [sd_v1|lsun_church|lsun_bedroom] [batch] [precision]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "q-diffusion_b200")):
    sys.path.insert(0, p)
from qdiff_b200 import first_stage as FS  # noqa: E402
from qdiff_b200.unet import randomize_  # noqa: E402

KIND = {1: "gemm", 3: "groupnorm", 5: "im2col", 8: "copy2d", 9: "nchw2nhwc", 10: "nhwc2nchw", 13: "split3", 14: "attention_fp", 15: "vq_lookup",
        16: "softmax_rows"}


def main(name="sd_v1", batch=2, precision=3):
    dev = torch.device("cuda:0")
    cfg = FS.CONFIGS[name]
    fs = randomize_(FS.build_first_stage(name, precision=precision), seed=3).to(dev)
    zc = cfg["ddconfig"]["z_channels"]
    res = 32 if name == "lsun_church" else 64
    z = torch.randn(batch, zc, res, res, device=dev)
    for _ in range(3):
        out = FS.decode_first_stage(fs, z, cfg["scale_factor"])
    torch.cuda.synchronize()
    prog = next(iter(fs._programs.values()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        out = FS.decode_first_stage(fs, z, cfg["scale_factor"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = float(sum(prog.op_flops))
    # per-kind serialised share (events around every op, eager replay)
    times = {}
    for i in range(prog.nops):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        prog.run_range(i, i + 1)
        b.record()
        torch.cuda.synchronize()
        k = KIND.get(prog.op_kinds[i], str(prog.op_kinds[i]))
        times[k] = times.get(k, 0.0) + a.elapsed_time(b)
    tot = sum(times.values())
    line = dict(workload=f"{name} first-stage decode {tuple(z.shape)} -> {tuple(out.shape)}", batch=batch, precision=precision,
                ms_per_batch=ms, ms_per_image=ms / batch, images_per_s=1000.0 * batch / ms, ops=prog.nops,
                algorithmic_tflop=flops / 1e12, algorithmic_tflops=flops / ms / 1e9,
                executed_bf16_tflops=precision * flops / ms / 1e9,
                share={k: round(v / tot, 3) for k, v in sorted(times.items(), key=lambda kv: -kv[1])})
    print(json.dumps(line))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"bench_decode_{name}_b{batch}_p{precision}.json"), "w") as f:
        json.dump(line, f)


if __name__ == "__main__":
    a = sys.argv[1:]
    main(a[0] if a else "sd_v1", int(a[1]) if len(a) > 1 else 2, int(a[2]) if len(a) > 2 else 3)
