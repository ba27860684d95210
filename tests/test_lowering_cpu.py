"""Host-side graph builders without a GPU: tools/dryrun_lowering.py replaces the C library by a recorder (every entry point
succeeds, tensors on the CPU) and lowers the first stage, the weight-only / full-precision UNet states and the INT8 DDIM
family.  Run in a subprocess (it patches module globals).  Nothing is computed: this pins the STRUCTURE of the recorded
programs - op counts, the copy-free decoder concat - so a lowering mistake shows up in the CPU suite already."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lowerings_dry_run():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dryrun_lowering.py")], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    out = r.stdout
    ops = {m.group(1).strip(): int(m.group(2)) for m in re.finditer(r"^(.*?): (\d+) ops", out, re.M)}
    # first stage: one conv = 1 / 2 / 3 accumulating launches with the precision; the VQ first stage adds the codebook lookup
    assert ops["first stage sd_v1 precision 1"] < ops["first stage sd_v1 precision 3"] < ops["first stage sd_v1 precision 6"]
    assert ops["first stage lsun_bedroom precision 3"] > 100
    # full-precision state: three weight planes -> more launches than the weight-only state of the same model
    for name in ("sd_tiny_w4_weightonly", "ldm_updown_w8_weightonly", "ldm_legacy_w4_weightonly", "ddim_w8_weightonly"):
        assert ops[f"{name} state (True, False)"] < ops[f"{name} state (False, False)"]
    # INT8 DDIM decoder: the concat costs no copies (and the copy form is still available as an A/B switch)
    m = re.search(r"QDIFF_DDIM_CAT=inplace: (\d+) ops, (\d+) copy2d", out)
    c = re.search(r"QDIFF_DDIM_CAT=copy: (\d+) ops, (\d+) copy2d", out)
    assert m and c and int(m.group(2)) == 0 and int(c.group(2)) == 8 and int(c.group(1)) - int(m.group(1)) == 8


def test_implicit_conv_rule():
    """Builder.implicit_conv_ok mirrors plan_gemm's tile rule (csrc/engine.cu): whole rows, whole small images, or 128-pixel
    segments of wide rows; everything else takes the explicit patch gather."""
    sys.path.insert(0, os.path.join(ROOT, "q-diffusion_b200"))
    from qdiff_b200.graph import Builder
    ok = Builder.implicit_conv_ok
    assert all(ok(h, w) for h, w in ((64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (128, 128), (256, 256), (512, 512), (2, 64)))
    assert not any(ok(h, w) for h, w in ((96, 96), (24, 24), (12, 12), (40, 40), (20, 20), (64, 48), (3, 64), (6, 6), (192, 192)))
