"""The three sampling scripts end to end on a GPU, with the reference's flags and the offline synthetic workloads
(--b200_synthetic: seeded weights + committed calibration fixture; there are no checkpoints on the GPU box), and from a
reference-format ckpt.pth written on the fly (--resume --cali_ckpt: the path a user of the reference takes)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    r = subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + "\n" + r.stderr[-4000:]
    return r.stdout + r.stderr


def test_txt2img_plms_synthetic(cuda, tmp_path):
    out = str(tmp_path / "z.pt")
    log = _run(["scripts/txt2img.py", "--plms", "--cond", "--ptq", "--quant_mode", "qdiff", "--quant_act", "--weight_bit", "4",
                "--act_bit", "8", "--sm_abit", "16", "--split", "--n_samples", "2", "--n_iter", "1", "--ddim_steps", "4",
                "--b200_synthetic", "sd_v1", "--b200_out", out])
    z = torch.load(out)["samples"]
    assert z.shape == (2, 4, 64, 64) and torch.isfinite(z).all(), log[-500:]


def test_sample_diffusion_ldm_ddim_and_dpm_synthetic(cuda, tmp_path):
    for extra in ([], ["--dpm"]):
        out = str(tmp_path / f"z{len(extra)}.pt")
        _run(["scripts/sample_diffusion_ldm.py", "--seed", "41", "-c", "4", "-e", "1.0", "--batch_size", "2", "-n", "2", "--ptq",
              "--quant_act", "--weight_bit", "8", "--b200_synthetic", "lsun_church", "--b200_out", out] + extra)
        z = torch.load(out)["samples"]
        assert z.shape == (2, 4, 32, 32) and torch.isfinite(z).all()


def test_sample_diffusion_ldm_decodes_images_on_the_engine(cuda, tmp_path):
    """--b200_decode: the first stage (kl-f8 decoder, seeded weights) runs on the engine after the loop: 32x32x4 latents ->
    256x256x3 images in [0, 1], saved next to the latents."""
    out = str(tmp_path / "img.pt")
    log = _run(["scripts/sample_diffusion_ldm.py", "--seed", "41", "-c", "4", "-e", "0.0", "--batch_size", "2", "-n", "2", "--ptq",
                "--quant_act", "--weight_bit", "8", "--b200_synthetic", "lsun_church", "--b200_decode", "--b200_out", out])
    blob = torch.load(out)
    img = blob["images"]
    assert blob["samples"].shape == (2, 4, 32, 32)
    assert img.shape == (2, 3, 256, 256) and torch.isfinite(img).all(), log[-500:]
    assert float(img.min()) >= 0.0 and float(img.max()) <= 1.0 and float(img.std()) > 1e-3


def test_sample_diffusion_ddim_from_reference_format_checkpoint(cuda, tmp_path):
    """--resume --cali_ckpt with a ckpt.pth in the reference's key format (the golden DDIM fixture's checkpoint) and a
    cifar10.yml-style config written next to it: the route a user of the reference takes."""
    import yaml
    from tests.test_oracle_golden import load_case
    g = load_case("ddim_w4a8_split")
    p = g["params"]
    cfg = dict(data=dict(dataset="CIFAR10", image_size=p["resolution"], channels=p["in_channels"]),
               model=dict(type="simple", in_channels=p["in_channels"], out_ch=p["out_ch"], ch=p["ch"], ch_mult=list(p["ch_mult"]),
                          num_res_blocks=p["num_res_blocks"], attn_resolutions=list(p["attn_resolutions"]), dropout=0.1,
                          resamp_with_conv=True),
               diffusion=dict(beta_schedule="linear", beta_start=0.0001, beta_end=0.02, num_diffusion_timesteps=1000),
               sampling=dict(batch_size=4, last_only=True))
    cfg_path, ckpt_path, out = str(tmp_path / "cfg.yml"), str(tmp_path / "ckpt.pth"), str(tmp_path / "img.pt")
    yaml.safe_dump(cfg, open(cfg_path, "w"))
    torch.save(g["ckpt"], ckpt_path)
    _run(["scripts/sample_diffusion_ddim.py", "--config", cfg_path, "--timesteps", "5", "--eta", "0", "--skip_type", "quad", "--ptq",
          "--weight_bit", "4", "--quant_mode", "qdiff", "--split", "--quant_act", "--act_bit", "8", "--a_sym", "--resume",
          "--cali_ckpt", ckpt_path, "--max_images", "4", "--b200_out", out])
    img = torch.load(out)["samples"]
    assert img.shape == (4, p["in_channels"], p["resolution"], p["resolution"])
    assert torch.isfinite(img).all() and float(img.min()) >= 0.0 and float(img.max()) <= 1.0
