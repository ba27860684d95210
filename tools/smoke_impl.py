"""smoke(): ONE small invocation of the hot path on cuda:0, checked against the CPU oracle.
Runs the SD-style tiny UNet fixture (spatial transformer + cross attention, asymmetric W4A8,
sm_abit 16, split shortcut) through QuantModel.forward and two PLMS steps of the sampler."""
import torch


def run():
    from tests.test_oracle_golden import load_case, noise_band_mse, oracle_forward
    from tests.test_unet_gpu import build_qnn
    from qdiff_b200 import _lib, samplers
    dev = torch.device("cuda:0")
    g = load_case("sd_tiny_w4a8_sm16")
    qnn = build_qnn(g, dev)
    n0 = _lib.lib().qd_launch_count()
    out = qnn(g["x"].to(dev), g["t"].to(dev), g["context"].to(dev)).cpu()
    n1 = _lib.lib().qd_launch_count()
    ref = oracle_forward(g)
    mse = ((out.double() - ref.double()) ** 2).mean().item()
    band = noise_band_mse(g, ref)
    assert torch.isfinite(out).all()
    assert n1 > n0, "no kernels were launched through libqdiff_b200.so"
    assert mse <= max(2 * band, 1e-6), f"smoke parity failed: mse {mse:.3e} vs noise band {band:.3e}"
    sampler = samplers.PLMSSampler(qnn, samplers.Schedule("linear", 1000, 0.00085, 0.0120))
    x, _ = sampler.sample(S=2, batch_size=2, shape=(4, 16, 16), conditioning=g["context"].to(dev),
                          unconditional_guidance_scale=1.0, x_T=g["x"])
    assert torch.isfinite(x).all()
    print(f"smoke ok: UNet eps mse vs oracle {mse:.3e} (reference fp32 noise band {band:.3e}); "
          f"{n1 - n0} kernels launched via the C ABI; 2-step PLMS latent finite")


if __name__ == "__main__":
    run()
