"""Command-line surface of the three sampling scripts, flag for flag (names, types, defaults, choices):

    scripts/sample_diffusion_ddim.py   <- reference scripts/sample_diffusion_ddim.py:350-477   (CIFAR-10 DDIM)
    scripts/sample_diffusion_ldm.py    <- reference scripts/sample_diffusion_ldm.py:191-349    (unconditional LDM)
    scripts/txt2img.py                 <- reference scripts/txt2img.py:107-331                 (Stable Diffusion)

Same flags, same meaning; what runs underneath is the engine.  Scope: the denoising loop on a
calibrated checkpoint -- `--ptq --resume --cali_ckpt ckpt.pth` (the checkpoint the reference's calibration wrote; it
carries the FP weights, the AdaRound parameters and the activation quantizers, SURVEY Appendix C, so no base checkpoint
is needed).  Calibration itself (`--ptq` without `--resume`), `--resume_w` and the text encoder
are outside the hot path: those flags parse, and the run stops with a message naming what to do instead.

Extra flags of this implementation (all prefixed so they cannot collide with future reference flags):
    --b200_synthetic NAME     seeded synthetic weights + the committed calibration fixture (offline runs, no checkpoint)
    --b200_context FILE       txt2img: pre-computed prompt embeddings {"c": [B,77,768], "uc": [1|B,77,768]} (torch.save);
                              without it a seeded N(0,1) context is used, like the reference's own dummy calibration input
    --b200_out FILE           where to save the latents / images tensor (default: <logdir or outdir>/samples.pt)
    --b200_decode             LDM / txt2img: decode the final latents with the first stage ON THE ENGINE (qdiff_b200.first_stage;
                              ddpm.py:710-767) and save the images in [0, 1] next to the latents.  Weights: --b200_first_stage
                              FILE (a first-stage or full LDM / SD checkpoint: keys `decoder.*`, `post_quant_conv.*`,
                              `quantize.embedding.weight`, optionally prefixed `first_stage_model.`), or seeded synthetic
                              weights with --b200_synthetic.  --b200_decode_precision 1|3|6: bfloat16 plane products per MAC.
Multi-GPU: run under `python -m torch.distributed.run`; the batch is sharded by images, every rank draws the full-batch
noise from the seed and keeps its slice, rank 0 gathers and saves (qdiff_b200/dist.py).
"""
import argparse
import os
import time

# ---------------------------------------------------------------------------------------------- flag tables
# (flags, kwargs).  Help strings are this implementation's own wording.
_QUANT = [
    (("--ptq",), dict(action="store_true", help="run the post-training-quantised UNet (the engine's only mode)")),
    (("--quant_act",), dict(action="store_true", help="activations are quantised too (W?A? instead of weight-only)")),
    (("--weight_bit",), dict(type=int, default=8, help="weight bits")),
    (("--act_bit",), dict(type=int, default=8, help="activation bits")),
]
_CALI = [
    (("--cali_st",), dict(type=int, default=1, help="calibration: timesteps sampled")),
    (("--cali_batch_size",), dict(type=int, default=32, help="calibration: reconstruction batch size")),
    (("--cali_n",), dict(type=int, default=1024, help="calibration: samples per timestep")),
    (("--cali_iters",), dict(type=int, default=20000, help="calibration: weight reconstruction iterations")),
    (("--cali_iters_a",), dict(default=5000, type=int, help="calibration: activation (LSQ) iterations")),
    (("--cali_lr",), dict(default=4e-4, type=float, help="calibration: LSQ learning rate")),
    (("--cali_p",), dict(default=2.4, type=float, help="calibration: L_p norm")),
    (("--cali_ckpt",), dict(type=str, help="calibrated checkpoint (ckpt.pth) to resume from")),
    (("--cali_data_path",), dict(type=str, default="sd_coco_sample1024_allst.pt", help="calibration data file")),
    (("--resume",), dict(action="store_true", help="load quantizer parameters from --cali_ckpt and sample")),
    (("--resume_w",), dict(action="store_true", help="load only the weight quantizers, then calibrate activations")),
    (("--cond",), dict(action="store_true", help="conditional model (cross-attention context)")),
]
_TAIL = [
    (("--sm_abit",), dict(type=int, default=8, help="bits of the attention-softmax quantizer")),
    (("--verbose",), dict(action="store_true", help="print the wrapped model")),
]
_B200 = [
    (("--b200_synthetic",), dict(type=str, default=None, help="seeded synthetic workload (cifar10 | lsun_bedroom | lsun_church | sd_v1)")),
    (("--b200_out",), dict(type=str, default=None, help="output tensor file")),
    (("--b200_decode",), dict(action="store_true", help="decode the latents with the first stage on the engine (LDM / txt2img)")),
    (("--b200_first_stage",), dict(type=str, default=None, help="first-stage checkpoint for --b200_decode")),
    (("--b200_decode_precision",), dict(type=int, default=3, choices=[1, 3, 6], help="bfloat16 plane products per MAC of the decoder")),
]


def _add(parser, table):
    for flags, kw in table:
        parser.add_argument(*flags, **kw)


def ddim_parser():
    p = argparse.ArgumentParser(description="CIFAR-10 DDIM sampling on the qdiff_b200 engine")
    _add(p, [
        (("--config",), dict(type=str, required=True, help="model config (the reference's configs/cifar10.yml)")),
        (("--seed",), dict(type=int, default=1234, help="random seed")),
        (("-l", "--logdir"), dict(type=str, nargs="?", default="none", help="log directory")),
        (("--use_pretrained",), dict(action="store_true")),
        (("--sample_type",), dict(type=str, default="generalized", help="generalized | ddpm_noisy | dpm_solver")),
        (("--skip_type",), dict(type=str, default="uniform", help="uniform | quad")),
        (("--timesteps",), dict(type=int, default=1000, help="number of sampling steps")),
        (("--eta",), dict(type=float, default=0.0, help="DDIM eta")),
        (("--sequence",), dict(action="store_true")),
    ])
    _add(p, _QUANT)
    _add(p, [(("--quant_mode",), dict(type=str, default="qdiff", choices=["qdiff"], help="quantisation mode")),
             (("--max_images",), dict(type=int, default=50000, help="number of images to sample"))])
    _add(p, _CALI)
    _add(p, [(("--a_sym",), dict(action="store_true", help="symmetric activation quantizers")),
             (("--running_stat",), dict(action="store_true", help="calibration: running statistics"))])
    _add(p, [_TAIL[0], (("--split",), dict(action="store_true", help="split-shortcut quantisation")), _TAIL[1]])
    _add(p, _B200)
    return p


def ldm_parser():
    p = argparse.ArgumentParser(description="unconditional LDM sampling on the qdiff_b200 engine")
    _add(p, [
        (("-r", "--resume_base"), dict(type=str, nargs="?", help="base model logdir or checkpoint (its config.yaml is read)")),
        (("-n", "--n_samples"), dict(type=int, nargs="?", default=50000, help="samples to draw")),
        (("-e", "--eta"), dict(type=float, nargs="?", default=1.0, help="DDIM eta")),
        (("-v", "--vanilla_sample"), dict(default=False, action="store_true", help="ancestral DDPM sampling")),
        (("--seed",), dict(type=int, required=True, help="random seed")),
        (("-l", "--logdir"), dict(type=str, nargs="?", default="none", help="log directory")),
        (("-c", "--custom_steps"), dict(type=int, nargs="?", default=50, help="DDIM steps")),
        (("--batch_size",), dict(type=int, nargs="?", default=10, help="batch size")),
    ])
    _add(p, _QUANT)
    _add(p, [(("--quant_mode",), dict(type=str, default="qdiff", choices=["qdiff"], help="quantisation mode"))])
    _add(p, _CALI)
    _add(p, [(("--a_sym",), dict(action="store_true", help="symmetric activation quantizers")),
             (("--a_min_max",), dict(action="store_true", help="calibration: min-max activation init")),
             (("--running_stat",), dict(action="store_true", help="calibration: running statistics")),
             (("--rs_sm_only",), dict(action="store_true", help="calibration: running statistics for softmax only")),
             _TAIL[0],
             (("--dpm",), dict(action="store_true", help="DPM-Solver sampling")),
             _TAIL[1]])
    _add(p, _B200)
    return p


def txt2img_parser():
    p = argparse.ArgumentParser(description="Stable Diffusion txt2img latents on the qdiff_b200 engine")
    _add(p, [
        (("--prompt",), dict(type=str, nargs="?", default="a painting of a virus monster playing guitar", help="prompt")),
        (("--outdir",), dict(type=str, nargs="?", default="outputs/txt2img-samples", help="output directory")),
        (("--skip_grid",), dict(action="store_true")),
        (("--skip_save",), dict(action="store_true")),
        (("--ddim_steps",), dict(type=int, default=50, help="sampling steps")),
        (("--plms",), dict(action="store_true", help="PLMS sampler")),
        (("--laion400m",), dict(action="store_true")),
        (("--fixed_code",), dict(action="store_true", help="same start code for every batch")),
        (("--ddim_eta",), dict(type=float, default=0.0, help="DDIM eta")),
        (("--n_iter",), dict(type=int, default=2, help="batches per prompt")),
        (("--H",), dict(type=int, default=512)), (("--W",), dict(type=int, default=512)),
        (("--C",), dict(type=int, default=4)), (("--f",), dict(type=int, default=8)),
        (("--n_samples",), dict(type=int, default=3, help="batch size")),
        (("--n_rows",), dict(type=int, default=0)),
        (("--scale",), dict(type=float, default=7.5, help="classifier-free guidance scale")),
        (("--from-file",), dict(type=str, help="file with one prompt per line")),
        (("--config",), dict(type=str, default="configs/stable-diffusion/v1-inference.yaml", help="model config")),
        (("--ckpt",), dict(type=str, default="models/ldm/stable-diffusion-v1/model.ckpt", help="base checkpoint")),
        (("--seed",), dict(type=int, default=42, help="random seed")),
        (("--precision",), dict(type=str, choices=["full", "autocast"], default="autocast")),
    ])
    _add(p, _QUANT)
    # reference quirk Q5: the default is not among the choices, so --ptq needs --quant_mode qdiff
    _add(p, [(("--quant_mode",), dict(type=str, default="symmetric", choices=["linear", "squant", "qdiff"], help="quantisation mode"))])
    _add(p, _CALI)
    _add(p, [(("--no_grad_ckpt",), dict(action="store_true")),
             (("--split",), dict(action="store_true", help="split-shortcut quantisation")),
             (("--running_stat",), dict(action="store_true")), (("--rs_sm_only",), dict(action="store_true")),
             _TAIL[0], _TAIL[1]])
    _add(p, _B200)
    _add(p, [(("--b200_context",), dict(type=str, default=None, help="pre-computed prompt embeddings (see module docstring)"))])
    return p


def surface(parser):
    """{dest: {flags, default, type, nargs, choices, required, action}} -- compared with the reference by tests."""
    out = {}
    for a in parser._actions:
        if a.dest == "help":
            continue
        out[a.dest] = dict(flags=sorted(a.option_strings), default=a.default,
                           type=getattr(a.type, "__name__", None) if a.type is not None else None,
                           nargs=None if isinstance(a, argparse._StoreTrueAction) else a.nargs,
                           choices=list(a.choices) if a.choices is not None else None, required=bool(a.required),
                           action=type(a).__name__)
    return out


# ---------------------------------------------------------------------------------------------- shared run helpers
def _dist_env():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    return rank, world, local


def _setup(seed):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("qdiff_b200 scripts need a CUDA device (sm_100a): the engine has no CPU fallback")
    rank, world, local = _dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    return rank, world, torch.device("cuda", local)


def _require_resume(args):
    if not args.ptq:
        raise SystemExit("the engine realises the quantised UNet only: pass --ptq (full-precision sampling is the "
                         "reference's own path)")
    if getattr(args, "quant_mode", "qdiff") != "qdiff":
        raise SystemExit("--quant_mode qdiff is the only mode (the reference's txt2img silently skips PTQ otherwise, "
                         "SURVEY Appendix D Q5)")
    if args.b200_synthetic:
        return
    if args.resume_w or not args.resume:
        raise SystemExit("calibration is not part of the sampling hot path (SURVEY section 8 f4): calibrate with the "
                         "reference, then run with --resume --cali_ckpt <ckpt.pth>")
    if not args.cali_ckpt or not os.path.exists(args.cali_ckpt):
        raise SystemExit(f"--cali_ckpt {args.cali_ckpt!r} not found")


def _wrap(model, args, a_sym, device):
    import qdiff_b200 as qd
    wq = {'n_bits': args.weight_bit, 'channel_wise': True, 'scale_method': 'max'}
    aq = {'n_bits': args.act_bit, 'symmetric': a_sym, 'channel_wise': False, 'scale_method': 'max',
          'leaf_param': args.quant_act}
    qnn = qd.QuantModel(model=model, weight_quant_params=wq, act_quant_params=aq, sm_abit=args.sm_abit)
    qd.resume_cali_model(qnn, args.cali_ckpt, None, args.quant_act, "qdiff", cond=bool(getattr(args, "cond", False)))
    if args.verbose:
        print(qnn)
    return qnn


def _synthetic(args, expect_family):
    from . import synth
    name = args.b200_synthetic
    if name not in synth.SPECS or synth.SPECS[name]["family"] != expect_family:
        raise SystemExit(f"--b200_synthetic {name!r}: expected one of "
                         f"{[k for k, v in synth.SPECS.items() if v['family'] == expect_family]}")
    qnn, _ = synth.build_qnn(name)
    return qnn, synth.SPECS[name]


def _load_yaml(path):
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)


def _save(args, default_dir, tensor, meta, rank):
    import torch
    if rank != 0:
        return None
    path = args.b200_out or os.path.join(default_dir, "samples.pt")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(dict(samples=tensor.cpu(), **meta), path)
    print(f"saved {tuple(tensor.shape)} -> {path}")
    return path


def _first_stage(args, name, cfg_params, dev):
    """(container on `dev`, scale_factor) for --b200_decode, or (None, 1.0).  name: synthetic workload; cfg_params: the
    `model.params` block of an LDM / SD yaml (first_stage_config, scale_factor)."""
    import torch
    from . import first_stage as FS
    from .unet import randomize_
    if not args.b200_decode:
        return None, 1.0
    if name is not None:
        if name not in FS.CONFIGS:
            raise SystemExit(f"--b200_decode: no first stage is defined for the synthetic workload {name!r}")
        cfg = FS.CONFIGS[name]
    else:
        fsc = cfg_params["first_stage_config"]
        p = fsc["params"]
        kind = "vq" if fsc["target"].endswith("VQModelInterface") else "kl" if fsc["target"].endswith("AutoencoderKL") else None
        if kind is None:
            raise SystemExit(f"--b200_decode: first stage {fsc['target']} is not AutoencoderKL / VQModelInterface")
        cfg = dict(kind=kind, embed_dim=p["embed_dim"], n_embed=p.get("n_embed"), ddconfig=dict(p["ddconfig"]),
                   scale_factor=float(cfg_params.get("scale_factor", 1.0)))
    fs = FS.build_first_stage(cfg, precision=args.b200_decode_precision)
    if args.b200_first_stage:
        sd = torch.load(args.b200_first_stage, map_location="cpu", weights_only=False)
        sd = sd.get("state_dict", sd)
        sd = {(k[len("first_stage_model."):] if k.startswith("first_stage_model.") else k): v for k, v in sd.items()}
        fs.load_state_dict(sd, strict=True)
    elif name is not None:
        randomize_(fs, seed=7)
    else:
        raise SystemExit("--b200_decode needs --b200_first_stage CKPT (or a --b200_synthetic workload)")
    return fs.to(dev), cfg["scale_factor"]


def _decode_images(fs, scale_factor, z, dev, chunk=4):
    """decode_first_stage in chunks, then the scripts' image range: clamp((x + 1) / 2, 0, 1) (txt2img.py:527)."""
    import torch
    from .first_stage import decode_first_stage
    outs = []
    for i in range(0, z.shape[0], chunk):
        x = decode_first_stage(fs, z[i:i + chunk].to(dev), scale_factor)
        outs.append(torch.clamp((x + 1.0) / 2.0, min=0.0, max=1.0).cpu())
    return torch.cat(outs)


def _shard(n_total, world):
    if n_total % world:
        raise SystemExit(f"batch size {n_total} is not divisible by the {world} ranks")
    return n_total // world


# ---------------------------------------------------------------------------------------------- sample_diffusion_ddim
def run_ddim(args):
    """Diffusion.sample / sample_fid / sample_image of the reference script (:110-347) on the engine."""
    import numpy as np
    import torch
    from . import dist as qdist, samplers, unet
    _require_resume(args)
    rank, world, dev = _setup(args.seed)
    if args.cond:
        raise SystemExit("--cond is not valid for the DDIM (CIFAR) script (the reference asserts the same)")
    cfg = _load_yaml(args.config) if os.path.exists(args.config) else None
    if args.b200_synthetic:
        qnn, spec = _synthetic(args, "ddim")
        ch, size = spec["in_shape"][0], spec["in_shape"][1]
        d = dict(beta_start=0.0001, beta_end=0.02, num_diffusion_timesteps=1000)
        batch = (cfg or {}).get("sampling", {}).get("batch_size", 64)
    else:
        if cfg is None:
            raise SystemExit(f"--config {args.config!r} not found")
        m, d = cfg["model"], cfg["diffusion"]
        ch, size = cfg["data"]["channels"], cfg["data"]["image_size"]
        model = unet.Model(unet.ddim_config(ch=m["ch"], out_ch=m["out_ch"], ch_mult=m["ch_mult"],
                                            num_res_blocks=m["num_res_blocks"], attn_resolutions=m["attn_resolutions"],
                                            in_channels=m["in_channels"], image_size=size,
                                            resamp_with_conv=m.get("resamp_with_conv", True), split_shortcut=args.split,
                                            num_diffusion_timesteps=d["num_diffusion_timesteps"]))
        qnn = _wrap(model, args, args.a_sym, dev)
        batch = cfg["sampling"]["batch_size"]
    if d.get("beta_schedule", "linear") != "linear":
        raise SystemExit("only the linear beta schedule of the reference's configs is supported")
    T = d["num_diffusion_timesteps"]
    betas = torch.from_numpy(np.linspace(d["beta_start"], d["beta_end"], T, dtype=np.float64)).float()
    if args.sample_type != "generalized":
        raise SystemExit(f"--sample_type {args.sample_type}: the engine implements 'generalized' (DDIM); DPM-Solver / "
                         "ddpm_noisy are listed under SURVEY section 8 f3")
    if args.skip_type == "uniform":
        seq = list(range(0, T, T // args.timesteps))
    elif args.skip_type == "quad":
        seq = [int(s) for s in list(np.linspace(0, np.sqrt(T * 0.8), args.timesteps) ** 2)]
    else:
        raise NotImplementedError(args.skip_type)
    per = _shard(batch, world)
    n_rounds = max(1, -(-args.max_images // batch))
    outs, t0 = [], time.time()
    for r in range(n_rounds):
        (x,) = qdist.shard_like_single_process((batch, ch, size, size), args.seed + r, rank, world)
        noise_gen = torch.Generator().manual_seed(args.seed + 7919 * (r + 1))
        full_noise = [torch.randn(batch, ch, size, size, generator=noise_gen) for _ in seq] if args.eta > 0 else None
        lo = rank * per
        x = samplers.generalized_steps(x.to(dev), seq, lambda xx, tt: qnn(xx, tt), betas, eta=args.eta,
                                       noise_fn=(lambda k, shape, d_: full_noise[k][lo:lo + per].to(d_)) if full_noise else None)
        x = qdist.gather_latents(x, world)
        outs.append(torch.clamp((x + 1.0) / 2.0, 0.0, 1.0))       # inverse_data_transform (rescaled data)
    imgs = torch.cat(outs)[:args.max_images]
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        print(f"{imgs.shape[0]} images, {len(seq)} steps each, {dt:.2f} s -> {imgs.shape[0] / dt:.2f} images/s on {world} GPU(s)")
    return _save(args, args.logdir if args.logdir != "none" else ".", imgs, dict(kind="images", steps=len(seq)), rank)


# ---------------------------------------------------------------------------------------------- sample_diffusion_ldm
def _ldm_config(args):
    """The reference resolves `<logdir>/config.yaml` next to the base checkpoint (sample_diffusion_ldm.py:379-410)."""
    base = args.resume_base
    if base is None:
        raise SystemExit("-r/--resume_base (logdir or checkpoint of the base model: its config.yaml is read) is required")
    logdir = base if os.path.isdir(base) else os.path.dirname(os.path.dirname(base)) or "."
    for cand in (os.path.join(logdir, "config.yaml"), os.path.join(os.path.dirname(base), "config.yaml")):
        if os.path.exists(cand):
            return _load_yaml(cand)
    raise SystemExit(f"no config.yaml found for {base!r}")


def run_ldm(args):
    """run / make_convolutional_sample / convsample_ddim of the reference script (:85-163) on the engine.  Saves the
    LATENTS; with --b200_decode also the images, decoded by the first stage on the engine (qdiff_b200.first_stage)."""
    import torch
    from . import dist as qdist, samplers, unet
    _require_resume(args)
    rank, world, dev = _setup(args.seed)
    if args.vanilla_sample:
        raise SystemExit("the engine implements DDIM (default) and DPM-Solver++ (--dpm) sampling for this script, not the "
                         "1000-step ancestral DDPM loop")
    if args.b200_synthetic:
        qnn, spec = _synthetic(args, "ldm")
        ch, size = spec["in_shape"][0], spec["in_shape"][1]
        sched = dict(timesteps=1000, linear_start=0.0015, linear_end=0.0195)
    else:
        cfg = _ldm_config(args)["model"]["params"]
        up = dict(cfg["unet_config"]["params"])
        model = unet.UNetModel(**up)
        qnn = _wrap(model, args, args.a_sym, dev)
        ch, size = cfg["channels"], cfg["image_size"]
        sched = dict(timesteps=cfg.get("timesteps", 1000), linear_start=cfg.get("linear_start", 1e-4),
                     linear_end=cfg.get("linear_end", 2e-2))
    schedule = samplers.Schedule("linear", sched["timesteps"], sched["linear_start"], sched["linear_end"])
    sampler = (samplers.DPMSolverSampler if args.dpm else samplers.DDIMSampler)(qnn, schedule)
    per = _shard(args.batch_size, world)
    outs, t0, r = [], time.time(), 0
    while sum(o.shape[0] for o in outs) < args.n_samples:
        (x_T,) = qdist.shard_like_single_process((args.batch_size, ch, size, size), args.seed + r, rank, world)
        gen = torch.Generator().manual_seed(args.seed + 7919 * (r + 1))
        lo = rank * per

        def noise_fn(i, shape, d_, gen=gen, lo=lo):
            return torch.randn(args.batch_size, ch, size, size, generator=gen)[lo:lo + per].to(d_)
        if args.dpm:      # convsample_dpm (sample_diffusion_ldm.py:96-103): deterministic, eta is not used
            z, _ = sampler.sample(S=args.custom_steps, batch_size=per, shape=(ch, size, size), x_T=x_T)
        else:
            z, _ = sampler.sample(S=args.custom_steps, batch_size=per, shape=(ch, size, size), eta=args.eta, x_T=x_T,
                                  noise_fn=noise_fn)
        outs.append(qdist.gather_latents(z, world))
        r += 1
    z = torch.cat(outs)[:args.n_samples]
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        print(f"{z.shape[0]} latents, {args.custom_steps} DDIM steps (eta {args.eta}), {dt:.2f} s -> {z.shape[0] / dt:.2f} /s on {world} GPU(s)")
    meta = dict(kind="latents", steps=args.custom_steps, eta=args.eta)
    if args.b200_decode and rank == 0:
        fs, sf = _first_stage(args, args.b200_synthetic, None if args.b200_synthetic else cfg, dev)
        t1 = time.time()
        meta["images"] = _decode_images(fs, sf, z, dev)
        torch.cuda.synchronize()
        print(f"decoded {tuple(meta['images'].shape)} in {time.time() - t1:.2f} s (first stage on the engine, precision {args.b200_decode_precision})")
    return _save(args, args.logdir if args.logdir != "none" else ".", z, meta, rank)


# ---------------------------------------------------------------------------------------------- txt2img
def run_txt2img(args):
    """The sampling loop of the reference's main() (:505-541) on the engine: PLMS / DDIM with classifier-free guidance.
    Prompt embeddings come from --b200_context (the CLIP text encoder is outside the scope); saves the latents,
    with --b200_decode also the decoded images."""
    import torch
    from . import dist as qdist, samplers, unet
    _require_resume(args)
    if not args.cond:
        raise SystemExit("txt2img needs --cond (the reference asserts the same)")
    rank, world, dev = _setup(args.seed)
    if args.precision == "autocast" and rank == 0:
        print("note: the engine computes the integer form of the fp32 (--precision full) path; autocast only affects "
              "the reference's fp16 simulation (SURVEY Appendix A.6)")
    if args.b200_synthetic:
        qnn, spec = _synthetic(args, "ldm")
        sched = dict(timesteps=1000, linear_start=0.00085, linear_end=0.0120)
        ctx_shape = spec["ctx"]
    else:
        cfg = _load_yaml(args.config)["model"]["params"]
        model = unet.UNetModel(**dict(cfg["unet_config"]["params"]))
        model.split = bool(args.split)
        qnn = _wrap(model, args, False, dev)
        sched = dict(timesteps=cfg.get("timesteps", 1000), linear_start=cfg["linear_start"], linear_end=cfg["linear_end"])
        ctx_shape = (77, cfg["unet_config"]["params"]["context_dim"])
    Sampler = samplers.PLMSSampler if args.plms else samplers.DDIMSampler
    sampler = Sampler(qnn, samplers.Schedule("linear", sched["timesteps"], sched["linear_start"], sched["linear_end"]))
    B = args.n_samples
    per = _shard(B, world)
    lo = rank * per
    if args.b200_context:
        emb = torch.load(args.b200_context, map_location="cpu")
        c_full, uc_full = emb["c"].float(), emb.get("uc")
        if c_full.shape[0] == 1:
            c_full = c_full.expand(B, -1, -1)
    else:
        g = torch.Generator().manual_seed(args.seed + 1)
        c_full = torch.randn(B, *ctx_shape, generator=g)
        uc_full = torch.randn(1, *ctx_shape, generator=g)
    c = c_full[lo:lo + per].contiguous().to(dev)
    uc = None
    if args.scale != 1.0:
        if uc_full is None:
            raise SystemExit("--scale != 1 needs the empty-prompt embedding 'uc' in --b200_context")
        uc = uc_full.float().expand(B, -1, -1)[lo:lo + per].contiguous().to(dev)
    shape = (args.C, args.H // args.f, args.W // args.f)
    start = None
    if args.fixed_code:
        (start,) = qdist.shard_like_single_process((B,) + shape, args.seed, rank, world)
    outs, t0 = [], time.time()
    for n in range(args.n_iter):
        x_T = start
        if x_T is None:
            (x_T,) = qdist.shard_like_single_process((B,) + shape, args.seed + 1 + n, rank, world)
        z, _ = sampler.sample(S=args.ddim_steps, conditioning=c, batch_size=per, shape=shape, verbose=False,
                              unconditional_guidance_scale=args.scale, unconditional_conditioning=uc, eta=args.ddim_eta,
                              x_T=x_T)
        outs.append(qdist.gather_latents(z, world))
    z = torch.cat(outs)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rank == 0:
        print(f"{z.shape[0]} latents {tuple(z.shape[1:])}, {args.ddim_steps} {'PLMS' if args.plms else 'DDIM'} steps, "
              f"scale {args.scale}: {dt:.2f} s -> {z.shape[0] / dt:.3f} images/s on {world} GPU(s)")
    meta = dict(kind="latents", steps=args.ddim_steps, scale=args.scale, prompt=args.prompt)
    if args.b200_decode and rank == 0:
        fs, sf = _first_stage(args, args.b200_synthetic, None if args.b200_synthetic else cfg, dev)
        t1 = time.time()
        meta["images"] = _decode_images(fs, sf, z, dev, chunk=2)
        torch.cuda.synchronize()
        print(f"decoded {tuple(meta['images'].shape)} in {time.time() - t1:.2f} s (first stage on the engine, precision {args.b200_decode_precision})")
    return _save(args, args.outdir, z, meta, rank)
