#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): Stable Diffusion v1-4 UNet, W4A8
(sm_abit 16, split shortcut), 50-step PLMS with classifier-free guidance, 8 images per GPU.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
    python bench.py --impl reference ...                     (CPU arm: the oracle port of the reference path)

A "step" is ONE denoising step of the batch: one UNet evaluation at batch 16 (8 images x [uncond, cond])
through the CUDA engine + the fused sampler update.  50 PLMS steps cost 51 UNet evaluations
(ldm/models/diffusion/plms.py:222-227), so images/s = 8 * N / (51 * step time).

Prints ONE JSON line (rank 0).  Weights are seeded synthetic (no checkpoints offline), activation
quantizers come from tests/golden/calib_sd_v1.json (reference 'max' quick-init on one seeded batch).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "q-diffusion_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOAD = "sd_v1"
IMAGES_PER_GPU = 8
UNET_EVALS_PER_IMAGE_BATCH = 51          # 50 PLMS steps (first step calls the UNet twice)
GOP_PER_IMAGE_EVAL = 803.3               # 401.64 GMAC per UNet evaluation of one image
CFG_SCALE = 7.5


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(bf16_burst=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], hbm=p["hbm_gbs"],
                    source="MEASURED_PEAKS.json")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md recipe), sampled through NVML
    (same counters as the nvidia-smi query, but fast enough to get several samples inside a sub-second region)."""

    def __init__(self, index):
        self.rows, self.stop, self.index = [], False, index
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                     "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                     "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self.stop:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((sm, mx, [n for n, bit in names.items() if r & bit]))
                time.sleep(0.02)
        except Exception as e:   # NVML unavailable: fall back to nvidia-smi polling
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            while not self.stop:
                try:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.rows.append((int(out[0]), int(out[1]), [n for i, n in enumerate(names) if out[2 + i].strip().lower().startswith("active")]))
                except Exception:
                    pass
                time.sleep(0.05)

    def __enter__(self):
        self.th.start()
        time.sleep(0.05)
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted({n for r in self.rows for n in r[2]})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max((r[1] for r in self.rows), default=None),
                    reasons=reasons, samples=len(self.rows))


def host_threads(cap=32):
    """CPU threads this process may really use: affinity mask and cgroup quota, not os.cpu_count() (the GPU boxes
    report 128+ logical CPUs to a container that owns far fewer; 128 torch threads there ran the oracle 10-50x slower)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            quota, period = f.read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        try:
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:
                quota = int(f.read())
            with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
                period = int(f.read())
            if quota > 0:
                n = min(n, max(1, quota // period))
        except (OSError, ValueError):
            pass
    return max(1, min(n, cap))


def cpu_baseline(ckpt=None, max_seconds=40.0):
    """The reference's CPU path (oracle port: fake-quant fp32 torch, host threads) on a bounded sample:
    ONE UNet evaluation of ONE image of the same SD workload; images/s = 1 / (2 * 51 * t_eval)."""
    import contextlib
    from oracle import synth_cfg
    from qdiff_b200 import synth
    if ckpt is None:
        with contextlib.redirect_stdout(sys.stderr):
            _, ckpt = synth.full_ckpt(WORKLOAD)
    ckpt = {k: (v.float() if k.endswith(".alpha") else v.cpu()) for k, v in ckpt.items()}
    x, t, ctx = synth.calib_inputs(WORKLOAD, batch=1, seed=99)
    with torch.no_grad():
        t0 = time.time()
        synth_cfg.oracle_forward(WORKLOAD, ckpt, x, t, ctx)
        t_eval = time.time() - t0
        if t_eval < max_seconds / 3:   # one more for a steadier number if it is cheap enough
            t0 = time.time()
            synth_cfg.oracle_forward(WORKLOAD, ckpt, x, t, ctx)
            t_eval = min(t_eval, time.time() - t0)
    evals_per_image = 2 * UNET_EVALS_PER_IMAGE_BATCH
    return dict(value=1.0 / (evals_per_image * t_eval), unit="images/s", cores=torch.get_num_threads(), kind="port",
                sample=f"1 UNet evaluation at batch 1 ({t_eval:.2f} s) of the {evals_per_image} per image; "
                       "oracle port of the reference fake-quant path (the reference itself is not on the GPU box)",
                unet_eval_s=t_eval)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(host_threads())
    import contextlib
    from qdiff_b200 import synth
    with contextlib.redirect_stdout(sys.stderr):
        _, ckpt_ref = synth.full_ckpt(WORKLOAD)
    times = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_baseline(ckpt_ref, max_seconds=0.0)
        if i >= args.warmup:
            times.append(cb["unet_eval_s"])
        if sum(times) > 150:
            break
    t_eval = sum(times) / len(times)
    # one "step" of the GPU arm = 16 image-evaluations; the CPU sample is 1 -> scale to the same unit
    ms_per_step = t_eval * 2 * IMAGES_PER_GPU * 1e3
    value = IMAGES_PER_GPU / (UNET_EVALS_PER_IMAGE_BATCH * ms_per_step * 1e-3)
    cb.update(value=value, sample=f"{len(times)} x 1 UNet evaluation at batch 1, scaled x16 to one batch-16 step")
    print(json.dumps({
        "impl": "reference", "metric": "images_per_sec", "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32 fake-quant (reference simulation)", "data": "synthetic",
        "config": {"workload": "SD v1-4 UNet (860M) W4A8 asymmetric, sm_abit 16, split shortcut; PLMS-50 + CFG 7.5; "
                               "8 images/GPU -> UNet batch 16, 64x64x4 latents, 77x768 context",
                   "step": "1 denoising step = 1 UNet evaluation at batch 16 (CPU arm: one batch-1 evaluation timed, x16)",
                   "host_threads": torch.get_num_threads()},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def gemm_roofline(prog, pk):
    """Replay the recorded program op by op; CUDA events around every INT8 GEMM launch (default stream =
    the stream the engine launches on).  achieved = sum(2*M*N*K) / sum(duration)."""
    from qdiff_b200 import _lib
    names = prog.op_names
    gemm_ids = [i for i, n in enumerate(names) if prog.op_kinds[i] == _lib.QD_OP_GEMM]
    prog.run_range(0, prog.nops)  # warm
    torch.cuda.synchronize()
    evs = []
    for i in range(prog.nops):
        if i in set(gemm_ids):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            prog.run_range(i, i + 1)
            e1.record()
            evs.append((i, e0, e1))
        else:
            prog.run_range(i, i + 1)
    torch.cuda.synchronize()
    tot_ms = sum(e0.elapsed_time(e1) for _, e0, e1 in evs)
    tot_ops = sum(prog.op_flops[i] for i, _, _ in evs)
    achieved = tot_ops / (tot_ms * 1e-3) / 1e12
    peak = 2.0 * pk["bf16_sustained"]
    # DRAM bytes per GEMM launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over one step's launches) from the
    # committed ncu pass over this same command: tools/launch_summary.py --traffic writes the file
    traffic, tsrc = None, None
    tf = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_roofline_traffic.json")
    if os.path.exists(tf):
        with open(tf) as f:
            tj = json.load(f)
        traffic, tsrc = tj.get("gemm_dram_bytes_per_launch"), tj.get("source")
    return dict(bound="tensor", achieved=achieved, peak=peak, unit="TOP/s", frac=achieved / peak, traffic=traffic,
                traffic_source=tsrc,
                kernel="gemm_i8_kernel (tcgen05.mma kind::i8)", launches=len(evs), gemm_ms_per_step=tot_ms,
                peak_source=f"2 x bf16_tflops_sustained ({pk['source']}); INT8 dense = 2x bf16 on sm_100a",
                note="events bracket each launch individually (serialised, includes launch gaps)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels individually (for ncu launch lists)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: qdiff_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if dist is not None:
        dist.barrier()
    from qdiff_b200 import _lib, samplers, synth
    L = _lib.lib()

    import contextlib
    with contextlib.redirect_stdout(sys.stderr):   # the reference-compatible loaders print; stdout carries ONE JSON line
        qnn, ckpt = synth.build_qnn(WORKLOAD, cuda_graph=not args.no_graph)
    B = IMAGES_PER_GPU
    # every rank draws the FULL batch from the same seed and keeps its shard (N-rank == 1-rank results)
    from qdiff_b200 import dist as qdist
    x_sh, c_sh = qdist.shard_like_single_process((world * B, 4, 64, 64), 42, rank, world,
                                                 extra_shapes=[(world * B, 77, 768)])
    uc = torch.randn(1, 77, 768, generator=torch.Generator().manual_seed(43)).expand(B, 77, 768)
    x_host = x_sh.pin_memory()
    c_host = torch.cat([uc, c_sh]).contiguous().pin_memory()   # [uncond; cond] as plms.py:187
    sched = samplers.Schedule("linear", 1000, 0.00085, 0.0120)               # configs/stable-diffusion/v1-inference.yaml
    sampler = samplers.PLMSSampler(qnn, sched)
    sampler.make_schedule(50)
    ts = list(reversed(sampler.ddim_timesteps.tolist()))

    x = x_host.to(dev)
    ctx = c_host.to(dev)
    nxt = torch.empty_like(x)
    e_t = torch.empty_like(x)
    old = [torch.randn_like(x) for _ in range(3)]

    def step(i, x, nxt):
        """One PLMS step at multistep order 4 (the steady state: 47 of the 50 steps)."""
        idx = i % len(ts)
        t = torch.full((2 * B,), int(ts[idx]), device=dev, dtype=torch.long)
        eps = qnn(torch.cat([x, x]), t, ctx)
        k = len(ts) - 1 - idx
        samplers._step(x, eps, nxt, a_t=sampler.ddim_alphas[k], a_prev=sampler.ddim_alphas_prev[k], sigma=0.0,
                       sqrt_one_minus_at=sampler.ddim_sqrt_one_minus_alphas[k], cfg_scale=CFG_SCALE,
                       coef=samplers.PLMSSampler._AB[3], olds=(old[0], old[1], old[2]), eps_out=e_t)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i, x, nxt)
    barrier()
    launches0 = L.qd_launch_count()
    with ClockSampler(local) as clocks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.profiler.start()    # no-op unless run under `ncu --profile-from-start off` (profiles/: launch list)
        e0.record()
        for i in range(args.steps):
            step(args.warmup + i, x, nxt)
        e1.record()
        barrier()
        torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    launches = L.qd_launch_count() - launches0
    prog = qnn.program(torch.cat([x, x]), ctx)
    # with a CUDA graph the kernels replay without passing through the C ABI: count them from the program
    if qnn.use_cuda_graph:
        launches = args.steps * (prog.kernel_launches + 1)

    # ---- e2e: same step through the public API with HOST buffers (H2D of latents+context, D2H of x_prev)
    out_host = torch.empty_like(x_host)
    def e2e_step(i):
        xd = x_host.to(dev, non_blocking=True)
        cd = c_host.to(dev, non_blocking=True)
        idx = i % len(ts)
        t = torch.full((2 * B,), int(ts[idx]), device=dev, dtype=torch.long)
        eps = qnn(torch.cat([xd, xd]), t, cd)
        k = len(ts) - 1 - idx
        samplers._step(xd, eps, nxt, a_t=sampler.ddim_alphas[k], a_prev=sampler.ddim_alphas_prev[k], sigma=0.0,
                       sqrt_one_minus_at=sampler.ddim_sqrt_one_minus_alphas[k], cfg_scale=CFG_SCALE,
                       coef=samplers.PLMSSampler._AB[3], olds=(old[0], old[1], old[2]))
        out_host.copy_(nxt, non_blocking=True)
    for i in range(3):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for i in range(args.steps):
        e2e_step(i)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    times = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        qdist.gather_latents(nxt, world)   # the path's only collective: final latent gather
    ms, ms_e2e = float(times[0]), float(times[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    ms_step = ms / args.steps
    value = world * B / (UNET_EVALS_PER_IMAGE_BATCH * ms_step * 1e-3)
    e2e_value = world * B / (UNET_EVALS_PER_IMAGE_BATCH * (ms_e2e / args.steps) * 1e-3)
    pk = peaks()
    step_tops = 2 * B * GOP_PER_IMAGE_EVAL / 1e3 / (ms_step * 1e-3)
    line = {
        "metric": "images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8 (W4 codes x A8 codes, int32 accumulate; fp32 norms/softmax)",
        "data": "synthetic",
        "config": {"workload": "SD v1-4 UNet (860M) W4A8 asymmetric, sm_abit 16, split shortcut; PLMS-50 + CFG 7.5; "
                               "8 images/GPU -> UNet batch 16, 64x64x4 latents, 77x768 context",
                   "step": "1 denoising step = 1 UNet evaluation at batch 16 + fused sampler update",
                   "unet_step_ms": ms_step, "unet_evals_per_image_batch": UNET_EVALS_PER_IMAGE_BATCH,
                   "l2": "working set per step (weights 0.86 GB as s8 + >10 GB activations) is far larger than L2",
                   "cuda_graph": bool(qnn.use_cuda_graph), "engine_ops_per_step": prog.nops,
                   "whole_step_int8_tops": step_tops, "parallelism": f"dp{world} (batch sharded, no collective in the loop)"},
        "e2e": {"value": e2e_value, "unit": "images/s",
                "h2d_bytes_per_step": x_host.numel() * 4 + c_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
    }
    if not args.no_roofline:
        line["roofline"] = gemm_roofline(prog, pk)
        line["roofline"]["whole_step_frac"] = step_tops / line["roofline"]["peak"]
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(host_threads())
        cb = cpu_baseline(ckpt)
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
