"""EXPERIMENTAL: build and run the CTA-pair INT8 GEMM probe
(q-diffusion_b200/csrc/experimental/gemm_i8_2cta_probe.cu, tcgen05.mma.cta_group::2) - correctness against torch, then
timing on the operand-delivery-bound conv shapes next to the shipped 1-CTA kernel.  Not part of the product library.

    python tools/probe_2cta.py            # on a B200 (gpurun)
    python tools/probe_2cta.py --build    # cross-compile only (no GPU needed)
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "q-diffusion_b200", "csrc", "experimental", "gemm_i8_2cta_probe.cu")
LIB = os.path.join(ROOT, "q-diffusion_b200", "csrc", "experimental", "libprobe2cta.so")


def build():
    if not os.path.exists(LIB) or os.path.getmtime(SRC) > os.path.getmtime(LIB):
        subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
                        "-Xcompiler", "-fPIC", "-o", LIB, SRC, "-lcudart"], check=True)
    return LIB


def main():
    build()
    if "--build" in sys.argv:
        print("built", LIB)
        return
    import torch
    sys.path.insert(0, os.path.join(ROOT, "q-diffusion_b200"))
    from qdiff_b200 import ops
    L = C.CDLL(LIB)
    L.probe_gemm_i8_2cta.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(C.c_float)]
    dev = torch.device("cuda:0")

    def run(M, N, K, BN, iters):
        g = torch.Generator(device=dev).manual_seed(M + N + K)
        a = torch.randint(0, 256, (M, K), dtype=torch.uint8, device=dev, generator=g)
        b = torch.randint(-7, 8, (N, K), dtype=torch.int8, device=dev, generator=g)
        out = torch.full((M, N), -1, dtype=torch.int32, device=dev)
        ms = C.c_float(0)
        rc = L.probe_gemm_i8_2cta(a.data_ptr(), b.data_ptr(), out.data_ptr(), M, N, K, BN, iters, C.byref(ms))
        torch.cuda.synchronize()
        return rc, a, b, out, ms.value

    # ---- correctness (exact integers) on a small and a multi-tile shape
    for (M, N, K, BN) in [(256, 64, 128, 64), (512, 320, 384, 160), (2048, 512, 1280, 256)]:
        rc, a, b, out, _ = run(M, N, K, BN, 0)
        assert rc == 0, rc
        ref = (a.double() @ b.double().t()).to(torch.int64)
        ok = torch.equal(out.to(torch.int64), ref)
        print(f"probe {M}x{N}x{K} BN={BN}: {'exact' if ok else 'MISMATCH'}", flush=True)
        if not ok:
            bad = (out.to(torch.int64) != ref).nonzero()
            print("  first mismatches (row, col):", bad[:8].tolist())
            return 1
    # ---- timing: the 64x64 conv shapes as plain GEMMs with K = 9*C (rounded to 128), next to the shipped kernel
    for (M, N, K, BN) in [(65536, 320, 8704, 160), (65536, 320, 2944, 160), (16384, 640, 5760, 160), (16384, 1280, 11520, 256), (4096, 1280, 11520, 256)]:
        rc, a, b, out, ms = run(M, N, K, BN, 10)
        assert rc == 0, rc
        tops2 = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        scale = torch.ones(N, device=dev)
        o1 = torch.empty(M, N, device=dev)
        d = ops.gemm_desc(a, b, scale, M=M, N=N, C=K, a_signed=False, out=o1, ldo=N)
        for _ in range(2):
            ops.qgemm(d)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.qgemm(d)
        e1.record()
        torch.cuda.synchronize()
        ms1 = e0.elapsed_time(e1) / 10
        print(f"M={M} N={N} K={K}: pair kernel (BN={BN}) {ms * 1e3:7.1f} us {tops2:7.1f} TOP/s | shipped 1-CTA kernel "
              f"{ms1 * 1e3:7.1f} us {2.0 * M * N * K / (ms1 * 1e-3) / 1e12:7.1f} TOP/s", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main() or 0)
