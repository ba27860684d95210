"""In-situ per-op parity (test infrastructure): replay a compiled engine program ONE OP AT A TIME and check every
op against the CPU oracle evaluated on the engine's OWN inputs of that op.

This is the deterministic counterpart of the noise-band gate in test_unet_gpu.py.  The fake-quant UNet amplifies
ulp-level differences, so comparing whole-network outputs can only be statistical; here every recorded op (each
QuantModule GEMM incl. split halves and W8 parts, every GroupNorm/LayerNorm/quantize kernel, every attention core,
im2col, layout ops) is fed exactly what the engine fed it and must reproduce the oracle's result for THAT input:
  * integer ops (im2col, plain quantizer of an fp32 tensor, copies): bit-exact;
  * INT8 GEMM with fp32 output: |err| <= 3e-6 x (|acc*scale| + |bias| + |rowvec| + |residual|)  (fp32 rounding only);
  * ops that emit codes behind fp32 arithmetic (norms, SiLU/GELU, requantising GEMM epilogues, attention with a fused
    consumer quantizer): <= 1 code at < 2e-3 of the positions;
  * attention with fp32 output: max |err| < 2e-3 x |ref|max (a few 1-step flips of P codes).
A systematic one-code bias in any layer fails these checks; ulp-level noise does not.  Together with the fold check
(folded integer weights x step == the oracle's fake-quant weights, bit-exact) this ties every QuantModule of every
golden case - ddim family included - to the oracle, which is itself pinned to the reference (test_oracle_golden.py).

Reference semantics restated by the checks: qdiff/quant_layer.py:82-89,248-279 (quantizer, QuantModule),
qdiff/quant_block.py:83-111,190-221,307-386 (blocks, attention), ldm util.py:151-171 / ddim diffusion.py:6-24
(timestep embedding).
"""
import torch
import torch.nn.functional as F

from oracle import ops_oracle as O

CODE_FRAC = 2e-3        # fraction of positions allowed to differ by one code behind fp32 arithmetic


# ----------------------------------------------------------------------------------------------- readers
def rd_f32(act):
    return act.logical().detach().to("cpu", torch.float64)


def rd_codes(act):
    t = act.logical().detach().cpu()
    if getattr(act, "f16", False):         # attention Q/K operands as fp16 (code - zero_point): back to codes
        return t.to(torch.float64).round().to(torch.int64) + int(act.zp[0])
    return t.to(torch.int64)               # uint8 -> 0..255, int8 -> -128..127


def vt_positions(T):
    """Token -> byte position inside the transposed V^T rows (attention.cuh att_vt_perm)."""
    t = torch.arange(T)
    return (t & ~15) | (((t >> 1) & 3) << 2) | (((t >> 3) & 1) << 1) | (t & 1)


def quant(y, q):
    """UniformAffineQuantizer codes (quant_layer.py:82-87) of a double tensor; q = (delta, zp, lo, hi).
    The division is carried out in fp32 like the reference (y is fp32-representable or is rounded to fp32 first)."""
    delta, zp, lo, hi = q
    yf = y.to(torch.float32)
    return torch.clamp(torch.round(yf / torch.tensor(delta, dtype=torch.float32)) + zp, lo, hi).to(torch.int64)


class Report:
    def __init__(self):
        self.rows = []

    def add(self, idx, label, kind, what, n, nbad, maxdiff, ok, note=""):
        self.rows.append(dict(idx=idx, label=label, kind=kind, what=what, n=int(n), nbad=int(nbad), maxdiff=float(maxdiff),
                              ok=bool(ok), note=note))

    def failures(self):
        return [r for r in self.rows if not r["ok"]]

    def text(self, only_interesting=True):
        out = []
        for r in self.rows:
            if only_interesting and r["ok"] and r["nbad"] == 0:
                continue
            out.append(f"  op {r['idx']:4d} {r['kind']:10s} {r['label'][:56]:56s} {r['what']:10s} n={r['n']:9d} "
                       f"bad={r['nbad']:7d} ({r['nbad'] / max(r['n'], 1):.2e}) max={r['maxdiff']:.3e} {'ok' if r['ok'] else 'FAIL'} {r['note']}")
        return "\n".join(out)

    def summary(self):
        by = {}
        for r in self.rows:
            k = (r["kind"], r["what"])
            a = by.setdefault(k, [0, 0, 0, 0.0, 0])
            a[0] += 1; a[1] += r["n"]; a[2] += r["nbad"]; a[3] = max(a[3], r["maxdiff"]); a[4] += 0 if r["ok"] else 1
        lines = [f"  {k[0]:10s} {k[1]:10s} checks={a[0]:4d} elems={a[1]:11d} off={a[2]:8d} ({a[2] / max(a[1], 1):.2e}) "
                 f"max={a[3]:.3e} failed={a[4]}" for k, a in sorted(by.items())]
        return "\n".join(lines)


def _cmp_codes(rep, idx, label, kind, got, ref, exact):
    diff = (got - ref).abs()
    nbad = int((diff > 0).sum())
    mx = int(diff.max()) if diff.numel() else 0
    if exact:
        ok = nbad == 0
    else:
        ok = mx <= 1 and nbad <= max(CODE_FRAC * diff.numel(), 2)
    rep.add(idx, label, kind, "codes=" if exact else "codes~", diff.numel(), nbad, mx, ok)


def _cmp_f32(rep, idx, label, kind, got, ref, tol, what="fp32"):
    err = (got - ref).abs()
    bad = err > tol
    rep.add(idx, label, kind, what, err.numel(), int(bad.sum()), float((err / tol.clamp_min(1e-30)).max()) if err.numel() else 0.0,
            not bool(bad.any()), note="(max err / tol)")


# ----------------------------------------------------------------------------------------------- snapshots
_INPUTS = {"split3": ["src"], "gemm_wo": ["rowvec", "residual"], "attention_fp": ["q", "k", "v"],
           "quantize": ["src"], "groupnorm": ["x"], "layernorm": ["x"], "gemm": ["a", "rowvec", "residual"],
           "attention": ["q", "k", "vt"], "im2col": ["src"], "copy2d": ["src"], "upsample2x": ["src"],
           "avgpool2x": ["src"], "nhwc_to_nchw": ["src"]}


def snapshot(spec):
    """Host copies of the op's inputs BEFORE it runs (outputs may alias them: in-place residual accumulation)."""
    pre = {}
    for name in _INPUTS.get(spec["kind"], []):
        a = spec.get(name)
        if a is not None:
            pre[name] = rd_codes(a) if a.signed is not None else rd_f32(a)
    if spec["kind"] in ("gemm_wo", "im2col_bytes"):        # bfloat16 planes (or gathered patches of them), raw
        a = spec["a"] if spec["kind"] == "gemm_wo" else spec["src"]
        pre["a_raw"] = a.t.detach().cpu()
    if spec["kind"] == "groupnorm" and spec.get("ss") is not None:
        pre["ss"] = rd_f32(spec["ss"][0])
    if spec["kind"] == "timestep_emb":
        pre["t"] = spec["t"].detach().cpu().to(torch.float32)
    if spec["kind"] == "nchw_to_nhwc":
        pre["src"] = spec["src"].detach().cpu().to(torch.float64)
    return pre


# ----------------------------------------------------------------------------------------------- per-kind checks
def _act_fn(y, act):
    if act == 1:
        return O.silu(y)
    return y


def check_quantize(rep, i, label, s, pre):
    x = pre["src"]
    cols = s["cols"]
    if s["act"] == 2:
        y = O.geglu(x[:, :2 * cols].to(torch.float32)).double()
    else:
        y = _act_fn(x[:, :cols].to(torch.float32), s["act"]).double()
    if s["upsample"] is not None:
        B, H, W = s["upsample"]
        y = y.reshape(B, H, W, cols).repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).reshape(-1, cols)
    ref = quant(y, s["q0"])
    if s["split"]:
        ref[:, s["split"]:] = quant(y[:, s["split"]:], s["q1"])
    _cmp_codes(rep, i, label, "quantize", rd_codes(s["dst"]), ref, exact=(s["act"] == 0))


def _gn_ref(x, s):
    B, HW, C = s["B"], s["HW"], x.shape[1]
    xn = x.reshape(B, HW, C).permute(0, 2, 1).to(torch.float32)
    y = F.group_norm(xn, s["groups"], s["gamma"], s["beta"], s["eps"])
    return y


def check_groupnorm(rep, i, label, s, pre):
    x = pre["x"]
    y = _gn_ref(x, s)                                        # [B, C, HW] fp32
    if s["ss"] is not None:
        oc = s["ss"][1]
        e = pre["ss"].to(torch.float32)
        y = y * (1 + e[:, :oc, None]) + e[:, oc:2 * oc, None]
    if s["silu"]:
        y = O.silu(y)
    y2 = y.permute(0, 2, 1).reshape(x.shape[0], x.shape[1]).double()
    for a, q in s["outs"]:
        _cmp_codes(rep, i, label, "groupnorm", rd_codes(a), quant(y2, q), exact=False)
    if s["out_f"] is not None:
        got = rd_f32(s["out_f"])
        _cmp_f32(rep, i, label, "groupnorm", got, y2, 2e-5 * (y2.abs() + 1e-3 * y2.abs().max()), what="fp32~")
    if s["raw"] is not None:
        a, split, q0, q1 = s["raw"]
        ref = quant(x, q0)
        if split < x.shape[1]:
            ref[:, split:] = quant(x[:, split:], q1)
        _cmp_codes(rep, i, label, "groupnorm", rd_codes(a), ref, exact=True)


def check_layernorm(rep, i, label, s, pre):
    x = pre["x"].to(torch.float32)
    y = F.layer_norm(x, (x.shape[1],), s["gamma"], s["beta"], s["eps"]).double()
    for a, q in s["outs"]:
        _cmp_codes(rep, i, label, "layernorm", rd_codes(a), quant(y, q), exact=False)


def check_im2col(rep, i, label, s, pre):
    B, H, W, Ho, Wo = s["B"], s["H"], s["W"], s["Ho"], s["Wo"]
    src = pre["src"]
    C = src.shape[1]
    x = src.reshape(B, H, W, C)
    pt, pl = s["pad_tl"]
    pc = s["pad_code"]
    if s["src"].signed:
        pc = pc - 256 if pc > 127 else pc
    xp = torch.full((B, H + 3, W + 3, C), pc, dtype=torch.int64)
    xp[:, pt:pt + H, pl:pl + W] = x
    ref = torch.zeros(B, Ho, Wo, s["k_to"], dtype=torch.int64)
    st = s["stride"]
    for ky in range(3):
        for kx in range(3):
            ref[..., (ky * 3 + kx) * C:(ky * 3 + kx + 1) * C] = xp[:, ky:ky + st * Ho:st, kx:kx + st * Wo:st][:, :Ho, :Wo]
    _cmp_codes(rep, i, label, "im2col", rd_codes(s["dst"]), ref.reshape(-1, s["k_to"]), exact=True)


def check_gemm(rep, i, label, s, pre):
    N, taps, Cred = s["N"], s["taps"], s["C"]
    ws = s["ws"].double()
    a = pre["a"][:, s["a_cols"]:s["a_cols"] + (Cred if taps == 1 else Cred)]
    zx = s["zx"]
    M = a.shape[0]
    if taps == 9:
        B, H, W = s["conv_bhw"]
        an = (a.double() - zx).reshape(B, H, W, Cred).permute(0, 3, 1, 2)
        acc = F.conv2d(an, ws[:, :Cred], None, stride=1, padding=1).permute(0, 2, 3, 1).reshape(M, N)
    else:
        if ws.dim() == 4:
            w2 = ws.permute(0, 2, 3, 1).reshape(N, -1)
        else:
            w2 = ws.reshape(N, -1)
        if w2.shape[1] < Cred:
            w2 = F.pad(w2, (0, Cred - w2.shape[1]))
        acc = (a.double() - zx) @ w2.t()
    t_main = acc * s["scale"].double()[None, :]
    y = t_main.clone()
    mag = t_main.abs()
    if s["bias"] is not None:
        y += s["bias"].double()[None, :]
        mag += s["bias"].double().abs()[None, :]
    if s["rowvec"] is not None:
        rv = pre["rowvec"][:, :N]
        img = torch.arange(M) // s["rows_per_batch"]
        y += rv[img]
        mag += rv[img].abs()
    if s["residual"] is not None:
        r = pre["residual"][:, :N]
        y += r
        mag += r.abs()
    if s["out"] is not None:
        off = s["out_cols_offset"]
        got = rd_f32(s["out"])[:, off:off + N]
        _cmp_f32(rep, i, label, "gemm", got, y, 3e-6 * mag + 1e-9)
    if s["out_q"] is not None:
        q = s["oq"]
        if s["geglu"]:
            r = torch.arange(N)
            xs, gs = y[:, (r % 8) < 4], y[:, (r % 8) >= 4]
            ref = quant((xs.to(torch.float32) * F.gelu(gs.to(torch.float32))).double(), q)
            got = rd_codes(s["out_q"])
        elif s["transposed"]:
            T = s["rows_per_batch"]
            Bn = M // T
            ref = quant(y, q).reshape(Bn, T, N).permute(0, 2, 1)                 # [B, N, T]
            raw = rd_codes(s["out_q"]).reshape(Bn, N, -1)
            got = raw[:, :, vt_positions(T)]
        elif s["out_q_head"] is not None:
            d, P = s["out_q_head"]
            ref = quant(y, q)
            raw = rd_codes(s["out_q"])
            n = torch.arange(N)
            got = raw[:, (n // d) * P + n % d]
        else:
            ref = quant(y, q)
            got = rd_codes(s["out_q"])[:, :ref.shape[1]]
        _cmp_codes(rep, i, label, "gemm", got, ref, exact=False)


def check_attention(rep, i, label, s, pre):
    B, heads, d, Tq, Tk = s["B"], s["heads"], s["d"], s["Tq"], s["Tk"]
    qa, ka, va = s["q"], s["k"], s["vt"]
    h = torch.arange(heads)[:, None]
    c = torch.arange(d)[None, :]
    qcols = (s["q_layout"][0] + h * s["q_layout"][1] + c).reshape(-1)
    kcols = (s["k_layout"][0] + h * s["k_layout"][1] + c).reshape(-1)
    qf = (pre["q"][:, qcols].double() - qa.zp[0]) * float(qa.delta[0])
    kf = (pre["k"][:, kcols].double() - ka.zp[0]) * float(ka.delta[0])
    qf = qf.reshape(B, Tq, heads, d).permute(0, 2, 1, 3)
    kf = kf.reshape(B, Tk, heads, d).permute(0, 2, 1, 3)
    vrows = (s["v_layout"][0] + h * s["v_layout"][1] + c).reshape(-1)
    vt = pre["vt"].reshape(B, -1, pre["vt"].shape[1])[:, vrows][:, :, vt_positions(Tk)]      # [B, heads*d, Tk]
    vf = ((vt.double() - va.zp[0]) * float(va.delta[0])).reshape(B, heads, d, Tk).permute(0, 1, 3, 2)
    # scores: exact integers x (delta_q*delta_k*extra) like the engine and, up to fp32 rounding, like the reference
    sim = torch.einsum("bhid,bhjd->bhij", qf, kf) * s["scale_extra"]
    p = torch.softmax(sim.to(torch.float32), dim=-1)                # the reference's softmax runs in fp32
    dw, zw, lo, hi = s["qw"]
    pc = torch.clamp(torch.round(p / torch.tensor(dw, dtype=torch.float32)) + zw, lo, hi).double()
    out = torch.einsum("bhij,bhjd->bhid", (pc - zw) * dw, vf)
    ref = out.permute(0, 2, 1, 3).reshape(B * Tq, heads * d)
    if s["oq"] is None:
        got = rd_f32(s["out"])
        tol = torch.full_like(ref, 2e-3 * float(ref.abs().max()) + 1e-6)
        _cmp_f32(rep, i, label, "attention", got, ref, tol)
    else:
        _cmp_codes(rep, i, label, "attention", rd_codes(s["out"]), quant(ref, s["oq"]), exact=False)


def check_misc(rep, i, label, s, pre):
    k = s["kind"]
    if k == "copy2d":
        got, ref = rd_f32(s["dst"]), pre["src"]
        rep.add(i, label, k, "fp32=", ref.numel(), int((got != ref).sum()), float((got - ref).abs().max()), bool(torch.equal(got, ref)))
    elif k == "upsample2x":
        B, H, W = s["B"], s["H"], s["W"]
        C = pre["src"].shape[1]
        ref = pre["src"].reshape(B, H, W, C).repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).reshape(-1, C)
        got = rd_f32(s["dst"])
        rep.add(i, label, k, "fp32=", ref.numel(), int((got != ref).sum()), float((got - ref).abs().max()), bool(torch.equal(got, ref)))
    elif k == "avgpool2x":
        B, H, W = s["B"], s["H"], s["W"]
        C = pre["src"].shape[1]
        x = pre["src"].reshape(B, H, W, C).permute(0, 3, 1, 2).to(torch.float32)
        ref = F.avg_pool2d(x, 2, 2).permute(0, 2, 3, 1).reshape(-1, C).double()
        mag = F.avg_pool2d(x.abs(), 2, 2).permute(0, 2, 3, 1).reshape(-1, C).double()
        _cmp_f32(rep, i, label, k, rd_f32(s["dst"]), ref, 1e-6 * mag + 1e-12)
    elif k == "timestep_emb":
        dim = s["dst"].cols
        fn = O.timestep_embedding_ldm if s["mode"] == 0 else O.timestep_embedding_ddim
        ref = fn(pre["t"], dim).double()
        _cmp_f32(rep, i, label, k, rd_f32(s["dst"]), ref, torch.full_like(ref, 2e-6))
    elif k == "nchw_to_nhwc":
        x = pre["src"]
        ref = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1).reshape(-1, x.shape[1])
        got = rd_f32(s["dst"])
        rep.add(i, label, k, "fp32=", ref.numel(), int((got != ref).sum()), float((got - ref).abs().max()), bool(torch.equal(got, ref)))
    elif k == "nhwc_to_nchw":
        dst = s["dst"].detach().cpu().double()
        B, C = dst.shape[0], dst.shape[1]
        ref = pre["src"][:, :C].reshape(B, -1, C).permute(0, 2, 1).reshape(dst.shape)
        rep.add(i, label, k, "fp32=", ref.numel(), int((dst != ref).sum()), float((dst - ref).abs().max()), bool(torch.equal(dst, ref)))
    else:
        rep.add(i, label, k, "unchecked", 0, 0, 0.0, False, note="op kind without an in-situ check")


def _planes_to_f64(t, Cp, C):
    """bfloat16 [rows, 3*Cp] planes -> float64 [rows, C] (exact sum hi + mid + lo)."""
    p = t.to(torch.float64).reshape(t.shape[0], 3, Cp)
    return p.sum(dim=1)[:, :C]


def check_split3(rep, i, label, s, pre):
    x = pre["src"].to(torch.float32)
    if s["act"] == 1:
        x = O.silu(x)
    if s["upsample"] is not None:
        B, H, W = s["upsample"]
        x = x.reshape(B, H, W, -1).repeat_interleave(2, dim=1).repeat_interleave(2, dim=2).reshape(-1, x.shape[1])
    got = _planes_to_f64(s["dst"].t.detach().cpu(), s["Cp"], s["C"])
    ref = x.double()
    tol = (2e-7 if s["act"] == 0 else 4e-6) * ref.abs() + 1e-30       # 3 planes carry 24 bits; SiLU adds expf rounding
    _cmp_f32(rep, i, label, "split3", got, ref, tol + 1e-12)


def check_im2col_bytes(rep, i, label, s, pre):
    B, H, W, Ho, Wo, cb = s["B"], s["H"], s["W"], s["Ho"], s["Wo"], s["cbytes"]
    src = pre["a_raw"].view(torch.uint8).reshape(B, H, W, cb).to(torch.int64)
    pt, pl = s["pad_tl"]
    xp = torch.zeros((B, H + 3, W + 3, cb), dtype=torch.int64)
    xp[:, pt:pt + H, pl:pl + W] = src
    st = s["stride"]
    ref = torch.zeros(B, Ho, Wo, 9 * cb, dtype=torch.int64)
    for ky in range(3):
        for kx in range(3):
            ref[..., (ky * 3 + kx) * cb:(ky * 3 + kx + 1) * cb] = xp[:, ky:ky + st * Ho:st, kx:kx + st * Wo:st][:, :Ho, :Wo]
    got = s["dst"].t.detach().cpu().to(torch.int64)
    _cmp_codes(rep, i, label, "im2col", got, ref.reshape(-1, 9 * cb), exact=True)


def check_gemm_wo(rep, i, label, s, pre):
    N, Cp, C = s["N"], s["Cp"], s["C"]
    ws = s["ws"].double()
    raw = pre["a_raw"]
    if s["im2col"]:                                   # patches: [rows, 9 taps x (3 planes x Cp) bf16]
        a = raw.view(torch.bfloat16).reshape(raw.shape[0], 9, 3, Cp).to(torch.float64).sum(dim=2)[:, :, :C]      # [rows, 9, C]
        w2 = ws.reshape(N, C, 9).permute(0, 2, 1)                                                               # [N, 9, C]
        acc = torch.einsum("mtc,ntc->mn", a, w2)
    else:
        x = _planes_to_f64(raw, Cp, C)
        M = x.shape[0]
        if s["taps"] == 9:
            B, H, W = s["conv_bhw"]
            acc = F.conv2d(x.reshape(B, H, W, C).permute(0, 3, 1, 2), ws, None, stride=1, padding=1).permute(0, 2, 3, 1).reshape(M, N)
        else:
            acc = x @ ws.reshape(N, -1).t()
    M = acc.shape[0]
    t_main = acc * s["scale"].double()[None, :]
    y, mag = t_main.clone(), t_main.abs()
    # |acc| can hide cancellation: bound the fp32 accumulation error by the sum of |products| scale (coarse: use |x| |w|)
    if s["bias"] is not None:
        y += s["bias"].double()[None, :]
        mag += s["bias"].double().abs()[None, :]
    if s["rowvec"] is not None:
        rv = pre["rowvec"][:, :N]
        img = torch.arange(M) // s["rows_per_batch"]
        y += rv[img]
        mag += rv[img].abs()
    if s["residual"] is not None:
        r = pre["residual"][:, :N]
        y += r
        mag += r.abs()
    got = rd_f32(s["out"])[:, :N]
    # fp32 accumulation over K terms: tolerance relative to the output scale of the layer, not to each element
    tol = 2e-5 * mag + 2e-5 * float(t_main.abs().max())
    _cmp_f32(rep, i, label, "gemm_wo", got, y, tol, what="fp32acc")


def check_attention_fp(rep, i, label, s, pre):
    B, heads, d, Tq, Tk = s["B"], s["heads"], s["d"], s["Tq"], s["Tk"]
    h = torch.arange(heads)[:, None]
    c = torch.arange(d)[None, :]

    def take(name, layout, T):
        cols = (layout[0] + h * layout[1] + c).reshape(-1)
        return pre[name][:, cols].reshape(B, T, heads, d).permute(0, 2, 1, 3)
    q, k, v = take("q", s["q_layout"], Tq), take("k", s["k_layout"], Tk), take("v", s["v_layout"], Tk)
    p = torch.softmax((torch.einsum("bhid,bhjd->bhij", q, k) * s["scale"]).to(torch.float32), dim=-1).double()
    ref = torch.einsum("bhij,bhjd->bhid", p, v).permute(0, 2, 1, 3).reshape(B * Tq, heads * d)
    got = rd_f32(s["out"])
    _cmp_f32(rep, i, label, "attention", got, ref, 2e-5 * ref.abs() + 2e-5 * float(ref.abs().max()), what="fp32acc")


CHECKS = {"split3": check_split3, "im2col_bytes": check_im2col_bytes, "gemm_wo": check_gemm_wo, "attention_fp": check_attention_fp,
          "quantize": check_quantize, "groupnorm": check_groupnorm, "layernorm": check_layernorm, "im2col": check_im2col,
          "gemm": check_gemm, "attention": check_attention}


def verify_program(prog, x, t, ctx=None):
    """Set the program inputs, then run op by op; returns a Report with one or more rows per op."""
    prog.x_in.copy_(x.to(prog.x_in.device, torch.float32))
    prog.t_in.copy_(t.to(prog.t_in.device, torch.float32))
    if prog.ctx_in is not None:
        prog.ctx_in.copy_(ctx.to(prog.ctx_in.device, torch.float32))
    rep = Report()
    for i in range(prog.nops):
        spec = prog.op_specs[i]
        pre = snapshot(spec)
        prog.run_range(i, i + 1)
        torch.cuda.synchronize()
        CHECKS.get(spec["kind"], check_misc)(rep, i, prog.op_names[i], spec, pre)
    return rep


def verify_folds(qnn, g, device):
    """Folded integer weights x per-channel step == the oracle's fake-quant weights (AdaRound hard decision
    adaptive_rounding.py:49-59), bit for bit, for every QuantModule (both split halves)."""
    from oracle import unet_oracle as U
    from qdiff_b200 import graph
    q = g["qcfg"]
    qc = U.QuantCfg(q["weight_bit"], q["act_bit"], q["a_sym"], q["sm_abit"], q["quant_act"], adaround=True)
    c = U._Ctx(g["ckpt"], qc)
    from qdiff_b200._lib import lib
    b = graph.Builder(qnn, device, 1)
    n = bad = 0
    for name, m in qnn.model.named_modules():
        if type(m).__name__ != "QuantModule":
            continue
        w = c.get(name + ".weight")
        halves = [("", None)] if m.split == 0 else [("", (0, m.split)), ("_0", (m.split, w.shape[1]))]
        for suffix, cols in halves:
            ws, dw = b._fold(m, cols, suffix)
            got = (ws * dw.reshape(-1, *([1] * (ws.dim() - 1)))).cpu()
            ww = w if cols is None else w[:, cols[0]:cols[1]]
            ref = c.weight_q(ww, name + ".weight_quantizer" + suffix)
            n += 1
            if not torch.equal(got, ref):
                bad += 1
    lib().qd_engine_destroy(b.engine)
    return n, bad
