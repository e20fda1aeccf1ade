"""Op-by-op parity report of the CUDA kernels against PyTorch fp32 references (run on a B200 via gpurun).

    python tests/gpu_opcheck.py [filter-substring] > gpurun_out/opcheck.log

Unlike the pytest suite it keeps going after a failure and prints error statistics for every op, which is what is
needed when bringing kernels up without an interactive GPU."""
import os
import sys
import traceback

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-feature-vit_b200"))
sys.path.insert(0, ROOT)

from mfvit import ops  # noqa: E402
from mfvit._lib import EPI_BF16, EPI_DGELU, EPI_F32, EPI_GELU, EPI_RESID_F32, FusionGrads  # noqa: E402

dev = torch.device("cuda:0")
RESULTS = []


def report(name, got, ref, tol, rel=False):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    maxerr = err.max().item()
    val = maxerr / denom if rel else maxerr
    cos = F.cosine_similarity(got.flatten().double(), ref.flatten().double(), dim=0).item()
    ok = bool(val <= tol) and not bool(torch.isnan(got).any())
    RESULTS.append((name, ok))
    print("%-52s %s maxerr=%.3e (ref max %.3e) %s=%.3e tol=%.1e cos=%.6f" % (
        name, "OK  " if ok else "FAIL", maxerr, denom, "rel" if rel else "abs", val, tol, cos), flush=True)
    return ok


def run(name, fn, flt):
    if flt and flt not in name:
        return
    try:
        fn()
        torch.cuda.synchronize()
    except Exception:  # noqa: BLE001
        RESULTS.append((name, False))
        print("%-52s EXCEPTION\n%s" % (name, traceback.format_exc()), flush=True)


def bf(x):
    return x.to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------------------ GEMM
def t_gemm_fwd(G, M, N, K, block_n=0):
    def f():
        torch.manual_seed(0)
        x = bf(torch.randn(G, M, K, device=dev))
        w = bf(torch.randn(G, N, K, device=dev) * 0.05)
        b = torch.randn(G, N, device=dev)
        ref = torch.einsum("gmk,gnk->gmn", x.float(), w.float()) + b[:, None, :]
        out = ops.linear_fwd(x, w, b, EPI_BF16, block_n=block_n)
        report("gemm_fwd bf16 G%d M%d N%d K%d bn%d" % (G, M, N, K, block_n), out, ref, 2e-2, rel=True)
        out32 = ops.linear_fwd(x, w, b, EPI_F32, block_n=block_n)
        report("gemm_fwd f32  G%d M%d N%d K%d bn%d" % (G, M, N, K, block_n), out32, ref, 1e-4, rel=True)
    return f


def t_gemm_rows96():
    """opt-in 192-row pair tiles (rows_per_cta=96) of the 384-wide tile: residual forward and a bf16 dgrad, ragged M"""
    for M in (6304, 197 * 3, 100):
        torch.manual_seed(3)
        x = bf(torch.randn(2, M, 256, device=dev)); w = bf(torch.randn(2, 384, 256, device=dev) * 0.05)
        b = torch.randn(2, 384, device=dev); res = torch.randn(2, M, 384, device=dev)
        ref = torch.einsum("gmk,gnk->gmn", x.float(), w.float()) + b[:, None, :] + res
        out = ops.linear_fwd(x, w, b, EPI_RESID_F32, aux=res, block_n=384, cta_group=2, rows_per_cta=96)
        report("gemm rows96 resid M%d" % M, out, ref, 1e-4, rel=True)
        dy = bf(torch.randn(2, M, 512, device=dev)); w2 = bf(torch.randn(2, 512, 384, device=dev) * 0.05)
        out = ops.linear_dgrad(dy, w2, EPI_BF16, block_n=384, cta_group=2, rows_per_cta=96)
        report("gemm rows96 dgrad M%d" % M, out, torch.einsum("gmn,gnk->gmk", dy.float(), w2.float()), 2e-2, rel=True)


def t_gemm_streamk():
    """opt-in stream-K (mfv_set_option("streamk", mask)): the (tile, k-block) space cut into equal ranges per CTA pair,
    partial accumulators of a shared tile exchanged through the workspace.  Every family, against the whole-tile launch."""
    from mfvit import _lib
    lib = _lib.load()
    M = 6304
    default_mask = int(os.environ.get("MFVIT_STREAMK", "0"))

    def both(fn):
        assert lib.mfv_set_option(b"streamk", 0) == 0
        a = fn()
        assert lib.mfv_set_option(b"streamk", 15) == 0 and lib.mfv_set_option(b"streamk_min_kb", 2) == 0
        try:
            b = fn()
        finally:
            assert lib.mfv_set_option(b"streamk", default_mask) == 0 and lib.mfv_set_option(b"streamk_min_kb", 12) == 0
        return a, b

    torch.manual_seed(11)
    x = bf(torch.randn(2, M, 1536, device=dev)); w = bf(torch.randn(2, 384, 1536, device=dev) * 0.05)
    b = torch.randn(2, 384, device=dev); res = torch.randn(2, M, 384, device=dev)
    o0, o1 = both(lambda: ops.linear_fwd(x, w, b, EPI_RESID_F32, aux=res, block_n=384))
    report("stream-K resid fc2 (50 tiles on 74 pairs) vs whole tiles", o1, o0, 2e-5, rel=True)
    assert float((o1 - o0).abs().max()) > 0.0, "stream-K did not engage (bit-identical to the whole-tile launch)"
    x2 = bf(torch.randn(2, 2 * M, 1536, device=dev)); res2 = torch.randn(2, 2 * M, 384, device=dev)
    p0, p1 = both(lambda: ops.linear_fwd(x2, w, b, EPI_RESID_F32, aux=res2, block_n=384))
    report("stream-K resid fc2 (100 tiles) vs whole tiles", p1, p0, 2e-5, rel=True)
    x3 = bf(torch.randn(2, 197 * 7, 1536, device=dev)); res3 = torch.randn(2, 197 * 7, 384, device=dev)
    s0, s1 = both(lambda: ops.linear_fwd(x3, w, b, EPI_RESID_F32, aux=res3, block_n=384))
    report("stream-K resid fc2 (12 tiles, ragged rows) vs whole tiles", s1, s0, 2e-5, rel=True)
    report("stream-K resid fc2 vs fp32", o1, torch.einsum("gmk,gnk->gmn", x.float(), w.float()) + b[:, None, :] + res, 1e-4, rel=True)
    gam = 1 + 0.1 * torch.randn(2, 384, device=dev); bet = 0.1 * torch.randn(2, 384, device=dev)
    r0, r1 = both(lambda: ops.linear_fwd_ln(x, w, b, res, gam, bet))
    for nm, t0, t1 in zip(("x", "ln", "copy", "mean", "rstd"), r0, r1):
        if t0 is not None and torch.is_tensor(t0):
            report("stream-K resid+LN %s vs whole tiles" % nm, t1.float(), t0.float(), 4e-3 if t0.dtype != torch.float32 else 2e-5, rel=True)
    dy = bf(torch.randn(2, M, 1152, device=dev)); w2 = bf(torch.randn(2, 1152, 384, device=dev) * 0.05)
    d0, d1 = both(lambda: ops.linear_dgrad(dy, w2, EPI_BF16, block_n=384))
    report("stream-K bf16 dgrad qkv vs whole tiles", d1.float(), d0.float(), 8e-3, rel=True)
    xs = bf(torch.randn(2, M, 384, device=dev)); w3 = bf(torch.randn(2, 1152, 384, device=dev) * 0.05); b3 = torch.randn(2, 1152, device=dev)
    q0, q1 = both(lambda: ops.linear_fwd(xs, w3, b3, EPI_BF16))
    report("stream-K bf16 qkv (250 tiles) vs whole tiles", q1.float(), q0.float(), 8e-3, rel=True)
    w4 = bf(torch.randn(2, 1536, 384, device=dev) * 0.05); b4 = torch.randn(2, 1536, device=dev)
    def fc1():
        u = torch.empty(2, M, 1536, device=dev, dtype=torch.bfloat16); g = torch.empty_like(u)
        ops.linear_fwd(xs, w4, b4, EPI_GELU, out=u, out2=g)
        return u, g
    g0, g1 = both(fc1)
    for nm, t0, t1 in zip(("u", "gelu"), g0, g1):
        report("stream-K fc1 GELU %s vs whole tiles" % nm, t1.float(), t0.float(), 8e-3, rel=True)


def t_gemm_multicast():
    """opt-in clusters of four CTAs (mfv_set_option("gemm_mc", 1)): two pairs on adjacent row tiles share the B operand by
    TMA multicast.  Same arithmetic in the same order: bit-identical to the plain pair kernel, odd row-tile counts included."""
    from mfvit import _lib
    lib = _lib.load()

    def both(fn):
        assert lib.mfv_set_option(b"gemm_mc", 0) == 0
        a = fn()
        assert lib.mfv_set_option(b"gemm_mc", 1) == 0
        try:
            b = fn()
        finally:
            assert lib.mfv_set_option(b"gemm_mc", 0) == 0
        return a, b

    for M in (6304, 197 * 3, 2 * 6304 + 77):
        torch.manual_seed(13)
        x = bf(torch.randn(2, M, 1536, device=dev)); w = bf(torch.randn(2, 384, 1536, device=dev) * 0.05)
        b = torch.randn(2, 384, device=dev); res = torch.randn(2, M, 384, device=dev)
        o0, o1 = both(lambda: ops.linear_fwd(x, w, b, EPI_RESID_F32, aux=res, block_n=384))
        report("multicast resid M%d vs pair kernel" % M, o1, o0, 0.0)
        report("multicast resid M%d vs fp32" % M, o1, torch.einsum("gmk,gnk->gmn", x.float(), w.float()) + b[:, None, :] + res, 1e-4, rel=True)
        gam = 1 + 0.1 * torch.randn(2, 384, device=dev); bet = 0.1 * torch.randn(2, 384, device=dev)
        r0, r1 = both(lambda: ops.linear_fwd_ln(x, w, b, res, gam, bet))
        for nm, t0, t1 in zip(("x", "ln", "copy", "mean", "rstd"), r0, r1):
            if torch.is_tensor(t0):
                report("multicast resid+LN %s M%d vs pair kernel" % (nm, M), t1.float(), t0.float(), 0.0)
        dy = bf(torch.randn(2, M, 1152, device=dev)); w2 = bf(torch.randn(2, 1152, 384, device=dev) * 0.05)
        d0, d1 = both(lambda: ops.linear_dgrad(dy, w2, EPI_BF16, block_n=384))
        report("multicast bf16 dgrad (B MN-major) M%d vs pair kernel" % M, d1.float(), d0.float(), 0.0)


def t_gemm_resid_ln():
    """MFV_EPI_RESID_LN: residual add + LayerNorm of the finished row inside the proj / fc2 epilogue."""
    for M, K, f16, rows in ((6304, 384, True, 0), (197 * 3, 1536, False, 0), (1379, 384, True, 96), (200, 256, False, 0)):
        torch.manual_seed(5)
        dt = torch.float16 if f16 else torch.bfloat16
        x = torch.randn(2, M, K, device=dev).to(dt)
        w = (torch.randn(2, 384, K, device=dev) * 0.05).to(dt)
        b = torch.randn(2, 384, device=dev) * 0.1
        res = torch.randn(2, M, 384, device=dev) * 2 + 0.7  # non-zero row means: the variance must not cancel
        gam = 1 + 0.1 * torch.randn(2, 384, device=dev)
        bet = 0.1 * torch.randn(2, 384, device=dev)
        xr = torch.einsum("gmk,gnk->gmn", x.float(), w.float()) + b[:, None, :] + res
        mu, var = xr.mean(-1), xr.var(-1, unbiased=False)
        yr = (xr - mu[..., None]) * torch.rsqrt(var[..., None] + 1e-6) * gam[:, None, :] + bet[:, None, :]
        tag = "M%d K%d %s rows%d" % (M, K, "f16" if f16 else "bf16", rows)
        x_new, y, ycopy, mean, rstd = ops.linear_fwd_ln(x, w, b, res, gam, bet, 1e-6, f16=f16, bf16_copy=f16, rows_per_cta=rows)
        report("gemm resid+LN x " + tag, x_new, xr, 1e-4, rel=True)
        report("gemm resid+LN mean " + tag, mean, mu, 1e-5)
        report("gemm resid+LN rstd " + tag, rstd, torch.rsqrt(var + 1e-6), 1e-4, rel=True)
        report("gemm resid+LN y " + tag, y, yr, 2e-3 if f16 else 1.6e-2, rel=False)
        if ycopy is not None:
            report("gemm resid+LN y bf16 copy " + tag, ycopy, yr, 1.6e-2)
        _, y32, _, _, _ = ops.linear_fwd_ln(x, w, b, res, gam, bet, 1e-6, out_f32=True)
        report("gemm resid+LN y f32 " + tag, y32, yr, 2e-5, rel=True)


def t_wgrad_pair():
    """mfv_gemm_wgrad_pair: two split-K weight gradients (+ folded bias gradients) as one grid, vs fp32 einsum."""
    for M, splits in ((6304, 5), (197 * 3, 2), (1000, 1)):
        torch.manual_seed(9)
        G = 2
        dy0 = bf(torch.randn(G, M, 384, device=dev)); x0 = bf(torch.randn(G, M, 1536, device=dev))      # fc2-like: dW [384][1536]
        dy1 = bf(torch.randn(G, M, 1536, device=dev)); x1 = bf(torch.randn(G, M, 384, device=dev))      # fc1-like: dW [1536][384]
        dw0 = torch.zeros(G, 384, 1536, device=dev); dw1 = torch.zeros(G, 1536, 384, device=dev)
        # both bias gradients live at one group stride inside a common buffer, as in the engine's flat gradient buffer
        dbuf = torch.zeros(G, 4096, device=dev)
        db1 = dbuf[:, :1536]
        ops.linear_wgrad_pair(dy0, x0, dw0, None, dy1, x1, dw1, None, splits=splits)
        report("wgrad pair dW0 M%d s%d" % (M, splits), dw0, torch.einsum("gmn,gmk->gnk", dy0.float(), x0.float()), 2e-3, rel=True)
        report("wgrad pair dW1 M%d s%d" % (M, splits), dw1, torch.einsum("gmn,gmk->gnk", dy1.float(), x1.float()), 2e-3, rel=True)
    # with the bias gradient of problem 1 folded in (row sums of dY through the ones-UMMA): strided view into a flat buffer
    from mfvit._lib import GemmArgs  # noqa: F401
    import ctypes as C
    M, G = 1379, 2
    dy0 = bf(torch.randn(G, M, 384, device=dev)); x0 = bf(torch.randn(G, M, 384, device=dev))          # proj-like
    dy1 = bf(torch.randn(G, M, 1152, device=dev)); x1 = bf(torch.randn(G, M, 384, device=dev))         # qkv-like
    flat = torch.zeros(G, 1152 * 384 + 384 * 384 + 1152 + 8, device=dev)
    P = flat.shape[1]
    o_w0, o_w1, o_b1 = 0, 384 * 384, 384 * 384 + 1152 * 384

    def args(dy, x, w_off, b_off):
        Gg, Mm, N = dy.shape
        K = x.shape[2]
        a = GemmArgs()
        a.A, a.B, a.C = dy.data_ptr(), x.data_ptr(), flat.data_ptr() + 4 * w_off
        a.M, a.N, a.K, a.G = N, K, Mm, Gg
        a.lda, a.ldb, a.ldc = N, K, K
        a.a_gstride, a.b_gstride, a.c_gstride = Mm * N, Mm * K, P
        a.a_mn_major, a.b_mn_major, a.epilogue, a.splits = 1, 1, 5, 4
        a.bias_gstride = P
        a.row_sum = flat.data_ptr() + 4 * b_off if b_off is not None else None
        return a
    from mfvit import _lib
    a0, a1 = args(dy0, x0, o_w0, None), args(dy1, x1, o_w1, o_b1)
    _lib.check(_lib.init(0).mfv_gemm_wgrad_pair(C.byref(a0), C.byref(a1), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "mfv_gemm_wgrad_pair")
    report("wgrad pair (flat buffer) dW proj", flat[:, o_w0:o_w0 + 384 * 384].reshape(G, 384, 384),
           torch.einsum("gmn,gmk->gnk", dy0.float(), x0.float()), 2e-3, rel=True)
    report("wgrad pair (flat buffer) dW qkv", flat[:, o_w1:o_w1 + 1152 * 384].reshape(G, 1152, 384),
           torch.einsum("gmn,gmk->gnk", dy1.float(), x1.float()), 2e-3, rel=True)
    report("wgrad pair (flat buffer) db qkv (folded)", flat[:, o_b1:o_b1 + 1152], dy1.float().sum(1), 2e-3, rel=True)


def t_patch_embed_tma():
    """im2col-free patch embedding (5-D TMA tiles from NCHW fp32, 16-bit conversion in shared memory, tcgen05) vs Conv2d +
    flatten + cls + pos on the same 16-bit-rounded operands; and the raw TMA box layout the converter warps rely on."""
    import ctypes as C
    from mfvit import _lib
    lib = _lib.init(0)
    HW, Bp = 224, 2
    img = torch.arange(Bp * 3 * HW * HW, device=dev, dtype=torch.float32).view(Bp, 3, HW, HW)
    raw = torch.zeros(8192, device=dev)
    kb, ph0, cb = 6, 9, 4  # channel 1 of image 1, pixel rows 8..11 of the patches, second band (5 real patch rows)
    rc = lib.mfv_debug_patch_tma_probe(C.c_void_p(img.data_ptr()), C.c_void_p(raw.data_ptr()), Bp, HW, kb, ph0, cb,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc
    gw, PH = 14, 9
    exp = torch.zeros(128, 4, 16, device=dev)
    exp[gw * PH:] = -777.0
    for ph in range(min(PH, gw - ph0)):
        for pw in range(gw):
            exp[ph * gw + pw] = img[1, 1, (ph0 + ph) * 16 + 8:(ph0 + ph) * 16 + 12, pw * 16:pw * 16 + 16]
    report("patch embed TMA box layout [patch][i][j] (+ zero fill past the image)", raw.view(128, 4, 16), exp, 0.0)
    for B, HW, G, f16 in ((3, 224, 2, True), (2, 384, 1, True), (5, 224, 1, False)):
        torch.manual_seed(8)
        gw = HW // 16
        S = 1 + gw * gw
        dt = torch.float16 if f16 else torch.bfloat16
        imgs = [torch.randn(B, 3, HW, HW, device=dev) * 1.5 + 0.3 for _ in range(G)]
        w = torch.randn(G, 384, 768, device=dev) * 0.05
        bias = torch.randn(G, 384, device=dev) * 0.1
        cls = torch.randn(G, 384, device=dev)
        pos = torch.randn(G, S, 384, device=dev)
        x = ops.patch_embed_tma(imgs, w, bias, cls, pos, f16=f16)
        ref = torch.empty_like(x)
        for g in range(G):
            t = F.conv2d(imgs[g].to(dt).float(), w[g].to(dt).float().view(384, 3, 16, 16), bias[g], stride=16)
            ref[g, :, 1:] = t.flatten(2).transpose(1, 2) + pos[g, 1:]
            ref[g, :, 0] = cls[g] + pos[g, 0]
        tag = "B%d %dpx G%d %s" % (B, HW, G, "f16" if f16 else "bf16")
        report("patch embed TMA cls row " + tag, x[:, :, 0], ref[:, :, 0], 1e-6)
        report("patch embed TMA tokens " + tag, x[:, :, 1:], ref[:, :, 1:], 2e-5, rel=True)


def t_gemm_epilogues():
    torch.manual_seed(1)
    G, M, N, K = 2, 1000, 384, 256
    x = bf(torch.randn(G, M, K, device=dev))
    w = bf(torch.randn(G, N, K, device=dev) * 0.05)
    b = torch.randn(G, N, device=dev)
    res = torch.randn(G, M, N, device=dev)
    acc = torch.einsum("gmk,gnk->gmn", x.float(), w.float()) + b[:, None, :]
    out = ops.linear_fwd(x, w, b, EPI_RESID_F32, aux=res)
    report("gemm epilogue resid_f32", out, acc + res, 1e-4, rel=True)
    u = torch.empty(G, M, N, device=dev, dtype=torch.bfloat16)
    g = torch.empty_like(u)
    ops.linear_fwd(x, w, b, EPI_GELU, out=u, out2=g)
    report("gemm epilogue gelu: u", u, acc, 2e-2, rel=True)
    report("gemm epilogue gelu: g", g, F.gelu(acc), 2e-2, rel=True)
    # dgelu: C = (dy @ W) * gelu'(u)
    dy = bf(torch.randn(G, M, N, device=dev))
    w2 = bf(torch.randn(G, N, K, device=dev) * 0.05)  # [N,K] -> dx [M,K]
    upre = bf(torch.randn(G, M, K, device=dev))
    dxr = torch.einsum("gmn,gnk->gmk", dy.float(), w2.float())
    uu = upre.float().requires_grad_(True)
    F.gelu(uu).backward(torch.ones_like(uu))
    out = ops.linear_dgrad(dy, w2, EPI_DGELU, aux=upre)
    report("gemm dgrad + dgelu", out, dxr * uu.grad, 2e-2, rel=True)
    g2 = torch.empty_like(upre)
    for bn in (128, 256):
        out = ops.linear_dgrad(dy, w2, EPI_DGELU, aux=upre, out2=g2, block_n=bn)
        report("gemm dgrad + dgelu (+gelu out) bn%d" % bn, out, dxr * uu.grad, 2e-2, rel=True)
        report("gemm dgrad recomputed gelu bn%d" % bn, g2, F.gelu(upre.float()), 1e-2, rel=True)
    out = ops.linear_dgrad(dy, w2, EPI_BF16)
    report("gemm dgrad (B mn-major)", out, dxr, 2e-2, rel=True)


def t_gemm_wgrad(G, M, N, K, splits):
    def f():
        torch.manual_seed(2)
        dy = bf(torch.randn(G, M, N, device=dev))
        x = bf(torch.randn(G, M, K, device=dev))
        ref = torch.einsum("gmn,gmk->gnk", dy.float(), x.float())
        dw = torch.zeros(G, N, K, device=dev)
        fold = K == 384 and N > 128  # 384-wide pair tile: the bias gradient rides along as an extra UMMA
        db = torch.zeros(G, N, device=dev) if fold else None
        ops.linear_wgrad(dy, x, dw, splits=splits, db=db)
        report("gemm wgrad G%d M%d N%d K%d s%d" % (G, M, N, K, splits), dw, ref, 1e-4, rel=True)
        if fold:
            report("gemm wgrad folded bias grad G%d M%d N%d" % (G, M, N), db, dy.float().sum(1), 1e-4, rel=True)
    return f


# ------------------------------------------------------------------------------------------------------------ LN
def t_ln():
    torch.manual_seed(3)
    G, rows, C = 2, 1234, 384
    x = torch.randn(G, rows, C, device=dev) * 2 + 0.5
    gam = torch.randn(G, C, device=dev)
    bet = torch.randn(G, C, device=dev)
    y16, y32, mean, rstd = ops.layernorm_fwd(x, gam, bet, 1e-6, want_bf16=True, want_f32=True)
    xr = x.clone().requires_grad_(True)
    gr = gam.clone().requires_grad_(True)
    br = bet.clone().requires_grad_(True)
    ref = torch.stack([F.layer_norm(xr[g], (C,), gr[g], br[g], 1e-6) for g in range(G)])
    report("layernorm fwd f32", y32, ref, 1e-5)
    report("layernorm fwd bf16", y16, ref, 2e-2, rel=True)
    dy = torch.randn_like(x)
    dres = torch.randn_like(x)
    ref.backward(dy)
    dgam = torch.zeros_like(gam)
    dbet = torch.zeros_like(bet)
    dxs = torch.zeros_like(gam)
    dx, dx16 = ops.layernorm_bwd(dy, x, mean, rstd, gam, dres=dres, dgamma=dgam, dbeta=dbet, dx_colsum=dxs)
    report("layernorm bwd colsum(dx) (folded bias grad)", dxs, (xr.grad + dres).sum(1), 1e-3, rel=True)
    report("layernorm bwd dx (f32 dy, +dres)", dx, xr.grad + dres, 1e-4)
    report("layernorm bwd dx bf16 copy", dx16, xr.grad + dres, 2e-2, rel=True)
    report("layernorm bwd dgamma", dgam, gr.grad, 1e-3, rel=True)
    report("layernorm bwd dbeta", dbet, br.grad, 1e-3, rel=True)
    dx2, _ = ops.layernorm_bwd(bf(dy), x, mean, rstd, gam)
    report("layernorm bwd dx (bf16 dy)", dx2, xr.grad, 3e-2, rel=True)


# ------------------------------------------------------------------------------------------------------------ attention
def t_attn(NB, S, H, D, f16=False, amp=1.0, check_bwd=True):
    """amp > 1 widens the score range (q and k scaled): block maxima then differ by more than the forward's lazy-rescaling
    threshold (2^8), so the path that waits for the previous P.V and rescales O is exercised too.  (The fp16-mode backward
    recomputes the scores from bf16-rounded q / k: with scores this wide that alone moves the gradients by 4-5 %, so the
    wide fp16 case checks the forward only.)"""
    def f():
        torch.manual_seed(4)
        qkv = torch.randn(NB, S, 3, H, D, device=dev)
        qkv[:, :, :2] *= amp
        qkv = qkv.half() if f16 else bf(qkv)
        if f16:
            o16, lse, o = ops.attn_fwd(qkv, H, f16=True, bf16_copy=True)
            report("attn fwd o fp16 vs bf16 copy", o16, o, 1e-2, rel=True)
        else:
            o, lse = ops.attn_fwd(qkv, H)
        q, k, v = [qkv[:, :, i].float().permute(0, 2, 1, 3).detach().requires_grad_(True) for i in range(3)]
        s = (q @ k.transpose(-1, -2)) * D ** -0.5
        ref = (s.softmax(-1) @ v)
        tag = "NB%d S%d H%d D%d%s%s" % (NB, S, H, D, " f16" if f16 else "", " amp%g" % amp if amp != 1.0 else "")
        report("attn fwd o " + tag, o.permute(0, 2, 1, 3), ref, 2e-2, rel=True)
        report("attn fwd lse " + tag, lse, torch.logsumexp(s, -1), 1e-3)
        if not check_bwd:
            return
        do = bf(torch.randn(NB, S, H, D, device=dev))
        ref.backward(do.float().permute(0, 2, 1, 3))
        dqkv = ops.attn_bwd(qkv, o, do, lse)
        for i, (nm, t) in enumerate((("dq", q), ("dk", k), ("dv", v))):
            report("attn bwd %s %s" % (nm, tag), dqkv[:, :, i].permute(0, 2, 1, 3), t.grad, 3e-2, rel=True)
    return f


# ------------------------------------------------------------------------------------------------------------ fusion
def t_fusion():
    from oracle.fusion_ref import Fus_CrossViT
    torch.manual_seed(5)
    B, S, C, heads, NC = 6, 197, 384, 3, 3

    class Stub:
        def features3D(self, x):
            return x
    fus = Fus_CrossViT(Stub(), Stub()).to(dev)
    with torch.no_grad():
        for p in fus.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    vh = [torch.nn.Linear(C, NC).to(dev) for _ in range(2)]
    tok = (torch.randn(2, B, S, C, device=dev) * 1.5).requires_grad_(True)
    fused_ref = fus.fuse(tok[0], tok[1])
    cf = fus.closed_form(tok[0], tok[1])
    report("fusion oracle closed-form == as-written", cf, fused_ref, 1e-5)
    x_ref = torch.stack([vh[0](tok[0][:, 0]), vh[1](tok[1][:, 0])])
    L = fus.multi_scale_transformers[0].cross_attn_layers[0]
    # direction 0: CXR cls queries ENH patches: PreNorm L[0], post-LN L[3], mlp_head_cxr, vit_cxr.head
    # direction 1: ENH cls queries CXR patches: PreNorm L[2], post-LN L[1], mlp_head_enh, vit_enh.head
    pre = (L[0], L[2]); post = (L[3], L[1]); heads_m = (fus.mlp_head_cxr[0], fus.mlp_head_enh[0])
    names = {
        "ln1_w": [m.norm.weight for m in pre], "ln1_b": [m.norm.bias for m in pre],
        "wq": [m.fn.wq.weight for m in pre], "wk": [m.fn.wk.weight for m in pre], "wv": [m.fn.wv.weight for m in pre],
        "proj_w": [m.fn.proj.weight for m in pre], "proj_b": [m.fn.proj.bias for m in pre],
        "ln2_w": [m.weight for m in post], "ln2_b": [m.bias for m in post],
        "head_w": [m.weight for m in heads_m], "head_b": [m.bias for m in heads_m],
        "vhead_w": [m.weight for m in vh], "vhead_b": [m.bias for m in vh],
    }
    prm = ops.fusion_param_struct({k: tuple(t.detach() for t in v) for k, v in names.items()})
    fused, x = ops.fusion_fwd(tok.detach(), prm, B, S, C, heads, NC)
    report("fusion fwd fused", fused, fused_ref, 1e-4)
    report("fusion fwd x (backbone heads)", x, x_ref, 1e-4)
    d_fused = torch.randn(B, NC, device=dev)
    d_x = torch.randn(2, B, NC, device=dev)
    (fused_ref * d_fused).sum().add((x_ref * d_x).sum()).backward()
    gt = {k: tuple(torch.zeros_like(t) for t in v) for k, v in names.items()}
    grads = ops.fusion_param_struct(gt, cls=FusionGrads)
    dtok = ops.fusion_bwd(tok.detach(), prm, grads, d_fused, d_x, B, S, C, heads, NC)
    report("fusion bwd dtok", dtok, tok.grad, 1e-3, rel=True)
    for k, v in names.items():
        for d in range(2):
            report("fusion bwd d%s[%d]" % (k, d), gt[k][d], v[d].grad, 2e-3, rel=True)
    # batched stage kernels: forward with a saved buffer, backward reusing its state (what the trainer / autograd do)
    for Bb in (B, 33):
        tk = (torch.randn(2, Bb, S, C, device=dev) * 1.5).requires_grad_(True)
        f_ref = fus.fuse(tk[0], tk[1])
        x_r = torch.stack([vh[0](tk[0][:, 0]), vh[1](tk[1][:, 0])])
        for v in names.values():
            for t in v:
                t.grad = None
        dfu, dxx = torch.randn(Bb, NC, device=dev), torch.randn(2, Bb, NC, device=dev)
        (f_ref * dfu).sum().add((x_r * dxx).sum()).backward()
        scratch = ops.fusion_scratch(tk, Bb, S, C, heads)
        fused_b, x_b = ops.fusion_fwd(tk.detach(), prm, Bb, S, C, heads, NC, saved=scratch)
        report("fusion batched fwd fused B%d" % Bb, fused_b, f_ref, 1e-4)
        report("fusion batched fwd x B%d" % Bb, x_b, x_r, 1e-4)
        gt2 = {k: tuple(torch.zeros_like(t) for t in v) for k, v in names.items()}
        grads2 = ops.fusion_param_struct(gt2, cls=FusionGrads)
        dtok2 = ops.fusion_bwd(tk.detach(), prm, grads2, dfu, dxx, Bb, S, C, heads, NC, scratch=scratch)
        report("fusion batched bwd dtok B%d" % Bb, dtok2, tk.grad, 1e-3, rel=True)
        worst = max(float((gt2[k][d] - v[d].grad).abs().max() / v[d].grad.abs().max().clamp_min(1e-20))
                    for k, v in names.items() for d in range(2))
        report("fusion batched bwd all 26 parameter grads B%d (worst rel)" % Bb, torch.tensor([worst]), torch.zeros(1), 2e-3)


# ------------------------------------------------------------------------------------------------------------ misc
def t_ema():
    torch.manual_seed(6)
    sizes = [384, 1152 * 384, 7, 1536 * 384 + 3, 100000]
    ks = [torch.randn(n, device=dev) for n in sizes]
    qs = [torch.randn(n, device=dev) for n in sizes]
    m = 0.99
    ref = [k * m + q * (1. - m) for k, q in zip(ks, qs)]
    chunks, n, mx = ops.make_ema_chunks(list(zip(ks, qs)), dev)
    ops.ema_update_(chunks, n, mx, m)
    bad = sum(int((a != b).sum().item()) for a, b in zip(ks, ref))
    RESULTS.append(("ema bit-exact", bad == 0))
    print("%-52s %s mismatching elements=%d (chunks=%d)" % ("ema bit-exact vs eager k*m+q*(1.-m)", "OK  " if bad == 0 else "FAIL", bad, n))
    flat_k = torch.randn(5_000_000, device=dev)
    flat_q = torch.randn(5_000_000, device=dev)
    r = flat_k * 0.996 + flat_q * (1. - 0.996)
    chunks, n, mx = ops.make_ema_chunks([(flat_k[:3_000_000], flat_q[:3_000_000]), (flat_k[3_000_000:], flat_q[3_000_000:])], dev)
    ops.ema_update_(chunks, n, mx, 0.996)
    bad = int((flat_k != r).sum().item())
    RESULTS.append(("ema bit-exact flat", bad == 0))
    print("%-52s %s mismatching elements=%d (merged chunks=%d)" % ("ema bit-exact flat", "OK  " if bad == 0 else "FAIL", bad, n))


def t_infonce():
    from oracle.moco_ref import infonce_logits
    torch.manual_seed(7)
    N, D, K, T = 128, 256, 65536, 0.2
    q = torch.randn(N, D, device=dev, requires_grad=True)
    k = torch.randn(N, D, device=dev)
    queue = F.normalize(torch.randn(D, K, device=dev), dim=0)
    torch.backends.cuda.matmul.allow_tf32 = False
    logits_ref, labels, qn_ref, kn_ref = infonce_logits(q, k, queue, T)
    loss_ref = F.cross_entropy(logits_ref, labels)
    loss_ref.backward()
    qn, kn, logits, lse, loss = ops.infonce_fwd(q.detach(), k, queue, T)
    report("infonce qn", qn, qn_ref, 1e-6)
    report("infonce logits", logits, logits_ref, 1e-4)
    report("infonce lse", lse[:N], torch.logsumexp(logits_ref, 1), 1e-4)
    report("infonce loss", loss, loss_ref.reshape(1), 1e-4)
    dq = ops.infonce_bwd(q.detach(), qn, kn, queue, logits, lse, T)
    report("infonce dq (fused CE)", dq, q.grad, 2e-3, rel=True)
    dl = torch.randn_like(logits_ref) * 1e-3
    q2 = q.detach().clone().requires_grad_(True)
    l2, _, _, _ = infonce_logits(q2, k, queue, T)
    (l2 * dl).sum().backward()
    dq2 = ops.infonce_bwd(q.detach(), qn, kn, queue, logits, lse, T, dlogits=dl)
    report("infonce dq (external dlogits)", dq2, q2.grad, 2e-3, rel=True)
    # ---- tensor-core path (fp16 operands, fp32 accumulation): north_star bound on logits 2e-3 abs
    q16 = ops.queue16_update_(queue, torch.empty(D, K, device=dev, dtype=torch.float16))
    report("queue fp16 shadow", q16.float(), queue, 2.5e-4)
    qn_t, kn_t, buf, lse_t, loss_t = ops.infonce_tc_fwd(q.detach(), k, q16, T)
    report("infonce tc qn", qn_t, qn_ref, 1e-6)
    report("infonce tc logits", buf[:, 7:], logits_ref, 2e-3)
    report("infonce tc lse", lse_t[:N], torch.logsumexp(logits_ref, 1), 1e-3)
    report("infonce tc loss", loss_t, loss_ref.reshape(1), 1e-3)
    dq_t = ops.infonce_tc_bwd(q.detach(), qn_t, kn_t, q16, buf, lse_t, T)
    report("infonce tc dq (fused CE)", dq_t, q.grad, 5e-3, rel=True)
    dbuf = torch.zeros_like(buf)
    dbuf[:, 7:] = dl
    dq2_t = ops.infonce_tc_bwd(q.detach(), qn_t, kn_t, q16, buf, lse_t, T, dlogits_buf=dbuf)
    report("infonce tc dq (external dlogits)", dq2_t, q2.grad, 5e-3, rel=True)
    # backward after an enqueue overwrote 256 columns: the saved fp32 columns must restore the forward's queue
    old = queue[:, 4096:4096 + 256].clone()
    q16b = q16.clone()
    q16b[:, 4096:4096 + 256] = 0
    dq3_t = ops.infonce_tc_bwd(q.detach(), qn_t, kn_t, q16b, buf, lse_t, T, override=old, ov_start=4096)
    report("infonce tc dq with overwritten columns", dq3_t, q.grad, 5e-3, rel=True)
    keys = torch.randn(256, D, device=dev)
    qq = queue.clone()
    ops.enqueue_keys_(keys, qq, 1024)
    ref = queue.clone()
    ref[:, 1024:1024 + 256] = keys.t()
    report("enqueue keys", qq, ref, 0.0)


def t_small():
    torch.manual_seed(8)
    B, C, NC = 37, 384, 3
    a, b, c = [torch.randn(B, NC, device=dev, requires_grad=True) for _ in range(3)]
    tgt = torch.randint(0, NC, (B,), device=dev)
    lr = F.cross_entropy(a + b + c, tgt)
    lr.backward()
    loss, dl = ops.ce_small(a.detach(), b.detach(), c.detach(), tgt)
    report("ce_small loss", loss, lr.reshape(1), 1e-5)
    report("ce_small dlogits", dl, a.grad, 1e-6)
    x = torch.randn(B, 197, C, device=dev, requires_grad=True)
    lin = torch.nn.Linear(C, NC).to(dev)
    y_ref = lin(x[:, 0])
    y = ops.linear_small_fwd(x.detach(), 197 * C, lin.weight.detach(), lin.bias.detach(), B)
    report("linear_small fwd", y, y_ref, 1e-4)
    dy = torch.randn(B, NC, device=dev)
    y_ref.backward(dy)
    dx = torch.zeros_like(x)
    dw = torch.zeros_like(lin.weight)
    db = torch.zeros_like(lin.bias)
    ops.linear_small_bwd(x.detach(), 197 * C, lin.weight.detach(), dy, dx, 197 * C, dw, db, B)
    report("linear_small bwd dx", dx, x.grad, 1e-5)
    report("linear_small bwd dw", dw, lin.weight.grad, 1e-4)
    report("linear_small bwd db", db, lin.bias.grad, 1e-4)
    # column sums
    xb = bf(torch.randn(2, 3001, 1536, device=dev))
    out = torch.zeros(2, 1536, device=dev)
    ops.colsum_bf16(xb, out)
    report("colsum bf16", out, xb.float().sum(1), 1e-3, rel=True)
    # cast / optimisers
    p = torch.randn(100003, device=dev)
    report("cast bf16", ops.cast_bf16(p[:100000]), p[:100000], 1e-2, rel=True)
    p0 = torch.randn(70001, device=dev)
    g0 = torch.randn(70001, device=dev)
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([pt], lr=0.1, momentum=0.9, weight_decay=1e-4)
    pm = p0.clone()
    buf = torch.zeros_like(pm)
    sh = torch.empty_like(pm, dtype=torch.bfloat16)
    for it in range(3):
        pt.grad = g0 * (it + 1)
        opt.step()
        ops.sgd_step_(pm, g0 * (it + 1), buf, sh, 0.1, 0.9, 1e-4, it == 0)
    report("sgd 3 steps", pm, pt.detach(), 1e-5)
    report("sgd shadow", sh, pm, 1e-2, rel=True)
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pt], lr=1e-3, weight_decay=0.1)
    pm = p0.clone()
    m1 = torch.zeros_like(pm)
    m2 = torch.zeros_like(pm)
    for it in range(3):
        pt.grad = g0 * (it + 1)
        opt.step()
        ops.adam_step_(pm, g0 * (it + 1), m1, m2, None, 1e-3, (0.9, 0.999), 1e-8, 0.1, True, it + 1)
    report("adamw 3 steps", pm, pt.detach(), 1e-5)


def main():
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    print("device:", torch.cuda.get_device_name(0), flush=True)
    run("gemm_fwd small", t_gemm_fwd(1, 128, 64, 64, 64), flt)
    run("gemm_fwd small128", t_gemm_fwd(1, 256, 128, 128, 128), flt)
    run("gemm_fwd qkv", t_gemm_fwd(2, 6304, 1152, 384), flt)
    run("gemm_fwd qkv bn256", t_gemm_fwd(2, 6304, 1536, 384, 256), flt)
    run("gemm_fwd ragged", t_gemm_fwd(2, 197 * 3, 384, 1536, 128), flt)
    run("gemm_fwd wide384", t_gemm_fwd(2, 6304, 384, 1536, 384), flt)
    run("gemm_fwd wide384 ragged", t_gemm_fwd(2, 197 * 3, 384, 384, 384), flt)
    run("gemm_fwd N768 bn384", t_gemm_fwd(1, 1000, 768, 128, 384), flt)
    run("gemm epilogues", t_gemm_epilogues, flt)
    run("gemm rows96", t_gemm_rows96, flt)
    run("gemm resid+LN", t_gemm_resid_ln, flt)
    run("gemm stream-K", t_gemm_streamk, flt)
    run("gemm multicast", t_gemm_multicast, flt)
    run("patch embed TMA", t_patch_embed_tma, flt)
    run("wgrad pair", t_wgrad_pair, flt)
    run("gemm wgrad small", t_gemm_wgrad(1, 256, 128, 128, 1), flt)
    run("gemm wgrad", t_gemm_wgrad(2, 6304, 1152, 384, 8), flt)
    run("gemm wgrad fc", t_gemm_wgrad(2, 1970, 384, 1536, 5), flt)
    run("gemm wgrad fc1", t_gemm_wgrad(2, 6304, 1536, 384, 6), flt)
    run("layernorm", t_ln, flt)
    run("attn 197/64", t_attn(4, 197, 6, 64), flt)
    run("attn 577/64", t_attn(2, 577, 6, 64), flt)
    run("attn 197/32", t_attn(2, 197, 12, 32), flt)
    run("attn 50/64", t_attn(3, 50, 2, 64), flt)
    run("attn 197/64 f16", t_attn(4, 197, 6, 64, True), flt)
    run("attn 577/32 f16", t_attn(1, 577, 12, 32, True), flt)
    run("attn 577/64 f16", t_attn(2, 577, 6, 64, True), flt)
    run("attn 300/64", t_attn(2, 300, 3, 64), flt)
    # persistent forward: several items per CTA (64 x 6 = 384 items on 148 CTAs), tile-edge token counts
    run("attn 197/64 f16 NB64", t_attn(64, 197, 6, 64, True), flt)
    run("attn 197/64 NB50", t_attn(50, 197, 6, 64), flt)
    run("attn 256/64 f16", t_attn(30, 256, 6, 64, True), flt)
    run("attn 128/64", t_attn(70, 128, 3, 64), flt)
    run("attn 129/64 f16", t_attn(70, 129, 3, 64, True), flt)
    run("attn 17/64", t_attn(200, 17, 2, 64), flt)
    run("attn 197/64 f16 wide scores", t_attn(8, 197, 6, 64, True, amp=3.0, check_bwd=False), flt)
    run("attn 197/64 wide scores", t_attn(8, 197, 6, 64, False, amp=3.0), flt)
    run("attn 577/64 wide scores", t_attn(2, 577, 6, 64, False, amp=3.0), flt)
    run("fusion", t_fusion, flt)
    run("ema", t_ema, flt)
    run("infonce", t_infonce, flt)
    run("small ops", t_small, flt)
    bad = [n for n, ok in RESULTS if not ok]
    print("\n%d checks, %d failed" % (len(RESULTS), len(bad)))
    for n in bad:
        print("  FAILED:", n)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
