"""Where the fused GEMM epilogues spend their time: clock64 sums per epilogue warp and phase (dtype_flags bit 20, buffer in
row_sum), averaged over the warps of all CTAs, at the step's shapes.  python tests/gpu_epi_prof.py [pairs=32]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-feature-vit_b200"))
from mfvit import ops
from mfvit._lib import EPI_BF16, EPI_GELU, EPI_DGELU, EPI_RESID_LN
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
M = B * 197
PH = ["acc wait", "tmem ld", "slot/aux wait", "convert+sts", "fence+store", "math", "ln barrier", "drain"]
def h(*s): return (torch.randn(*s, device=dev) * 0.3).half()
def bf(*s): return (torch.randn(*s, device=dev) * 0.3).bfloat16()
def f32(*s): return torch.randn(*s, device=dev)
prof = torch.zeros(148 * 17 * 8, device=dev, dtype=torch.int64)
def report(name, fn):
    for _ in range(3): fn(0)
    prof.zero_(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(1 << 20); b.record(); torch.cuda.synchronize()
    pall = prof.view(148, 17, 8).double()
    p = pall[:, :16]
    mm = pall[:, 16]
    mu = mm[:, 2] > 0
    issuer = ("  || issuer per tile: wait for accumulator %.0f, k-loop %.0f clk (%.1f tiles per pair)"
              % (float(mm[mu, 0].sum() / mm[mu, 2].sum()), float(mm[mu, 1].sum() / mm[mu, 2].sum()), float(mm[mu, 2].mean()))) if mu.any() else ""
    used = p.sum(2) > 0
    n = int(used.sum())
    mean = (p * used[..., None]).sum((0, 1)) / max(n, 1)
    tot = p.sum(2)
    print("%-18s %6.1f us | warps %4d | total/warp mean %7.0f max %7.0f clk | " % (name, a.elapsed_time(b) * 1e3, n, float(tot.sum() / max(n, 1)), float(tot.max()))
          + "  ".join("%s %.0f" % (PH[k], float(mean[k])) for k in range(8)) + issuer, flush=True)
x, w, b = h(2, M, 384), h(2, 1536, 384), f32(2, 1536)
u, g16 = torch.empty(2, M, 1536, device=dev, dtype=torch.bfloat16), torch.empty(2, M, 1536, device=dev, dtype=torch.float16)
def gemm_kw(**kw): return kw
report("fc1+GELU", lambda fl: ops.gemm(x, w, u, M=M, N=1536, K=384, G=2, lda=384, ldb=384, ldc=1536, a_gstride=M * 384, b_gstride=1536 * 384,
       c_gstride=M * 1536, bias=b, bias_gstride=1536, C2=g16, epilogue=EPI_GELU, dtype_flags=7 | fl, row_sum=prof if fl else None))
w3, b3 = h(2, 1152, 384), f32(2, 1152)
q = torch.empty(2, M, 1152, device=dev, dtype=torch.float16)
report("qkv", lambda fl: ops.gemm(x, w3, q, M=M, N=1152, K=384, G=2, lda=384, ldb=384, ldc=1152, a_gstride=M * 384, b_gstride=1152 * 384,
       c_gstride=M * 1152, bias=b3, bias_gstride=1152, epilogue=EPI_BF16, dtype_flags=7 | fl, row_sum=prof if fl else None))
for tag, K in (("fc2+LN", 1536), ("proj+LN", 384)):
    xa, wa, ba = h(2, M, K), h(2, 384, K), f32(2, 384)
    res, gam, bet = f32(2, M, 384), f32(2, 384), f32(2, 384)
    xn = torch.empty(2, M, 384, device=dev); y = torch.empty(2, M, 384, device=dev, dtype=torch.float16)
    yc = torch.empty(2, M, 384, device=dev, dtype=torch.bfloat16); mean = torch.empty(2, M, device=dev); rstd = torch.empty(2, M, device=dev)
    report(tag, lambda fl: ops.gemm(xa, wa, xn, M=M, N=384, K=K, G=2, lda=K, ldb=K, ldc=384, a_gstride=M * K, b_gstride=384 * K,
           c_gstride=M * 384, bias=ba, bias_gstride=384, aux=res, aux_ld=384, aux_gstride=M * 384, C2=y, C3=yc, epilogue=EPI_RESID_LN,
           dtype_flags=7 | fl, row_sum=prof if fl else None, ln=dict(gamma=gam, beta=bet, mean=mean, rstd=rstd, eps=1e-6, out_f32=False)))
dy, w2 = bf(2, M, 384), bf(2, 384, 1536)
uu = bf(2, M, 1536); du = torch.empty_like(uu); gg = torch.empty_like(uu)
report("fc2 dgrad+GELU'", lambda fl: ops.gemm(dy, w2, du, M=M, N=1536, K=384, G=2, lda=384, ldb=1536, ldc=1536, a_gstride=M * 384,
       b_gstride=384 * 1536, c_gstride=M * 1536, aux=uu, aux_ld=1536, aux_gstride=M * 1536, b_mn=True, epilogue=EPI_DGELU, C2=gg, dtype_flags=fl,
       row_sum=prof if fl else None))
dyy, ww = bf(2, M, 1536), bf(2, 1536, 384)
dx = torch.empty(2, M, 384, device=dev, dtype=torch.bfloat16)
report("fc1 dgrad", lambda fl: ops.gemm(dyy, ww, dx, M=M, N=384, K=1536, G=2, lda=1536, ldb=384, ldc=384, a_gstride=M * 1536, b_gstride=1536 * 384,
       c_gstride=M * 384, b_mn=True, epilogue=EPI_BF16, dtype_flags=fl, row_sum=prof if fl else None))
