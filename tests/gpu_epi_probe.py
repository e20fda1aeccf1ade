"""Epilogue probe at the step's shapes (32 pairs: 2 x 6304 rows): each fused-epilogue GEMM with the whole epilogue, with the
bulk stores skipped (dtype_flags bit 9) and with the accumulators released unread (bit 8).  Graph-timed, fp16 forward mode.
    python tests/gpu_epi_probe.py [pairs=32]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-feature-vit_b200"))
from mfvit import ops
from mfvit._lib import EPI_BF16, EPI_GELU, EPI_DGELU, EPI_RESID_LN
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
M = B * 197
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
    return best * 1e3
def h(*s): return (torch.randn(*s, device=dev) * 0.3).half()
def bf(*s): return (torch.randn(*s, device=dev) * 0.3).bfloat16()
def f32(*s): return torch.randn(*s, device=dev)
rows = []
# fc1 + GELU: u bf16, gelu fp16, gelu bf16 twin
x, w, b = h(2, M, 384), h(2, 1536, 384), f32(2, 1536)
u, g16, gb = torch.empty(2, M, 1536, device=dev, dtype=torch.bfloat16), torch.empty(2, M, 1536, device=dev, dtype=torch.float16), torch.empty(2, M, 1536, device=dev, dtype=torch.bfloat16)
for name, fl in (("full", 0), ("nostore", 512), ("noepi", 256)):
    rows.append(("fc1+GELU " + name, timeit(lambda: ops.linear_fwd(x, w, b, EPI_GELU, out=u, out2=g16, out3=gb, dtype_flags=7 | fl))))
# qkv bf16 epilogue (fp16 out)
w3, b3 = h(2, 1152, 384), f32(2, 1152)
q = torch.empty(2, M, 1152, device=dev, dtype=torch.float16)
for name, fl in (("full", 0), ("nostore", 512), ("noepi", 256)):
    rows.append(("qkv " + name, timeit(lambda: ops.linear_fwd(x, w3, b3, EPI_BF16, out=q, dtype_flags=7 | fl))))
# fc2 + residual + LayerNorm (K = 1536) and proj + residual + LayerNorm (K = 384)
for tag, K in (("fc2+LN", 1536), ("proj+LN", 384)):
    xa, wa, ba = h(2, M, K), h(2, 384, K), f32(2, 384)
    res, gam, bet = f32(2, M, 384), f32(2, 384), f32(2, 384)
    xn = torch.empty(2, M, 384, device=dev); y = torch.empty(2, M, 384, device=dev, dtype=torch.float16)
    yc = torch.empty(2, M, 384, device=dev, dtype=torch.bfloat16); mean = torch.empty(2, M, device=dev); rstd = torch.empty(2, M, device=dev)
    for name, fl in (("full", 0), ("nostore", 512), ("noepi", 256)):
        rows.append((tag + " " + name, timeit(lambda: ops.gemm(xa, wa, xn, M=M, N=384, K=K, G=2, lda=K, ldb=K, ldc=384, a_gstride=M * K, b_gstride=384 * K,
            c_gstride=M * 384, bias=ba, bias_gstride=384, aux=res, aux_ld=384, aux_gstride=M * 384, C2=y, C3=yc, epilogue=EPI_RESID_LN,
            dtype_flags=7 | fl, ln=dict(gamma=gam, beta=bet, mean=mean, rstd=rstd, eps=1e-6, out_f32=False)))))
# fc2 dgrad + GELU' (bf16): dy [M,384] x W2 [384,1536] -> du = . * gelu'(u), C2 = gelu(u) bf16
dy, w2 = bf(2, M, 384), bf(2, 384, 1536)
uu = bf(2, M, 1536); du = torch.empty_like(uu); gg = torch.empty_like(uu)
for name, fl in (("full", 0), ("nostore", 512), ("noepi", 256)):
    rows.append(("fc2 dgrad+GELU' " + name, timeit(lambda: ops.gemm(dy, w2, du, M=M, N=1536, K=384, G=2, lda=384, ldb=1536, ldc=1536, a_gstride=M * 384,
        b_gstride=384 * 1536, c_gstride=M * 1536, aux=uu, aux_ld=1536, aux_gstride=M * 1536, b_mn=True, epilogue=EPI_DGELU, C2=gg, dtype_flags=fl))))
# 384-wide bf16 dgrads (fc1 dgrad K = 1536, qkv dgrad K = 1152, proj dgrad K = 384)
for tag, K in (("fc1 dgrad", 1536), ("qkv dgrad", 1152), ("proj dgrad", 384)):
    dyy, ww = bf(2, M, K), bf(2, K, 384)
    dx = torch.empty(2, M, 384, device=dev, dtype=torch.bfloat16)
    for name, fl in (("full", 0), ("nostore", 512), ("noepi", 256)):
        rows.append((tag + " " + name, timeit(lambda: ops.gemm(dyy, ww, dx, M=M, N=384, K=K, G=2, lda=K, ldb=384, ldc=384, a_gstride=M * K, b_gstride=K * 384,
            c_gstride=M * 384, b_mn=True, epilogue=EPI_BF16, dtype_flags=fl))))
for n, t in rows:
    print("%-28s %7.1f us" % (n, t), flush=True)
