import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi-feature-vit_b200"))
from mfvit import ops
from mfvit._lib import EPI_F32, EPI_ATOMIC_F32
dev = "cuda"
torch.manual_seed(0)
G, M, N, K = 1, 512, 256, 384
x = torch.randn(G, M, K, device=dev); w = torch.randn(G, N, K, device=dev) * 0.05
ref = torch.einsum("gmk,gnk->gmn", x, w)
def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item()
for fa, fb in ((0, 0), (1, 1), (0, 1), (1, 0)):
    xa = x.half() if fa else x.bfloat16(); wb = w.half() if fb else w.bfloat16()
    out = torch.empty(G, M, N, device=dev)
    ops.gemm(xa, wb, out, M=M, N=N, K=K, G=G, lda=K, ldb=K, ldc=N, epilogue=EPI_F32, dtype_flags=fa | (fb << 1))
    exact = torch.einsum("gmk,gnk->gmn", xa.float(), wb.float())
    print("A %s B %s: err vs exact-product-of-rounded-inputs %.3e   vs fp32 %.3e" % ("f16" if fa else "bf16", "f16" if fb else "bf16", rel(out, exact), rel(out, ref)))
# wgrad style (both MN-major), A bf16 (dy) B fp16 (x)
dy = torch.randn(G, M, N, device=dev).bfloat16(); xx = torch.randn(G, M, K, device=dev).half()
dw = torch.zeros(G, N, K, device=dev)
ops.gemm(dy, xx, dw, M=N, N=K, K=M, G=G, lda=N, ldb=K, ldc=K, a_mn=True, b_mn=True, epilogue=EPI_ATOMIC_F32, splits=2, block_n=128, dtype_flags=2)
print("wgrad mixed err %.3e" % rel(dw, torch.einsum("gmn,gmk->gnk", dy.float(), xx.float())))
