"""Tile width against round quantisation for the wide GEMMs of the step (fc1 + GELU, qkv, fc2 dgrad + GELU'): block_n 256 / 128."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-feature-vit_b200"))
from mfvit import ops
from mfvit._lib import EPI_BF16, EPI_GELU, EPI_DGELU
dev = "cuda"
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
for B in (32, 64):
    M = B * 197
    x, w, b = h(2, M, 384), h(2, 1536, 384), f32(2, 1536)
    u, g16 = torch.empty(2, M, 1536, device=dev, dtype=torch.bfloat16), torch.empty(2, M, 1536, device=dev, dtype=torch.float16)
    w3, b3 = h(2, 1152, 384), f32(2, 1152)
    q = torch.empty(2, M, 1152, device=dev, dtype=torch.float16)
    dy, w2 = bf(2, M, 384), bf(2, 384, 1536)
    uu = bf(2, M, 1536); du = torch.empty_like(uu); gg = torch.empty_like(uu)
    for bn in (256, 128):
        t1 = timeit(lambda: ops.gemm(x, w, u, M=M, N=1536, K=384, G=2, lda=384, ldb=384, ldc=1536, a_gstride=M * 384, b_gstride=1536 * 384,
             c_gstride=M * 1536, bias=b, bias_gstride=1536, C2=g16, epilogue=EPI_GELU, dtype_flags=7, block_n=bn))
        t2 = timeit(lambda: ops.gemm(x, w3, q, M=M, N=1152, K=384, G=2, lda=384, ldb=384, ldc=1152, a_gstride=M * 384, b_gstride=1152 * 384,
             c_gstride=M * 1152, bias=b3, bias_gstride=1152, epilogue=EPI_BF16, dtype_flags=7, block_n=bn))
        t3 = timeit(lambda: ops.gemm(dy, w2, du, M=M, N=1536, K=384, G=2, lda=384, ldb=1536, ldc=1536, a_gstride=M * 384,
             b_gstride=384 * 1536, c_gstride=M * 1536, aux=uu, aux_ld=1536, aux_gstride=M * 1536, b_mn=True, epilogue=EPI_DGELU, C2=gg, block_n=bn))
        print("B%d block_n %d: fc1+GELU %.1f  qkv %.1f  fc2 dgrad+GELU' %.1f us" % (B, bn, t1, t2, t3), flush=True)
