"""Operand delivery at full occupancy: mainloop time per k-block with both operands loaded, only the shared one (B: weights,
the same tile for every row tile of a group) or only the private one (A), 2 ... 74 pair tiles of 256 x 384 / 256 x 256."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-feature-vit_b200"))
from mfvit import ops
from mfvit._lib import EPI_BF16
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
G = 2
for N, bn in ((384, 384), (256, 256)):
    for tiles in (2, 26, 50, 74):
        M = tiles // 2 * 256
        t = {}
        for K in (768, 1536):
            A = torch.randn(G, M, K, device=dev).bfloat16(); Bm = torch.randn(G, N, K, device=dev).bfloat16()
            out = torch.zeros(G, M, N, device=dev, dtype=torch.bfloat16)
            for md in (0, 1, 13, 14):
                t[(K, md)] = timeit(lambda: ops.gemm(A, Bm, out, M=M, N=N, K=K, G=G, lda=K, ldb=K, ldc=N, a_gstride=M * K, b_gstride=N * K,
                    c_gstride=M * N, epilogue=EPI_BF16, block_n=bn, dtype_flags=256 | (md << 16)))
        per = {md: (t[(1536, md)] - t[(768, md)]) / 12 * 1965 for md in (0, 1, 13, 14)}
        print("N%d tiles %2d: clk per k-block: A+B %.0f | none %.0f | B only %.0f | A only %.0f" % (N, tiles, per[0], per[1], per[13], per[14]), flush=True)
