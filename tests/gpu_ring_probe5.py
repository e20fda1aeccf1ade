"""UMMA rate against operand majorness (stale shared memory, no TMA loads): 256 x 384 pair tiles, K-major / MN-major A and B."""
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
    for tiles in (2, 74):
        M = tiles // 2 * 256
        for a_mn in (False, True):
            for b_mn in (False, True):
                t = {}
                for K in (1536, 3072):
                    A = torch.randn(G, K, M, device=dev).bfloat16() if a_mn else torch.randn(G, M, K, device=dev).bfloat16()
                    Bm = torch.randn(G, K, N, device=dev).bfloat16() if b_mn else torch.randn(G, N, K, device=dev).bfloat16()
                    out = torch.zeros(G, M, N, device=dev, dtype=torch.bfloat16)
                    for md in (0, 1):
                        t[(K, md)] = timeit(lambda: ops.gemm(A, Bm, out, M=M, N=N, K=K, G=G, lda=M if a_mn else K, ldb=N if b_mn else K, ldc=N,
                            a_gstride=M * K, b_gstride=N * K, c_gstride=M * N, a_mn=a_mn, b_mn=b_mn, epilogue=EPI_BF16, block_n=bn,
                            dtype_flags=256 | (md << 16)))
                per = [(t[(3072, md)] - t[(1536, md)]) / 24 * 1965 for md in (0, 1)]
                print("N%d tiles %2d A %s B %s: clk per k-block with loads %.0f, without %.0f" % (N, tiles, "MN" if a_mn else "K ", "MN" if b_mn else "K ", per[0], per[1]), flush=True)
