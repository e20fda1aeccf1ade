"""Weight-gradient GEMMs of a block: whole kernel / bulk reduce-adds skipped / accumulators released unread, and the phase
clocks of the epilogue warps.  python tests/gpu_wgrad_probe2.py"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi-feature-vit_b200"))
from mfvit import ops
from mfvit._lib import EPI_ATOMIC_F32
dev = "cuda"
G, M = 2, 32 * 197
PH = ["acc wait", "tmem ld", "slot/aux wait", "convert+sts", "fence+store", "math", "ln barrier", "drain"]
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
for tag, N, K, sp in (("qkv ", 1152, 384, 6), ("proj", 384, 384, 12), ("fc1 ", 1536, 384, 6), ("fc2 ", 384, 1536, 3)):
    dy = torch.randn(G, M, N, device=dev).bfloat16(); x = torch.randn(G, M, K, device=dev).bfloat16()
    dw = torch.zeros(G, N, K, device=dev)
    def run(fl):
        return ops.gemm(dy, x, dw, M=N, N=K, K=M, G=G, lda=N, ldb=K, ldc=K, a_gstride=M * N, b_gstride=M * K, c_gstride=N * K,
                        a_mn=True, b_mn=True, epilogue=EPI_ATOMIC_F32, splits=sp, dtype_flags=fl)
    print("%s splits %d: full %.1f  nostore %.1f  noepi %.1f us" % (tag, sp, timeit(lambda: run(0)), timeit(lambda: run(512)), timeit(lambda: run(256))), flush=True)
# mainloop only, with and without the TMA loads (stale shared memory): what bounds the MN-major mainloop
for tag, N, K, sp in (("qkv ", 1152, 384, 6), ("fc1 ", 1536, 384, 6), ("fc2 ", 384, 1536, 3)):
    dy = torch.randn(G, M, N, device=dev).bfloat16(); x = torch.randn(G, M, K, device=dev).bfloat16()
    dw = torch.zeros(G, N, K, device=dev)
    def run(fl):
        return ops.gemm(dy, x, dw, M=N, N=K, K=M, G=G, lda=N, ldb=K, ldc=K, a_gstride=M * N, b_gstride=M * K, c_gstride=N * K,
                        a_mn=True, b_mn=True, epilogue=EPI_ATOMIC_F32, splits=sp, dtype_flags=fl)
    print("%s splits %d mainloop only: with loads %.1f  without %.1f us" % (tag, sp, timeit(lambda: run(256)), timeit(lambda: run(256 | (1 << 16)))), flush=True)
