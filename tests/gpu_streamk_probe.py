"""Stream-K cost split (profiles/r02_summary.md): the fc2-shaped bf16 GEMM with stream-K off / on, and with the partial
dump (dtype_flags 2048) and / or the partial add (1024) suppressed - wrong results, timing only.
    python tests/gpu_streamk_probe.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-feature-vit_b200"))
from mfvit import ops, _lib
from mfvit._lib import EPI_BF16, EPI_RESID_F32
lib = _lib.load()
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
for B in (32, 64):
    M = B * 197; N = 384; K = 1536
    x = (torch.randn(2, M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(2, N, K, device=dev) * 0.05).bfloat16()
    o16 = torch.zeros(2, M, N, device=dev, dtype=torch.bfloat16)
    for sk in (0, 1):
        lib.mfv_set_option(b"streamk", sk)
        r = []
        for fl, nm in ((0, "full"), (256, "noepi"), (512, "nostore"), (1024, "noadd"), (2048, "nodump"), (3072, "noadd+nodump")):
            r.append("%s %.1f" % (nm, timeit(lambda: ops.linear_fwd(x, w, None, EPI_BF16, out=o16, block_n=384, dtype_flags=fl))))
        print("B%d sk%d bf16 fc2: %s" % (B, sk, " | ".join(r)), flush=True)
lib.mfv_set_option(b"streamk", 0)
