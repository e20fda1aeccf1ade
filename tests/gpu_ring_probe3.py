"""UMMA shape probe: mainloop without TMA loads (stale shared memory), single round of tiles."""
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
for N, bn, modes in ((384, 384, tuple(range(13))),):
    for K in (1536, 3072):
        for tiles in (2,):
            M = tiles // 2 * 256
            x = (torch.randn(2, M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(2, N, K, device=dev) * 0.05).bfloat16()
            o16 = torch.zeros(2, M, N, device=dev, dtype=torch.bfloat16)
            r = []
            for md in modes:
                t = timeit(lambda: ops.linear_fwd(x, w, None, EPI_BF16, out=o16, block_n=bn, dtype_flags=256 | (md << 16)))
                r.append("m%d:%.1f" % (md, t))
            print("N%d K%d tiles %d noepi us: %s" % (N, K, tiles, " ".join(r)), flush=True)
