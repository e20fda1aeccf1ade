"""UMMA cost per instruction, cta_group::1 (M = 128) and ::2 (M = 256), N = 64 / 128 / 256: mainloop on stale shared memory."""
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
for cg in (1, 2):
    for bn in (64, 128, 256):
        if cg == 2 and bn == 64: continue
        N = bn
        M = 128 * cg
        t = {}
        for K in (1536, 3072):
            x = (torch.randn(2, M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(2, N, K, device=dev) * 0.05).bfloat16()
            o16 = torch.zeros(2, M, N, device=dev, dtype=torch.bfloat16)
            t[K] = timeit(lambda: ops.linear_fwd(x, w, None, EPI_BF16, out=o16, block_n=bn, cta_group=cg, dtype_flags=256 | (1 << 16)))
        per = (t[3072] - t[1536]) / (24 * 4) * 1965
        print("cta_group %d M %d N %d: %.1f / %.1f us -> %.0f clk per UMMA (K=16)" % (cg, M, N, t[1536], t[3072], per), flush=True)
