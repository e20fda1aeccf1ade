"""Operand-ring depth probe (mainloop only: dtype_flags bit 8 releases the accumulators unread, bits 12..15 set the number
of operand stages - stages past the kernel's own count lie over the unused epilogue rings).  Graph-timed, one B200.
    python tests/gpu_ring_probe.py"""
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
for B in (32, 64):
    M = B * 197
    for tag, N, K, bn in (("fc2 ", 384, 1536, 384), ("qkvd", 384, 1152, 384), ("proj", 384, 384, 384), ("fc1 ", 1536, 384, 256),
                          ("qkv ", 1152, 384, 256), ("big ", 1536, 1536, 256)):
        x = (torch.randn(2, M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(2, N, K, device=dev) * 0.05).bfloat16()
        o16 = torch.zeros(2, M, N, device=dev, dtype=torch.bfloat16)
        wt = w.transpose(1, 2).contiguous()
        cub = timeit(lambda: torch.bmm(x, wt, out=o16))
        r = []
        for st in (0,):
            t = timeit(lambda: ops.linear_fwd(x, w, None, EPI_BF16, out=o16, block_n=bn, dtype_flags=256 | (st << 12)))
            r.append("st%d %.1f" % (st, t))
        full = timeit(lambda: ops.linear_fwd(x, w, None, EPI_BF16, out=o16, block_n=bn))
        print("B%d %s M%d N%d K%d bn%d: cuBLAS %.1f | noepi %s | bf16 epi %.1f" % (B, tag, M, N, K, bn, cub, " ".join(r), full), flush=True)
