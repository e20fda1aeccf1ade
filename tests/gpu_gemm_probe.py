"""GEMM probes behind profiles/r02_summary.md section 3 (graph-timed, one B200): the 384-wide shapes of the step against
cuBLAS on the bare contraction, with / without the epilogue (dtype_flags 256 = accumulators released unread) and with 128 / 96
rows per CTA.  Toggle MFVIT_STREAMK / MFVIT_GEMM_MC in the environment for the stream-K and multicast variants.
    python tests/gpu_gemm_probe.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-feature-vit_b200"))
from mfvit import ops
from mfvit._lib import EPI_BF16, EPI_RESID_F32
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
    for tag, N, K in (("fc2 ", 384, 1536), ("proj", 384, 384), ("qkvd", 384, 1152)):
        x = (torch.randn(2, M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(2, N, K, device=dev) * 0.05).bfloat16()
        b = torch.randn(2, N, device=dev); aux = torch.randn(2, M, N, device=dev)
        out = torch.zeros(2, M, N, device=dev); o16 = torch.zeros(2, M, N, device=dev, dtype=torch.bfloat16)
        c = torch.empty(2, M, N, device=dev, dtype=torch.bfloat16)
        wt = w.transpose(1, 2).contiguous()
        cub = timeit(lambda: torch.bmm(x, wt, out=c))
        r = []
        for rpc in (128, 96):
            full = timeit(lambda: ops.linear_fwd(x, w, b, EPI_RESID_F32, out=out, aux=aux, block_n=384, rows_per_cta=rpc))
            noepi = timeit(lambda: ops.linear_fwd(x, w, b, EPI_RESID_F32, out=out, aux=aux, block_n=384, rows_per_cta=rpc, dtype_flags=256))
            b16 = timeit(lambda: ops.linear_fwd(x, w, None, EPI_BF16, out=o16, block_n=384, rows_per_cta=rpc))
            r.append("rpc%d resid %.1f noepi %.1f bf16 %.1f" % (rpc, full, noepi, b16))
        b256 = timeit(lambda: ops.linear_fwd(x, w, None, EPI_BF16, out=o16, block_n=128))
        print("B%d %s M%d N%d K%d: cuBLAS %.1f | %s | bn128 bf16 %.1f" % (B, tag, M, N, K, cub, " | ".join(r), b256), flush=True)
