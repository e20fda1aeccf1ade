"""A/B: 384-wide pair tiles with 128 vs 96 rows per CTA on the step's N=384 GEMMs (graph-timed) + correctness vs fp32."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi-feature-vit_b200"))
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
    return best


B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
M = B * 197
print(torch.cuda.get_device_name(0), "pairs", B)
for tag, N, K, epi in (("proj fwd ", 384, 384, EPI_RESID_F32), ("fc2 fwd  ", 384, 1536, EPI_RESID_F32)):
    x = (torch.randn(2, M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(2, N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(2, N, device=dev); aux = torch.randn(2, M, N, device=dev)
    ref = torch.einsum("gmk,gnk->gmn", x.float(), w.float()) + b[:, None, :] + aux
    res = {}
    for rpc in (128, 96):
        out = torch.zeros(2, M, N, device=dev)
        ops.linear_fwd(x, w, b, epi, out=out, aux=aux, block_n=384, rows_per_cta=rpc)
        torch.cuda.synchronize()
        err = (out - ref).abs().max().item()
        ms = timeit(lambda: ops.linear_fwd(x, w, b, epi, out=out, aux=aux, block_n=384, rows_per_cta=rpc))
        res[rpc] = ms
        print("%s rows/CTA %3d: %6.1f us  max err %.2e" % (tag, rpc, ms * 1e3, err), flush=True)
for tag, N, K in (("fc1 dgrad", 1536, 384), ("qkv dgrad", 1152, 384), ("proj dgrd", 384, 384)):
    dy = (torch.randn(2, M, N, device=dev) * 0.5).bfloat16(); w = (torch.randn(2, N, K, device=dev) * 0.05).bfloat16()
    ref = torch.einsum("gmn,gnk->gmk", dy.float(), w.float())
    for rpc in (128, 96):
        out = torch.zeros(2, M, K, device=dev, dtype=torch.bfloat16)
        ops.linear_dgrad(dy, w, EPI_BF16, out=out, block_n=384, rows_per_cta=rpc)
        torch.cuda.synchronize()
        err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
        ms = timeit(lambda: ops.linear_dgrad(dy, w, EPI_BF16, out=out, block_n=384, rows_per_cta=rpc))
        print("%s rows/CTA %3d: %6.1f us  rel err %.2e" % (tag, rpc, ms * 1e3, err), flush=True)
