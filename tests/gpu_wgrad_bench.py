"""The four weight-gradient GEMMs of a block at the step's shapes (32 pairs: 2 x 6304 tokens), graph-timed, against the
forward GEMM of the same FLOPs (bare mainloop) and cuBLAS.  python tests/gpu_wgrad_bench.py [pairs=32]"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi-feature-vit_b200"))
from mfvit import ops
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
G, M = 2, B * 197
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
for tag, N, K in (("qkv ", 1152, 384), ("proj", 384, 384), ("fc1 ", 1536, 384), ("fc2 ", 384, 1536)):
    dy = torch.randn(G, M, N, device=dev).bfloat16(); x = torch.randn(G, M, K, device=dev).bfloat16()
    dw = torch.zeros(G, N, K, device=dev); db = torch.zeros(G, N, device=dev)
    out = torch.empty(G, N, K, device=dev, dtype=torch.bfloat16)
    cub = timeit(lambda: torch.bmm(dy.transpose(1, 2), x, out=out))
    r = []
    for sp in (2, 3, 4, 6, 8, 12):
        r.append("sp%d %.1f" % (sp, timeit(lambda: ops.linear_wgrad(dy, x, dw, splits=sp, db=db if K == 384 else None))))
    fl = 2.0 * G * M * N * K
    print("%s dW[%d x %d] over %d tokens: cuBLAS %.1f us | %s | %.1f GF" % (tag, N, K, M, cub, " ".join(r), fl / 1e9), flush=True)
