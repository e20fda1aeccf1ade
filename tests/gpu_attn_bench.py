"""Attention micro-benchmark (graph-timed): mfv_attn_fwd / mfv_attn_bwd at the step's shape. usage: gpu_attn_bench.py [NB=64] [S=197]"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi-feature-vit_b200"))
from mfvit import ops
dev = "cuda"
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 197
H, D = 6, 64


def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
    return best


print(torch.cuda.get_device_name(0), "NB", NB, "S", S)
for f16 in (False, True):
    qkv = torch.randn(NB, S, 3, H, D, device=dev)
    qkv = qkv.half() if f16 else qkv.bfloat16()
    fwd_flops = 4.0 * NB * H * S * S * D
    ms = timeit(lambda: ops.attn_fwd(qkv, H, f16=f16, bf16_copy=f16))
    print("attn fwd %s: %7.1f us  %6.1f TF/s" % ("f16" if f16 else "bf16", ms * 1e3, fwd_flops / ms / 1e9), flush=True)
    r = ops.attn_fwd(qkv, H, f16=f16, bf16_copy=f16)
    o, lse = (r[2], r[1]) if f16 else r
    do = torch.randn(NB, S, H, D, device=dev).bfloat16()
    ms = timeit(lambda: ops.attn_bwd(qkv, o, do, lse))
    print("attn bwd %s: %7.1f us  %6.1f TF/s (5 GEMMs incl. recompute)" % ("f16" if f16 else "bf16", ms * 1e3, 2.5 * fwd_flops / ms / 1e9), flush=True)
