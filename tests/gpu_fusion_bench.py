"""Fusion kernels micro-benchmark (graph-timed): mfv_fusion_fwd / mfv_fusion_bwd at B pairs. usage: gpu_fusion_bench.py [B=32]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "multi-feature-vit_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from mfvit import ops
from mfvit._lib import FusionGrads
from mfvit.functions import FUSION_FIELDS
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S, C, heads, NC = 197, 384, 3, 3
torch.manual_seed(0)
shapes = {"ln1_w": (C,), "ln1_b": (C,), "wq": (C, C), "wk": (C, C), "wv": (C, C), "proj_w": (C, C), "proj_b": (C,),
          "ln2_w": (C,), "ln2_b": (C,), "head_w": (NC, C), "head_b": (NC,), "vhead_w": (NC, C), "vhead_b": (NC,)}
params = {n: (torch.randn(*shapes[n], device=dev) * 0.05, torch.randn(*shapes[n], device=dev) * 0.05) for n in FUSION_FIELDS}
grads = {n: (torch.zeros_like(a), torch.zeros_like(b)) for n, (a, b) in params.items()}
ps = ops.fusion_param_struct(params)
gs = ops.fusion_param_struct(grads, cls=FusionGrads)
tok = torch.randn(2, B, S, C, device=dev)
dlog = torch.randn(B, NC, device=dev)
dx = torch.randn(2, B, NC, device=dev)
dtok = torch.empty_like(tok)


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


print(torch.cuda.get_device_name(0), "B", B)
print("fusion fwd (single kernel): %.1f us" % (timeit(lambda: ops.fusion_fwd(tok, ps, B, S, C, heads, NC)) * 1e3))
scr = ops.fusion_scratch(tok, B, S, C, heads)
print("fusion fwd (batched stages): %.1f us" % (timeit(lambda: ops.fusion_fwd(tok, ps, B, S, C, heads, NC, saved=scr)) * 1e3))


def fwd_bwd():
    ops.fusion_fwd(tok, ps, B, S, C, heads, NC, saved=scr)
    ops.fusion_bwd(tok, ps, gs, dlog, dx, B, S, C, heads, NC, dtok=dtok, scratch=scr)


print("fusion fwd + bwd (+wgrad), batched, state reused: %.1f us" % (timeit(fwd_bwd) * 1e3))
print("fusion bwd (+wgrad): %.1f us" % (timeit(lambda: ops.fusion_bwd(tok, ps, gs, dlog, dx, B, S, C, heads, NC, dtok=dtok)) * 1e3))
