"""GEMM micro-benchmark: us and TFLOP/s of mfv_gemm on the block's shapes (CUDA events, 20 reps after 3 warm-ups).

`noepi` columns re-run the same launch with dtype_flags bit 8 (accumulators released unread): mainloop-only time.
usage: python tests/gpu_gemm_bench.py [pairs_per_gpu=32]
"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi-feature-vit_b200"))
from mfvit import ops
from mfvit._lib import EPI_BF16, EPI_F32, EPI_ATOMIC_F32, EPI_RESID_F32, EPI_GELU, EPI_DGELU
dev = "cuda"
NAMES = {EPI_BF16: "bf16", EPI_F32: "f32", EPI_RESID_F32: "resid", EPI_GELU: "gelu", EPI_DGELU: "dgelu"}


def timeit(fn, reps=20):
    """reps launches captured in one CUDA graph: device time per launch without Python / ctypes launch overhead"""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    g.replay(); g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):  # min of 5 graph replays: the box-to-box / clock-ramp noise is +-30 % on single replays
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
    return best


def fwd(G, M, N, K, bn=0, epi=EPI_BF16, cg=0, tag="fwd "):
    x = torch.randn(G, M, K, device=dev).bfloat16(); w = torch.randn(G, N, K, device=dev).bfloat16()
    out = torch.empty(G, M, N, device=dev, dtype=torch.float32 if epi in (EPI_F32, EPI_RESID_F32) else torch.bfloat16)
    out2 = torch.empty_like(out) if epi == EPI_GELU else None
    aux = torch.randn(G, M, N, device=dev) if epi == EPI_RESID_F32 else None
    b = torch.randn(G, N, device=dev)
    ms = timeit(lambda: ops.linear_fwd(x, w, b, epi, out=out, out2=out2, aux=aux, block_n=bn, cta_group=cg))
    ms0 = timeit(lambda: ops.linear_fwd(x, w, b, epi, out=out, out2=out2, aux=aux, block_n=bn, cta_group=cg, dtype_flags=256))
    ms1 = timeit(lambda: ops.linear_fwd(x, w, b, epi, out=out, out2=out2, aux=aux, block_n=bn, cta_group=cg, dtype_flags=512))
    print("%s G%d M%5d N%5d K%5d bn%3d cg%d %-5s: %7.1f us %7.1f TF/s | noepi %7.1f nostore %7.1f us" % (
        tag, G, M, N, K, bn, cg, NAMES[epi], ms * 1e3, 2.0 * G * M * N * K / ms / 1e9, ms0 * 1e3, ms1 * 1e3), flush=True)


def dgrad(G, M, N, K, bn=0, epi=EPI_BF16, cg=0, out2=False):
    """dy [G,M,N] x w [G,N,K] -> dx [G,M,K]"""
    dy = torch.randn(G, M, N, device=dev).bfloat16(); w = torch.randn(G, N, K, device=dev).bfloat16()
    out = torch.empty(G, M, K, device=dev, dtype=torch.bfloat16)
    aux = torch.randn(G, M, K, device=dev).bfloat16() if epi == EPI_DGELU else None
    o2 = torch.empty_like(out) if out2 else None
    ms = timeit(lambda: ops.linear_dgrad(dy, w, epi, aux=aux, out=out, block_n=bn, cta_group=cg, out2=o2))
    print("dgrd G%d M%5d N%5d K%5d bn%3d cg%d %-5s%s: %7.1f us %7.1f TF/s" % (
        G, M, N, K, bn, cg, NAMES[epi], "+g" if out2 else "  ", ms * 1e3, 2.0 * G * M * N * K / ms / 1e9), flush=True)


def wgrad(G, M, N, K, splits, cg=0):
    dy = torch.randn(G, M, N, device=dev).bfloat16(); x = torch.randn(G, M, K, device=dev).bfloat16()
    dw = torch.zeros(G, N, K, device=dev)
    ms = timeit(lambda: ops.linear_wgrad(dy, x, dw, splits=splits, cta_group=cg))
    print("wgrd G%d M%5d N%5d K%5d s%2d cg%d     : %7.1f us %7.1f TF/s" % (
        G, M, N, K, splits, cg, ms * 1e3, 2.0 * G * M * N * K / ms / 1e9), flush=True)


print(torch.cuda.get_device_name(0))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
M = B * 197
fwd(1, 8192, 8192, 8192, 256)
fwd(2, M, 1152, 384, 256, tag="qkv ")
fwd(2, M, 1536, 384, 256, EPI_GELU, tag="fc1 ")
for bn, cg in ((128, 1), (128, 2), (384, 2)):
    fwd(2, M, 384, 384, bn, EPI_RESID_F32, cg, tag="proj")
    fwd(2, M, 384, 1536, bn, EPI_RESID_F32, cg, tag="fc2 ")
dgrad(2, M, 384, 1536, 256, EPI_DGELU, out2=True)   # fc2 dgrad
for bn in (128, 384):
    dgrad(2, M, 1536, 384, bn)   # fc1 dgrad
    dgrad(2, M, 1152, 384, bn)   # qkv dgrad
    dgrad(2, M, 384, 384, bn)    # proj dgrad
for s in (4, 7):
    wgrad(2, M, 1152, 384, s)
for s in (4, 6):
    wgrad(2, M, 1536, 384, s)
for s in (3, 4, 6):
    wgrad(2, M, 384, 1536, s)
wgrad(2, M, 384, 384, 16); wgrad(2, M, 384, 384, 8)
# reference point: cuBLAS through torch on the same shapes
for (m, n, k) in ((8192, 8192, 8192), (2 * M, 1536, 384), (2 * M, 1152, 384), (2 * M, 384, 1536), (2 * M, 384, 384)):
    a = torch.randn(m, k, device=dev).bfloat16(); b = torch.randn(n, k, device=dev).bfloat16()
    ms = timeit(lambda: torch.matmul(a, b.t()))
    print("cuBLAS M%5d N%5d K%5d: %7.1f us  %7.1f TF/s" % (m, n, k, ms * 1e3, 2.0 * m * n * k / ms / 1e9))
