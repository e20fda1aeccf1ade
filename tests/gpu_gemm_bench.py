"""GEMM micro-benchmark: TFLOP/s of mfv_gemm per shape / tile width (CUDA events, 20 reps after 3 warm-ups)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi-feature-vit_b200"))
from mfvit import ops
from mfvit._lib import EPI_BF16, EPI_F32, EPI_ATOMIC_F32, EPI_RESID_F32, EPI_GELU
dev = "cuda"

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def fwd(G, M, N, K, bn, epi=EPI_BF16):
    x = torch.randn(G, M, K, device=dev).bfloat16(); w = torch.randn(G, N, K, device=dev).bfloat16()
    out = torch.empty(G, M, N, device=dev, dtype=torch.float32 if epi in (EPI_F32, EPI_RESID_F32) else torch.bfloat16)
    out2 = torch.empty_like(out) if epi == EPI_GELU else None
    aux = torch.randn(G, M, N, device=dev) if epi == EPI_RESID_F32 else None
    ms = timeit(lambda: ops.linear_fwd(x, w, None, epi, out=out, out2=out2, aux=aux, block_n=bn))
    print("fwd  G%d M%5d N%5d K%5d bn%3d epi%d: %7.1f us  %7.1f TF/s" % (G, M, N, K, bn, epi, ms * 1e3, 2.0 * G * M * N * K / ms / 1e9))

def wgrad(G, M, N, K, splits):
    dy = torch.randn(G, M, N, device=dev).bfloat16(); x = torch.randn(G, M, K, device=dev).bfloat16()
    dw = torch.zeros(G, N, K, device=dev)
    ms = timeit(lambda: ops.linear_wgrad(dy, x, dw, splits=splits))
    print("wgrd G%d M%5d N%5d K%5d s%2d      : %7.1f us  %7.1f TF/s" % (G, M, N, K, splits, ms * 1e3, 2.0 * G * M * N * K / ms / 1e9))

print(torch.cuda.get_device_name(0))
for bn in (128, 256):
    fwd(1, 8192, 8192, 8192, bn)
    fwd(1, 8192, 8192, 1024, bn)
    fwd(1, 8192, 8192, 384, bn)
    fwd(2, 6304, 1536, 384, bn)
    fwd(2, 6304, 1536, 384, bn, EPI_GELU)
    fwd(2, 6304, 1152, 384, bn)
fwd(2, 6304, 384, 384, 128); fwd(2, 6304, 384, 384, 128, EPI_RESID_F32); fwd(2, 6304, 384, 1536, 128, EPI_RESID_F32)
fwd(2, 12608, 1536, 384, 256); fwd(2, 12608, 384, 1536, 128, EPI_RESID_F32)
for s in (2, 4, 6, 8, 12):
    wgrad(2, 6304, 1152, 384, s)
wgrad(2, 6304, 1536, 384, 4); wgrad(2, 6304, 384, 1536, 4); wgrad(2, 6304, 384, 384, 16)
# reference point: cuBLAS through torch on the same shapes
for (M, N, K) in ((8192, 8192, 8192), (12608, 1536, 384), (12608, 1152, 384), (12608, 384, 1536)):
    a = torch.randn(M, K, device=dev).bfloat16(); b = torch.randn(N, K, device=dev).bfloat16()
    ms = timeit(lambda: torch.matmul(a, b.t()))
    print("cuBLAS M%5d N%5d K%5d: %7.1f us  %7.1f TF/s" % (M, N, K, ms * 1e3, 2.0 * M * N * K / ms / 1e9))
