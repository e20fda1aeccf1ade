"""HBM-bound kernels of the path against the measured copy bandwidth (MEASURED_PEAKS.json, else the profiling guide's
fallback): LayerNorm fwd/bwd, EMA, InfoNCE fwd/bwd, SGD step, shadow cast.  Algorithmic bytes per launch (DESIGN.md
section 3) / CUDA-event time.  Buffers are rotated through a ring larger than the 126 MB L2 so every launch streams from
DRAM.   usage: python tests/gpu_hbm_bench.py
"""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-feature-vit_b200"))
from mfvit import ops
dev = "cuda"
peak = 6550.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak))
except Exception:  # noqa: BLE001
    pass


def timeit(fns, reps=3):
    """fns: list of closures over DIFFERENT buffers (ring); returns device ms per launch.  The ring x reps sequence is
    captured in one CUDA graph so that Python / allocator time per call (~20 us) does not pollute 10 us kernels."""
    for f in fns: f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for f in fns: f()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / (reps * len(fns)))
    return best


def report(name, nbytes, ms):
    gbs = nbytes / ms / 1e6
    print("%-44s %8.1f MB  %8.1f us  %7.0f GB/s  %5.1f %% of %.0f GB/s" % (name, nbytes / 1e6, ms * 1e3, gbs, 100 * gbs / peak, peak), flush=True)


print(torch.cuda.get_device_name(0))
G, rows, C = 2, 32 * 197, 384
ring = 8  # 8 x (19.4 + ...) MB > L2
# ---- LayerNorm forward: reads x f32, writes fp16 + bf16 copy, mean / rstd
xs = [torch.randn(G, rows, C, device=dev) for _ in range(ring)]
gam, bet = torch.randn(G, C, device=dev), torch.randn(G, C, device=dev)
ms = timeit([lambda x=x: ops.layernorm_fwd(x, gam, bet, 1e-6, f16=True, bf16_copy=True) for x in xs])
report("layernorm fwd (f32 -> fp16 + bf16 copy)", G * rows * (C * (4 + 2 + 2) + 8), ms)
ms = timeit([lambda x=x: ops.layernorm_fwd(x, gam, bet, 1e-6) for x in xs])
report("layernorm fwd (f32 -> bf16)", G * rows * (C * (4 + 2) + 8), ms)
# ---- LayerNorm backward: reads x f32, dy bf16, dres f32; writes dx f32 + bf16
y16, _, mean, rstd = ops.layernorm_fwd(xs[0], gam, bet, 1e-6, want_bf16=True, want_f32=True)
dys = [torch.randn(G, rows, C, device=dev).bfloat16() for _ in range(ring)]
dres = [torch.randn(G, rows, C, device=dev) for _ in range(ring)]
dg, db, dc = torch.zeros(G, C, device=dev), torch.zeros(G, C, device=dev), torch.zeros(G, C, device=dev)
ms = timeit([lambda i=i: ops.layernorm_bwd(dys[i], xs[i], mean, rstd, gam, dres=dres[i], dgamma=dg, dbeta=db, dx_colsum=dc)
             for i in range(ring)])
report("layernorm bwd (+dres, dgamma, dbeta, colsum)", G * rows * (C * (4 + 2 + 4 + 4 + 2) + 8), ms)
del xs, dys, dres
# ---- EMA over the whole MoCo base encoder (41 080 704 fp32 parameters): k = k*m + q*(1-m)
n = 41_080_704
k, q = torch.randn(n, device=dev), torch.randn(n, device=dev)
chunks, nc, mx = ops.make_ema_chunks([(k, q)], dev)
ms = timeit([lambda: ops.ema_update_(chunks, nc, mx, 0.99)], reps=10)
report("EMA momentum update (41.08 M params)", 12 * n, ms)
# ---- SGD-momentum step with both 16-bit shadows (MF-ViT CA: 2 x 21.67 M parameters)
n2 = 2 * 21_665_664
p, g, buf = torch.randn(n2, device=dev), torch.randn(n2, device=dev), torch.zeros(n2, device=dev)
sh, sh16 = torch.empty(n2, device=dev, dtype=torch.bfloat16), torch.empty(n2, device=dev, dtype=torch.float16)
ms = timeit([lambda: ops.sgd_step_(p, g, buf, sh, 1e-3, 0.9, 0.0, False, shadow16=sh16)], reps=10)
report("SGD-momentum step + bf16/fp16 shadows", n2 * (12 + 8 + 4), ms)
ms = timeit([lambda: ops.cast_shadow(p, sh, sh16)], reps=10)
report("shadow cast f32 -> bf16 + fp16", n2 * 8, ms)
del p, g, buf, sh, sh16, k, q
# ---- InfoNCE (MoCo v2 loss): q @ queue over 65 536 keys, dim 256, 128 queries; queue streamed once per pass
N, D, K = 128, 256, 65536
qr, kr = torch.randn(N, D, device=dev), torch.randn(N, D, device=dev)
queues = [torch.nn.functional.normalize(torch.randn(D, K, device=dev), dim=0) for _ in range(3)]
outs = [ops.infonce_fwd(qr, kr, qu, 0.2) for qu in queues]
ms = timeit([lambda qu=qu: ops.infonce_fwd(qr, kr, qu, 0.2) for qu in queues], reps=5)
report("InfoNCE fwd (queue 64 MiB read, logits 32 MiB written)", D * K * 4 + N * (K + 1) * 4, ms)
q16s = [ops.queue16_update_(qu, torch.empty(D, K, device=dev, dtype=torch.float16)) for qu in queues]
ms = timeit([lambda q16=q16: ops.infonce_tc_fwd(qr, kr, q16, 0.2) for q16 in q16s], reps=5)
report("InfoNCE tc fwd (fp16 queue 32 MiB rd, logits 32 MiB wr + rd)", D * K * 2 + 2 * N * K * 4, ms)
tq = ops.infonce_tc_fwd(qr, kr, q16s[0], 0.2)
ms = timeit([lambda q16=q16: ops.infonce_tc_bwd(qr, tq[0], tq[1], q16, tq[2], tq[3], 0.2) for q16 in q16s], reps=5)
report("InfoNCE tc bwd (logits rd, fp16 grad 16 MiB wr + rd, queue rd)", D * K * 2 + N * K * 4 + 2 * N * K * 2, ms)
try:
    qn, kn, logits, lse, loss = outs[0]
    ms = timeit([lambda qu=qu: ops.infonce_bwd(qr, qn, kn, qu, logits, lse, 0.2) for qu in queues], reps=5)
    report("InfoNCE bwd (queue + logits read)", D * K * 4 + N * (K + 1) * 4, ms)
except Exception as e:  # noqa: BLE001
    print("infonce bwd skipped:", e)
