"""Input pipeline alone: augment kernel against the HBM roofline, host staging time per batch, loader-only pairs/s.

usage: python tests/gpu_loader_bench.py [batch] [img]
"""
import json, os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multi-feature-vit_b200"))
from mfvit import data, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
img = int(sys.argv[2]) if len(sys.argv) > 2 else 224
dev = torch.device("cuda", 0)
peak = 6550.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak))
except Exception:  # noqa: BLE001
    pass
print(torch.cuda.get_device_name(0), "batch", B, "img", img)

# ---- kernel alone: ring of inputs larger than L2, graph-timed
g = torch.Generator().manual_seed(0)
ring = 12
srcs = [torch.randint(0, 256, (B, img, img, 3), dtype=torch.uint8, generator=g).to(dev) for _ in range(ring)]
outs = [torch.empty(B, 3, img, img, device=dev) for _ in range(ring)]
mean, std = (torch.tensor(v, device=dev) for v in data.STATS["Train_Mix"])
for name, deg in (("flip only", 0.0), ("flip + rotation +-1 deg", 1.0), ("flip + rotation +-30 deg", 30.0)):
    params = data.pack_params(data.draw_train_params_batch(B, img, img, img, deg, g), img, img).to(dev)
    for s, o in zip(srcs, outs):
        ops.augment_u8(s, params, mean, std, img, out=o)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(3):
            for s, o in zip(srcs, outs):
                ops.augment_u8(s, params, mean, std, img, out=o)
    graph.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); graph.replay(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / (3 * ring))
    nbytes = B * img * img * 3 * (1 + 4)
    print("augment_u8 %-26s %6.1f MB  %7.1f us  %6.0f GB/s  %5.1f %% of %.0f GB/s" %
          (name, nbytes / 1e6, best * 1e3, nbytes / best / 1e6, 100 * nbytes / best / 1e6 / peak, peak), flush=True)

# ---- host staging per batch and loader-only rate
n_store = 16 * B
store = data.PairedU8Store(torch.randint(0, 256, (n_store, img, img, 3), dtype=torch.uint8, generator=g),
                           torch.randint(0, 256, (n_store, img, img, 3), dtype=torch.uint8, generator=g),
                           torch.randint(0, 3, (n_store,), generator=g))
loader = data.PairedDeviceLoader(store, B, crop=img, degrees=True, training=True, device=dev, drop_last=True)
gen = torch.Generator().manual_seed(1)
idx = torch.randperm(n_store)[:B]
t0 = time.perf_counter()
for i in range(20):
    loader._stage(loader.slots[i % 3], idx, loader._draw(B, gen))
torch.cuda.synchronize()
print("host staging (gather + draws + pack + H2D / transform enqueue): %.2f ms / batch" % ((time.perf_counter() - t0) / 20 * 1e3))
for s in loader.slots:
    s.used = False
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    for epoch in range(4):
        loader.set_epoch(epoch)
        for xc, xe, y in loader:
            n += xc.shape[0]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("loader alone: %d pairs in %.1f ms = %.0f pairs/s (%.2f ms / batch)" % (n, dt * 1e3, n / dt, dt / (n / B) * 1e3), flush=True)
