"""Where does the loader-fed step lose time?  A: resident inputs; B: + augment kernels from a resident uint8 batch;
C: + H2D of the uint8 batch on a copy stream (main thread); D: full PairedDeviceLoader."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "multi-feature-vit_b200"), os.path.join(ROOT, "multi-feature-vit_b200", "dropin"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import e2e_common as E
from mfvit import data, ops
from mfvit.trainer import MFViTCATrainer
B, img, steps = 32, 224, 40
dev = torch.device("cuda", 0)
_, (fus, cxr, enh) = E.build_mfvit_pair(seed=0)
tr = MFViTCATrainer(fus, cxr, enh, lr=1e-3, momentum=0.9, train_backbones=True)
batch = E.synthetic_pair(B, img, device="cuda")
for _ in range(3):
    tr.step(*batch)
tr.capture_graph(*batch)
bufs = tuple(tr._g_inputs)
g = torch.Generator().manual_seed(0)
u8 = [torch.randint(0, 256, (B, img, img, 3), dtype=torch.uint8, generator=g) for _ in range(2)]
u8p = [t.pin_memory() for t in u8]
u8d = [t.to(dev) for t in u8]
params = data.pack_params(data.draw_train_params_batch(B, img, img, img, 1.0, g), img, img).to(dev)
stats = [tuple(torch.tensor(v, device=dev) for v in data.STATS[k]) for k in ("data", "Train_Mix")]

def timed(name, body):
    for i in range(5):
        body(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for i in range(steps):
        body(i)
    b.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("%-60s %.3f ms/step (host enqueue %.3f ms/step)" % (name, a.elapsed_time(b) / steps, (t1 - t0) / steps * 1e3), flush=True)

timed("A resident inputs", lambda i: tr.step(*bufs))
def body_b(i):
    for t in range(2):
        ops.augment_u8(u8d[t], params, stats[t][0], stats[t][1], img, out=bufs[t])
    tr.step(*bufs)
timed("B + 2 augment kernels (resident uint8)", body_b)
cs = torch.cuda.Stream()
ev = torch.cuda.Event()
def body_c(i):
    with torch.cuda.stream(cs):
        for t in range(2):
            u8d[t].copy_(u8p[t], non_blocking=True)
        ev.record(cs)
    torch.cuda.current_stream().wait_event(ev)
    body_b(i)
timed("C + H2D of uint8 on a copy stream (serialised with the step)", body_c)
store = data.PairedU8Store(torch.randint(0, 256, (8 * B, img, img, 3), dtype=torch.uint8, generator=g),
                           torch.randint(0, 256, (8 * B, img, img, 3), dtype=torch.uint8, generator=g),
                           torch.randint(0, 3, (8 * B,), generator=g))
for nstore, label in ((8, "8 batches / epoch"), (64, "64 batches / epoch")):
    if nstore != 8:
        store = data.PairedU8Store(store.cxr.repeat(8, 1, 1, 1), store.enh.repeat(8, 1, 1, 1), store.labels.repeat(8))
    loader = data.PairedDeviceLoader(store, B, crop=img, degrees=True, training=True, device=dev, drop_last=True)
    def run(n):
        done, epoch = 0, 0
        while done < n:
            loader.set_epoch(epoch)
            for xc, xe, y in loader:
                tr.step(xc, xe, y)
                done += 1
                if done == n:
                    return
            epoch += 1
    run(5)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(steps); b.record(); torch.cuda.synchronize()
    print("%-60s %.3f ms/step" % ("D PairedDeviceLoader, " + label, a.elapsed_time(b) / steps), flush=True)
timed("A again", lambda i: tr.step(*bufs))

# ---- per-step loss reads: where does the turnaround go?
def sync_loop(name, batches_iter):
    turn, stept = [], []
    t_sync = None
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    n = 0
    for xc, xe, y in batches_iter:
        t0 = time.perf_counter()
        if t_sync is not None:
            turn.append(t0 - t_sync)
        loss = tr.step(xc, xe, y)
        t1 = time.perf_counter()
        float(loss)
        t_sync = time.perf_counter()
        stept.append(t1 - t0)
        n += 1
    b.record(); torch.cuda.synchronize()
    import statistics
    print("%-44s %.3f ms/step; host: hand-back->next batch %.3f ms (median), step() call %.3f ms" %
          (name, a.elapsed_time(b) / n, statistics.median(turn) * 1e3, statistics.median(stept) * 1e3), flush=True)

sync_loop("S1 resident inputs, loss read every step", ((bufs[0], bufs[1], bufs[2]) for _ in range(steps)))
other = tuple(t.clone() for t in bufs)
sync_loop("S2 other device tensors (3 D2D copies)", (other for _ in range(steps)))
loader = data.PairedDeviceLoader(store, B, crop=img, degrees=True, training=True, device=dev, drop_last=True)
def gen(n):
    done = 0
    for xc, xe, y in loader:
        yield xc, xe, y
        done += 1
        if done == n:
            return
sync_loop("S3 PairedDeviceLoader (64 batches / epoch)", gen(steps))
