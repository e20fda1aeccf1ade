"""Phase times of the MF-ViT CA step from cumulative CUDA graphs (forward | + fusion forward / loss / fusion backward |
+ encoder backward | whole step with the optimizer), 20 replays each.  python tests/gpu_phase_probe.py [pairs=32]"""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import e2e_common as E  # noqa: E402
import vits_returnftrs as vits  # noqa: E402
from mfvit import ops  # noqa: E402
from mfvit.trainer import MFViTCATrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
fm = importlib.import_module(E.FUS_MOD)
cxr, enh = vits.vit_small(), vits.vit_small()
for v in (cxr, enh):
    E.reference_head_init_(v)
fus = fm.Fus_CrossViT(cxr, enh)
dev = torch.device("cuda:0")
cxr.to(dev), enh.to(dev), fus.to(dev)
tr = MFViTCATrainer(fus, cxr, enh, train_backbones=True)
c, e, t = E.synthetic_pair(B, 224, device=dev)
for _ in range(2):
    tr.step(c, e, t)
torch.cuda.synchronize()
eng, lay = tr.engine, tr.engine.layout


def phase(upto):
    tok, lease = eng.forward([c, e], save=True)
    if upto == 1:
        lease.release()
        return
    dtok, d_x, scratch, d_fused, dl3 = tr._bufs[(B, dev)]
    ops.fusion_bwd_join(dev)
    fused, x = ops.fusion_fwd(tok, tr._pstruct, B, lay.S, lay.C, tr.heads, tr.NC, saved=scratch)
    loss, dlogits = ops.ce_small(fused, x[0], x[1], t)
    ops.fill_(tr._small.grad, 0.0)
    dl3.copy_(dlogits.unsqueeze(0).expand_as(dl3))
    ops.fusion_bwd(tok, tr._pstruct, tr._gstruct, d_fused, d_x, B, lay.S, lay.C, tr.heads, tr.NC, dtok=dtok, scratch=scratch, defer=True)
    if upto == 2:
        ops.fusion_bwd_join(dev)
        lease.release()
        return
    eng.backward(lease, dtok, zero=True)
    ops.fusion_bwd_join(dev)


def time_graph(fn, reps=20):
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2): fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


t1 = time_graph(lambda: phase(1)); t2 = time_graph(lambda: phase(2)); t3 = time_graph(lambda: phase(3))
tr.capture_graph(c, e, t)
for _ in range(3): tr.step(c, e, t)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): tr.step(c, e, t)
b.record(); torch.cuda.synchronize()
t4 = a.elapsed_time(b) / 20
print("pairs %d: forward %.3f ms | + fusion fwd/loss/bwd %.3f (%.3f) | + encoder backward %.3f (%.3f) | whole step %.3f (optimizer etc. %.3f)"
      % (B, t1, t2, t2 - t1, t3, t3 - t2, t4, t4 - t3), flush=True)
