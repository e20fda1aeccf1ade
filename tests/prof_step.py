"""Minimal driver for ncu: 2 warm-up steps + N profiled steps of the MF-ViT CA trainer at config-2 size."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import e2e_common as E  # noqa: E402
import vits_returnftrs as vits  # noqa: E402
from mfvit.trainer import MFViTCATrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
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
    loss = tr.step(c, e, t)
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off: only the steps below are captured
for _ in range(steps):
    loss = tr.step(c, e, t)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(loss))
