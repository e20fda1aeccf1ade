"""MoCo-v3-structure / v2-loss pretraining step (BASELINE configs[3] shape on one GPU): 128 images per view, K = 65 536.

The loop body is the reference's (MAIN_PRE:520-548): cosine moco momentum, autocast, criterion(output, target),
GradScaler, AdamW - run unchanged over the drop-in MoCo_ViT, whose encoders / EMA / InfoNCE / enqueue are libmfvit
kernels.  usage: python tests/gpu_moco_bench.py [batch=128] [steps=10]
"""
import math, os, sys, time
from functools import partial
from types import SimpleNamespace
import torch
import torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "multi-feature-vit_b200"), os.path.join(ROOT, "multi-feature-vit_b200", "dropin"),
          os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import importlib
import e2e_common as E
import vits
bm = importlib.import_module("moco.builder_vit_mocov3structure_mocov2loss")

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
args = SimpleNamespace(arch="vit_small", epochs=100, moco_m=0.99)
torch.manual_seed(0)
model = bm.MoCo_ViT(partial(vits.vit_small, stop_grad_conv1=True), args, 256, 4096, 0.2).cuda()
criterion = nn.CrossEntropyLoss().cuda()
optimizer = torch.optim.AdamW(model.parameters(), 1.5e-4, weight_decay=0.1)
scaler = torch.cuda.amp.GradScaler()
im_q, im_k, _ = E.synthetic_pair(B, 224, device="cuda")
iters_per_epoch = 100


def step(i):
    moco_m = 1. - 0.5 * (1. + math.cos(math.pi * (i / iters_per_epoch) / args.epochs)) * (1. - args.moco_m)
    with torch.cuda.amp.autocast(True):
        output, target = model(im_q, im_k, moco_m)
        loss = criterion(output, target)
    optimizer.zero_grad()
    scaler.scale(loss).backward()
    scaler.step(optimizer)
    scaler.update()
    return loss


for i in range(3):
    loss = step(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(steps):
    loss = step(3 + i)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
flops = 4.74e12 * B / 128  # q path fwd+bwd + k path fwd + l_neg
print("%s  MoCo step B=%d: %.2f ms/step  %.0f img/s  %.0f TFLOP/s algorithmic  loss %.4f" % (
    torch.cuda.get_device_name(0), B, ms, B / ms * 1e3, flops / ms / 1e9, float(loss)))
