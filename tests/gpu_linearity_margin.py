"""Margin of tests/test_gpu_parity.py::test_gradients_are_linear_in_upstream: the worst per-tensor cosine between the CXR
branch's gradients at upstream scale 3 and 1, over a few repetitions (split-K reduce-add order is not deterministic)."""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import e2e_common as E
_, (o_f, o_c, o_e) = E.build_mfvit_pair(seed=7)
img_c, img_e, tgt = E.synthetic_pair(32, 224, device="cuda")
def grads(scale):
    for m in (o_f, o_c, o_e):
        m.zero_grad(set_to_none=True)
    fused, x_c, x_e = o_f(o_c, o_e, img_c, img_e)
    (F.cross_entropy(fused + x_c + x_e, tgt) * scale).backward()
    return {n: p.grad.clone() for n, p in o_c.named_parameters() if p.grad is not None}
for rep in range(4):
    g1, g3 = grads(1.0), grads(3.0)
    c = sorted((E.cos(g3[n], g1[n]), n) for n in g1)
    print("rep %d: worst %.6f %s | second %.6f %s | third %.6f %s" % (rep, c[0][0], c[0][1], c[1][0], c[1][1], c[2][0], c[2][1]), flush=True)
