"""Shared builders / metrics for the end-to-end parity tests (drop-in modules on CUDA kernels vs the oracle)."""
import importlib
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "multi-feature-vit_b200"), os.path.join(ROOT, "multi-feature-vit_b200", "dropin")):
    if p not in sys.path:
        sys.path.insert(0, p)

FUS_MOD = ("model.crossvit_2vits_2additionaloutputs_changenormlayer_location_removeextralclayer_"
           "changemodelinputlocation_std002_sum")

# the reference's own normalisation of ToTensor output (aihc_utils/image_transform.py:12-16)
CXR_MEAN, CXR_STD = 0.5045, 0.2462
ENH_MEAN, ENH_STD = (0.2243, 0.5507, 0.6865), (0.1026, 0.2995, 0.3300)


def synthetic_pair(B, hw, rank=0, device="cpu"):
    g = torch.Generator().manual_seed(2024 + rank)
    u1 = torch.rand(B, 3, hw, hw, generator=g)
    u2 = torch.rand(B, 3, hw, hw, generator=g)
    cxr = (u1 - CXR_MEAN) / CXR_STD
    enh = (u2 - torch.tensor(ENH_MEAN).view(1, 3, 1, 1)) / torch.tensor(ENH_STD).view(1, 3, 1, 1)
    tgt = torch.randint(0, 3, (B,), generator=g)
    return cxr.to(device), enh.to(device), tgt.to(device)


def perturb_(module, std=0.02, seed=11):
    """LayerNorm affine and all biases moved off their trivial (1 / 0) init by N(0, 0.02) so that every term of the
    computation is exercised; weight matrices keep the reference's own random init."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in module.named_parameters():
            if p.requires_grad and p.dim() == 1:
                p.add_(torch.randn(p.shape, generator=g) * std)


def reference_head_init_(vit, num_classes=3):
    """MAIN_CA:309-316 / MAIN_LPFT:288-296: head = Linear(384, 3), weight ~ N(0, 0.01), bias = 0."""
    vit.head = torch.nn.Linear(vit.head.in_features, num_classes)
    vit.head.weight.data.normal_(mean=0.0, std=0.01)
    vit.head.bias.data.zero_()


def build_vit_pair(img_size=224, num_heads=6, num_classes=3, seed=0, device="cuda"):
    """(oracle_vit, dropin_vit) with identical weights."""
    import vits_returnftrs as vits
    from oracle import vit_ref
    torch.manual_seed(seed)
    ref = vit_ref.vit_small(img_size=img_size, num_heads=num_heads)
    reference_head_init_(ref, num_classes)
    perturb_(ref, seed=seed + 11)
    ours = vits.vit_small(img_size=img_size, num_heads=num_heads)
    ours.head = torch.nn.Linear(ours.head.in_features, num_classes)
    ours.load_state_dict(ref.state_dict(), strict=True)
    return ref.to(device), ours.to(device)


def build_mfvit_pair(img_size=224, seed=0, device="cuda"):
    """((ref_fus, ref_cxr, ref_enh), (fus, cxr, enh)) with identical weights."""
    from oracle import fusion_ref
    fm = importlib.import_module(FUS_MOD)
    r_c, o_c = build_vit_pair(img_size, seed=seed, device=device)
    r_e, o_e = build_vit_pair(img_size, seed=seed + 1, device=device)
    torch.manual_seed(seed + 2)
    r_f = fusion_ref.Fus_CrossViT(r_c, r_e)
    perturb_(r_f, seed=seed + 13)
    o_f = fm.Fus_CrossViT(o_c, o_e)
    o_f.load_state_dict(r_f.state_dict(), strict=True)
    return (r_f.to(device), r_c, r_e), (o_f.to(device), o_c, o_e)


def cos(a, b):
    return F.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0).item()


def grad_report(named_ref, named_ours, to_cpu=False, skip_zero=False):
    """per-tensor cosine similarity of gradients; returns (min_cos, worst_name, table).  skip_zero: leave out tensors
    whose reference gradient is zero up to rounding (< 1e-6 of the largest gradient entry of the model) - e.g. the final
    LayerNorm bias in front of a bias-free Linear + train-mode BatchNorm, whose true gradient is exactly 0."""
    ours = dict(named_ours)
    rows = []
    named_ref = list(named_ref)
    top = max([float(p.grad.abs().max()) for _, p in named_ref if p.grad is not None] + [0.0])
    for n, p in named_ref:
        if p.grad is None:
            continue
        if skip_zero and float(p.grad.abs().max()) < 1e-6 * top:
            continue
        g = ours[n].grad
        assert g is not None, "missing gradient for %s" % n
        if to_cpu:
            g = g.cpu()
        rows.append((cos(g, p.grad), n, float(p.grad.abs().max())))
    rows.sort()
    return rows[0][0], rows[0][1], rows


def mfvit_step(fus, cxr, enh, img_c, img_e, tgt, dedup=None):
    kw = {} if dedup is None else {"dedup": dedup}
    fused, x_c, x_e = fus(cxr, enh, img_c, img_e, **kw)
    out = fused + x_c + x_e  # MAIN_CA:868
    loss = F.cross_entropy(out, tgt)  # MAIN_CA:873
    loss.backward()
    return out.detach(), loss.detach(), (fused.detach(), x_c.detach(), x_e.detach())
