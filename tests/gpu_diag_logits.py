"""Diagnostic: distribution of the MF-ViT CA logits error vs the fp32 oracle over seeds (B = 32)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "multi-feature-vit_b200"), os.path.join(ROOT, "multi-feature-vit_b200", "dropin"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import e2e_common as E
from mfvit import _lib
torch.backends.cuda.matmul.allow_tf32 = False
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
for seed in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    (r_f, r_c, r_e), (o_f, o_c, o_e) = E.build_mfvit_pair(seed=seed)
    img_c, img_e, tgt = E.synthetic_pair(B, 224, rank=seed, device="cuda")
    with torch.no_grad():
        fr, xc, xe = r_f(r_c, r_e, img_c, img_e, dedup=True)
        tok_r = torch.stack([r_c.features3D(img_c), r_e.features3D(img_e)])
        row = []
        for fuse in (0, 1):
            lib.mfv_set_option(b"patch_tma", fuse)
            fo, oc, oe = o_f(o_c, o_e, img_c, img_e)
            from mfvit.engine import encode, engine_for
            tok_o = encode(engine_for(o_c, o_e), [img_c, img_e])
            row.append((float((fo - fr).abs().max()), float((oc - xc).abs().max()), float((oe - xe).abs().max()),
                        float((tok_o - tok_r).abs().max()), float((tok_o[:, :, 0] - tok_r[:, :, 0]).pow(2).mean().sqrt())))
    print("seed %d  " % seed + "   ".join("patch_tma=%d fused %.2e cxr %.2e enh %.2e tok max %.2e cls rms %.2e" % ((i,) + r) for i, r in enumerate(row)), flush=True)
