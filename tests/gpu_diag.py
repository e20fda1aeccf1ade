"""Diagnostics: error of the CUDA path vs the fp32 oracle next to the error of a bf16-autocast oracle (same precision
class), to separate precision effects from bugs.  python tests/gpu_diag.py > gpurun_out/diag.log"""
import os
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import e2e_common as E

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def stats(tag, a, b):
    err = (a - b).abs()
    print("%-40s maxabs=%.3e meanabs=%.3e refmax=%.3e rel=%.3e cos=%.6f" % (
        tag, err.max().item(), err.mean().item(), b.abs().max().item(), err.max().item() / b.abs().max().item(), E.cos(a, b)))


ref, ours = E.build_vit_pair()
img, _, tgt = E.synthetic_pair(16, 224, device="cuda")
with torch.no_grad():
    tok_r = ref.features3D(img)
    tok_o = ours.features3D(img)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        tok_b = ref.features3D(img).float()
    stats("tokens ours vs fp32", tok_o, tok_r)
    stats("tokens autocast-bf16 vs fp32", tok_b, tok_r)
    stats("cls ours vs fp32", tok_o[:, 0], tok_r[:, 0])
    stats("cls autocast vs fp32", tok_b[:, 0], tok_r[:, 0])
    lo, lr = ours(img), ref(img)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lb = ref(img).float()
    stats("logits ours vs fp32", lo, lr)
    stats("logits autocast vs fp32", lb, lr)
    # per-block drift: run the oracle blocks on our embedding and compare residual stream norms
    x = ref.patch_embed(img)
    x = torch.cat((ref.cls_token.expand(x.shape[0], -1, -1), x), 1) + ref.pos_embed
    print("embed absmax %.3f" % x.abs().max().item())
    for i, blk in enumerate(ref.blocks):
        x = blk(x)
        if i in (0, 5, 11):
            print("block %d residual stream absmax %.3f rms %.3f" % (i, x.abs().max().item(), x.pow(2).mean().sqrt().item()))
out_r = ref(img); F.cross_entropy(out_r, tgt).backward()
out_o = ours(img); F.cross_entropy(out_o, tgt).backward()
mn, worst, rows = E.grad_report(ref.named_parameters(), ours.named_parameters())
print("grad cos min %.6f at %s" % (mn, worst))
for r in rows[:8]:
    print("   cos=%.6f %-40s gmax=%.3e" % r)
# autocast grads for comparison
ref2, _ = E.build_vit_pair()
with torch.autocast("cuda", dtype=torch.bfloat16):
    o2 = ref2(img)
F.cross_entropy(o2.float(), tgt).backward()
mn, worst, rows = E.grad_report(ref.named_parameters(), ref2.named_parameters())
print("autocast grad cos min %.6f at %s" % (mn, worst))

# ---- reference-init protocol (MAIN_CA:309-316 head ~ N(0,0.01)); logits abs error, both precision modes
import mfvit.engine as ME
for mode in ("fp16", "bf16"):
    ME.FWD_PRECISION = mode
    ref, ours = E.build_vit_pair(seed=21)
    with torch.no_grad():
        ref.head.weight.normal_(0, 0.01); ref.head.bias.zero_()
        ours.head.load_state_dict(ref.head.state_dict())
    img, _, tgt = E.synthetic_pair(32, 224, device="cuda")
    with torch.no_grad():
        stats("[%s] tokens vs fp32" % mode, ours.features3D(img), ref.features3D(img))
        stats("[%s] logits (ref init) vs fp32" % mode, ours(img), ref(img))
    out_r = ref(img); F.cross_entropy(out_r, tgt).backward()
    out_o = ours(img); F.cross_entropy(out_o, tgt).backward()
    mn, worst, rows = E.grad_report(ref.named_parameters(), ours.named_parameters())
    print("[%s] grad cos min %.6f at %s" % (mode, mn, worst))
