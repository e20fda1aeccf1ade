"""Where the attention backward (S <= 224 kernel) spends its time: clock64 sums per compute warp and phase, averaged per CTA
(MFVIT_ATTN_PROF carries the counter buffer's address).  python tests/gpu_attn_prof.py [NB=64]"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi-feature-vit_b200"))
dev = "cuda"
prof = torch.zeros(16 * 8 + 1, device=dev, dtype=torch.int64)
os.environ["MFVIT_ATTN_PROF"] = str(prof.data_ptr())
from mfvit import ops
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S, H, D = 197, 6, 64
PH = ["prologue", "wait S/dP", "tmem+math", "wait prev MMAs", "P/dS to smem", "dK/dV out", "dQ out + drain", "-"]
for f16 in (False, True):
    qkv = torch.randn(NB, S, 3, H, D, device=dev)
    qkv = qkv.half() if f16 else qkv.bfloat16()
    r = ops.attn_fwd(qkv, H, f16=f16, bf16_copy=f16)
    o, lse = (r[2], r[1]) if f16 else r
    do = torch.randn(NB, S, H, D, device=dev).bfloat16()
    for _ in range(3): ops.attn_bwd(qkv, o, do, lse)
    torch.cuda.synchronize(); prof.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.attn_bwd(qkv, o, do, lse); b.record(); torch.cuda.synchronize()
    n = int(prof[128])
    p = prof[:128].view(16, 8).double() / max(n, 1)
    print("%s: %.1f us, %d CTAs; per CTA, mean over the 16 compute warps (clk): " % ("f16" if f16 else "bf16", a.elapsed_time(b) * 1e3, n)
          + "  ".join("%s %.0f" % (PH[k], float(p[:, k].mean())) for k in range(7)) + "  | total %.0f" % float(p.sum(1).mean()), flush=True)
