"""One GEMM shape, a few launches - the target of `ncu --set full` captures.

usage: gpu_gemm_one.py fc1|qkv|fc2|proj|fc2ln|dgelu|wgrad
"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi-feature-vit_b200"))
from mfvit import ops
from mfvit._lib import EPI_BF16, EPI_RESID_F32, EPI_GELU, EPI_DGELU
dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "fc1"
G, M = 2, 32 * 197
bf = lambda *s: torch.randn(*s, device=dev).bfloat16()
for _ in range(3):
    if which == "fc1":
        out = torch.empty(G, M, 1536, device=dev, dtype=torch.bfloat16)
        ops.linear_fwd(bf(G, M, 384), bf(G, 1536, 384), torch.randn(G, 1536, device=dev), EPI_GELU, out=out,
                       out2=torch.empty_like(out), block_n=256)
    elif which == "qkv":
        ops.linear_fwd(bf(G, M, 384), bf(G, 1152, 384), torch.randn(G, 1152, device=dev), EPI_BF16, block_n=256)
    elif which == "fc2":
        ops.linear_fwd(bf(G, M, 1536), bf(G, 384, 1536), torch.randn(G, 384, device=dev), EPI_RESID_F32,
                       aux=torch.randn(G, M, 384, device=dev), block_n=128)
    elif which == "proj":
        ops.linear_fwd(bf(G, M, 384), bf(G, 384, 384), torch.randn(G, 384, device=dev), EPI_RESID_F32,
                       aux=torch.randn(G, M, 384, device=dev), block_n=128)
    elif which == "dgelu":
        ops.linear_dgrad(bf(G, M, 384), bf(G, 384, 1536), EPI_DGELU, aux=bf(G, M, 1536),
                         out2=torch.empty(G, M, 1536, device=dev, dtype=torch.bfloat16), block_n=256)
    elif which == "fc2ln":  # fc2 + residual + LayerNorm of the next block (MFV_EPI_RESID_LN), fp16 operands + bf16 copy
        ops.linear_fwd_ln(torch.randn(G, M, 1536, device=dev).half(), (torch.randn(G, 384, 1536, device=dev) * 0.05).half(),
                          torch.randn(G, 384, device=dev), torch.randn(G, M, 384, device=dev),
                          torch.ones(G, 384, device=dev), torch.zeros(G, 384, device=dev), f16=True, bf16_copy=True)
    elif which == "wgrad":
        ops.linear_wgrad(bf(G, M, 1536), bf(G, M, 384), torch.zeros(G, 1536, 384, device=dev), splits=4)
torch.cuda.synchronize()
print("done", which)
