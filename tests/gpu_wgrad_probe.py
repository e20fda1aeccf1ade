"""Probe for ncu: the fc1 weight-gradient GEMM (1536 x 384 output, K = 6304 tokens, G = 2) with different split counts.
usage: gpu_wgrad_probe.py <splits>"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multi-feature-vit_b200"))
from mfvit import ops
dev = "cuda"
splits = int(sys.argv[1]) if len(sys.argv) > 1 else 6
G, M = 2, 32 * 197
dy = torch.randn(G, M, 1536, device=dev).bfloat16()
x = torch.randn(G, M, 384, device=dev).bfloat16()
dw = torch.zeros(G, 1536, 384, device=dev)
for _ in range(3):
    ops.linear_wgrad(dy, x, dw, splits=splits)
torch.cuda.synchronize()
print("done", splits)
