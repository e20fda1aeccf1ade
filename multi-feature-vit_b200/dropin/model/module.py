"""Drop-in for moco_pretraining/moco/model/module.py: the names FUS:6 imports.  PreNorm / CrossAttention hold the
parameters (same attribute names -> same state-dict keys); their arithmetic is executed by the fused CLS cross-attention
kernel inside Fus_CrossViT.forward (mfv_fusion_fwd/bwd), so they have no standalone eager forward.  Residual,
FeedForward and Attention are dead code in the reference's live paths and are kept as names only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _path  # noqa: E402,F401
import torch.nn as nn  # noqa: E402
from mfvit import MfvError  # noqa: E402


def _no_eager(self, *a, **k):
    raise MfvError("%s has no eager forward; it is evaluated inside Fus_CrossViT.forward by the fused sm_100a "
                   "cross-attention kernel" % type(self).__name__)


class PreNorm(nn.Module):
    def __init__(self, dim, fn):  # MOD:15-21
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn

    forward = _no_eager


class CrossAttention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):  # MOD:108-121
        super().__init__()
        if qkv_bias or qk_scale is not None or attn_drop or proj_drop:
            raise MfvError("fused CrossAttention supports the reference's live configuration only "
                           "(qkv_bias=False, default scale, no dropout)")
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.wq = nn.Linear(dim, dim, bias=False)
        self.wk = nn.Linear(dim, dim, bias=False)
        self.wv = nn.Linear(dim, dim, bias=False)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    forward = _no_eager


class _Unused(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("%s is dead code in the reference's live paths and is not part of the MF-ViT CA "
                                  "hot path" % type(self).__name__)


class Residual(_Unused):
    pass


class FeedForward(_Unused):
    pass


class Attention(_Unused):
    pass
