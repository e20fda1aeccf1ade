"""Restatement of the absent `vits.py` / `vits_returnftrs.py` (facebookresearch/moco-v3 VisionTransformerMoCo on top of
timm-0.4.x VisionTransformer), pure PyTorch.  Call sites it must satisfy: MAIN_CA:44,289-290; MAIN_LPFT:44,276;
MAIN_PRE:39,274; BLD:28-30,217-222; FUS:80,83,128-135.  In-tree restatements of the same math followed here: attention
MOD:52-64, GELU-MLP MOD:26-32, pre-norm residual block fuseattention.py:75-81, token assembly crossvit.py:130-146,
patch embedding FUS:197-221.  State-dict keys: SURVEY.md section 3.5.
"""
import math
from functools import partial, reduce
from operator import mul

import torch
import torch.nn as nn


class PatchEmbed(nn.Module):
    """Conv2d(k=16, s=16) -> flatten(2).transpose(1, 2)   (commented restatement FUS:197-221)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=384):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class Attention(nn.Module):
    """timm Attention: qkv reshape (B,N,3,H,C/H).permute(2,0,3,1,4), scale = head_dim**-0.5 (same math MOD:52-64)."""

    def __init__(self, dim, num_heads, qkv_bias=True):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)) * self.scale
        attn = attn.softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj(x)


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()  # exact erf GELU
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


def sincos_pos_embed(grid_h, grid_w, embed_dim, temperature=10000.0):
    """MoCo-v3 build_2d_sincos_position_embedding (fixed table, requires_grad=False, CLS slot = zeros)."""
    gw = torch.arange(grid_w, dtype=torch.float32)
    gh = torch.arange(grid_h, dtype=torch.float32)
    gw, gh = torch.meshgrid(gw, gh, indexing="ij")
    pos_dim = embed_dim // 4
    omega = torch.arange(pos_dim, dtype=torch.float32) / pos_dim
    omega = 1.0 / (temperature ** omega)
    out_w = torch.einsum("m,d->md", [gw.flatten(), omega])
    out_h = torch.einsum("m,d->md", [gh.flatten(), omega])
    pos = torch.cat([torch.sin(out_w), torch.cos(out_w), torch.sin(out_h), torch.cos(out_h)], dim=1)[None]
    return torch.cat([torch.zeros(1, 1, embed_dim), pos], dim=1)


class VisionTransformerMoCo(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=384, depth=12, num_heads=6,
                 mlp_ratio=4.0, stop_grad_conv1=False, norm_eps=1e-6):
        super().__init__()
        self.num_classes = num_classes
        self.embed_dim = self.num_features = embed_dim
        self.num_tokens = 1
        norm_layer = partial(nn.LayerNorm, eps=norm_eps)
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads, mlp_ratio, norm_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        # MoCo-v3 initialisation
        self.pos_embed = nn.Parameter(sincos_pos_embed(*self.patch_embed.grid_size, embed_dim))
        self.pos_embed.requires_grad = False
        for name, m in self.named_modules():
            if isinstance(m, nn.Linear):
                if "qkv" in name:
                    val = math.sqrt(6.0 / float(m.weight.shape[0] // 3 + m.weight.shape[1]))
                    nn.init.uniform_(m.weight, -val, val)
                else:
                    nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)
        nn.init.normal_(self.cls_token, std=1e-6)
        val = math.sqrt(6.0 / float(3 * reduce(mul, self.patch_embed.patch_size, 1) + embed_dim))
        nn.init.uniform_(self.patch_embed.proj.weight, -val, val)
        nn.init.zeros_(self.patch_embed.proj.bias)
        if stop_grad_conv1:
            self.patch_embed.proj.weight.requires_grad = False
            self.patch_embed.proj.bias.requires_grad = False

    def features3D(self, x):
        """All tokens after the final LayerNorm, [B, N+1, C]  (FUS:128 'b, 197, 384'; crossvit.py:130-146)."""
        x = self.patch_embed(x)
        cls = self.cls_token.expand(x.shape[0], -1, -1)
        x = torch.cat((cls, x), dim=1)
        x = x + self.pos_embed
        x = self.blocks(x)
        return self.norm(x)

    def forward_features(self, x):
        return self.features3D(x)[:, 0]

    def forward(self, x):
        return self.head(self.forward_features(x))


def vit_small(**kwargs):
    kwargs.setdefault("num_heads", 6)  # north_star: head_dim 64; upstream MoCo-v3 uses 12
    return VisionTransformerMoCo(patch_size=16, embed_dim=384, depth=12, mlp_ratio=4, **kwargs)


def vit_base(**kwargs):
    kwargs.setdefault("num_heads", 12)
    return VisionTransformerMoCo(patch_size=16, embed_dim=768, depth=12, mlp_ratio=4, **kwargs)
