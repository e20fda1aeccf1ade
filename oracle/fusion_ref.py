"""Restatement of the reference fusion model, pure PyTorch, same module tree / state-dict keys as the reference:
PreNorm MOD:15-21, CrossAttention MOD:108-137, MultiScaleTransformerEncoder FUS:12-65, Fus_CrossViT FUS:72-157
(init FUS:117-124).  `forward` follows the as-written graph (torch.cat copies, LayerNorm over all 197 rows);
`closed_form` is the de-duplicated CLS-row formula of SURVEY.md section 3.2 that the CUDA kernel implements.
"""
import torch
import torch.nn as nn


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)  # default eps 1e-5 (MOD:18)
        self.fn = fn

    def forward(self, x):
        return self.fn(self.norm(x))


class CrossAttention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5  # MOD:114
        self.wq = nn.Linear(dim, dim, bias=qkv_bias)
        self.wk = nn.Linear(dim, dim, bias=qkv_bias)
        self.wv = nn.Linear(dim, dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):  # MOD:123-137
        B, N, C = x.shape
        h = self.num_heads
        q = self.wq(x[:, 0:1]).reshape(B, 1, h, C // h).permute(0, 2, 1, 3)
        k = self.wk(x).reshape(B, N, h, C // h).permute(0, 2, 1, 3)
        v = self.wv(x).reshape(B, N, h, C // h).permute(0, 2, 1, 3)
        attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, 1, C)
        return self.proj(x)


class MultiScaleTransformerEncoder(nn.Module):
    def __init__(self, small_dim=384, large_dim=384, cross_attn_depth=1, cross_attn_heads=3):
        super().__init__()
        self.cross_attn_layers = nn.ModuleList([])
        for _ in range(cross_attn_depth):
            self.cross_attn_layers.append(nn.ModuleList([
                PreNorm(large_dim, CrossAttention(large_dim, num_heads=cross_attn_heads)),
                nn.LayerNorm(large_dim, eps=1e-6),
                PreNorm(small_dim, CrossAttention(small_dim, num_heads=cross_attn_heads)),
                nn.LayerNorm(small_dim, eps=1e-6),
            ]))

    def forward(self, xs, xl):  # FUS:35-65
        for cross_attn_s, n_l, cross_attn_l, n_s in self.cross_attn_layers:
            small_class, x_small = xs[:, 0], xs[:, 1:]
            large_class, x_large = xl[:, 0], xl[:, 1:]
            cal_q = large_class.unsqueeze(1)
            cal_out = cal_q + cross_attn_l(torch.cat((cal_q, x_small), dim=1))
            xl = n_l(torch.cat((cal_out, x_large), dim=1))
            cal_q = small_class.unsqueeze(1)
            cal_out = cal_q + cross_attn_s(torch.cat((cal_q, x_large), dim=1))
            xs = n_s(torch.cat((cal_out, x_small), dim=1))
        return xs, xl


class Fus_CrossViT(nn.Module):
    def __init__(self, model_vit_cxr, model_vit_enh, num_classes=3, small_dim=384, large_dim=384, cross_attn_depth=1,
                 multi_scale_enc_depth=1, heads=3, dropout=0.0, pool="cls"):
        super().__init__()
        self.vit_features_cxr = model_vit_cxr.features3D  # bound method, not a sub-module (FUS:80)
        self.vit_features_enh = model_vit_enh.features3D
        self.multi_scale_transformers = nn.ModuleList([
            MultiScaleTransformerEncoder(small_dim, large_dim, cross_attn_depth, heads)
            for _ in range(multi_scale_enc_depth)])
        self.pool = pool
        self.num_classes = num_classes
        self.mlp_head_cxr = nn.Sequential(nn.Linear(small_dim, num_classes))
        self.mlp_head_enh = nn.Sequential(nn.Linear(large_dim, num_classes))
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):  # FUS:117-124
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def fuse(self, cxr_ftrs, enh_ftrs):
        """FUS:137-155 given the two token tensors."""
        bs = cxr_ftrs.shape[0]
        for mst in self.multi_scale_transformers:
            cxr_ca, enh_ca = mst(cxr_ftrs, enh_ftrs)
        cxr_fus = cxr_ftrs + cxr_ca
        enh_fus = enh_ftrs + enh_ca
        cxr_cls = cxr_fus.mean(dim=1) if self.pool == "mean" else cxr_fus[:, 0]
        enh_cls = enh_fus.mean(dim=1) if self.pool == "mean" else enh_fus[:, 0]
        cxr_ds = self.mlp_head_cxr(cxr_cls).view(bs, 1, self.num_classes)
        enh_ds = self.mlp_head_enh(enh_cls).view(bs, 1, self.num_classes)
        return torch.sum(torch.cat([cxr_ds, enh_ds], dim=1), dim=1).squeeze(dim=1)

    def forward(self, vit_cxr, vit_enh, img_cxr, img_enh, dedup=False):
        """As written (FUS:126-157): 4 backbone passes.  dedup=True runs each backbone once (identical function)."""
        cxr_ftrs = self.vit_features_cxr(img_cxr)
        enh_ftrs = self.vit_features_enh(img_enh)
        if dedup:
            x_cxr = vit_cxr.head(cxr_ftrs[:, 0])
            x_enh = vit_enh.head(enh_ftrs[:, 0])
        else:
            x_cxr = vit_cxr(img_cxr)
            x_enh = vit_enh(img_enh)
        return self.fuse(cxr_ftrs, enh_ftrs), x_cxr, x_enh

    def closed_form(self, f_c, f_e):
        """only the CLS rows are consumed.  Returns the fused logits from final-normed tokens."""
        layers = self.multi_scale_transformers[0].cross_attn_layers[0]
        ca0, ln1, ca2, ln3 = layers[0], layers[1], layers[2], layers[3]

        def ca(pre, q_cls, patches):
            return pre(torch.cat((q_cls.unsqueeze(1), patches), dim=1))[:, 0]

        e = f_e[:, 0] + ln1(f_e[:, 0] + ca(ca2, f_e[:, 0], f_c[:, 1:]))
        c = f_c[:, 0] + ln3(f_c[:, 0] + ca(ca0, f_c[:, 0], f_e[:, 1:]))
        return self.mlp_head_cxr(c) + self.mlp_head_enh(e)
