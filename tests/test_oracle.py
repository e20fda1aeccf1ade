"""CPU: pins the oracle (oracle/*.py) against (1) golden vectors produced by the REAL reference modules
(tests/golden/*.pt, generator: oracle/gen_golden.py) and (2) torchvision's independent ViT for the encoder whose source
is absent from the reference."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import fusion_ref, moco_ref, vit_ref


class TokenStub(torch.nn.Module):
    def __init__(self, dim, nc):
        super().__init__()
        self.head = torch.nn.Linear(dim, nc)

    def features3D(self, x):
        return x

    def forward(self, x):
        return self.head(x[:, 0])


@pytest.fixture(scope="module")
def fus_golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "fusion_ref.pt"))


def _build_fusion(g):
    v_c, v_e = TokenStub(g["dim"], 3), TokenStub(g["dim"], 3)
    v_c.load_state_dict(g["head_c"])
    v_e.load_state_dict(g["head_e"])
    m = fusion_ref.Fus_CrossViT(v_c, v_e, small_dim=g["dim"], large_dim=g["dim"], heads=g["heads"])
    m.load_state_dict(g["state_dict"], strict=True)
    return m, v_c, v_e


def test_fusion_restatement_matches_reference_forward_backward(fus_golden):
    g = fus_golden
    m, v_c, v_e = _build_fusion(g)
    t_c = g["t_c"].clone().requires_grad_(True)
    t_e = g["t_e"].clone().requires_grad_(True)
    fused, x_c, x_e = m(v_c, v_e, t_c, t_e)
    assert torch.equal(fused, g["fused"]) or torch.allclose(fused, g["fused"], atol=1e-6, rtol=0)
    assert torch.allclose(x_c, g["x_c"], atol=1e-6, rtol=0)
    assert torch.allclose(x_e, g["x_e"], atol=1e-6, rtol=0)
    loss = F.cross_entropy(fused + x_c + x_e, g["target"])
    assert abs(float(loss) - float(g["loss"])) < 1e-6
    loss.backward()
    assert torch.allclose(t_c.grad, g["d_t_c"], atol=1e-6, rtol=1e-5)
    assert torch.allclose(t_e.grad, g["d_t_e"], atol=1e-6, rtol=1e-5)
    for k, p in m.named_parameters():
        assert torch.allclose(p.grad, g["grads"][k], atol=1e-6, rtol=1e-5), k
    assert sum(p.numel() for p in m.parameters()) == g["n_params"]


def test_fusion_closed_form_and_dedup_equal_as_written(fus_golden):
    g = fus_golden
    m, v_c, v_e = _build_fusion(g)
    with torch.no_grad():
        fused, x_c, x_e = m(v_c, v_e, g["t_c"], g["t_e"])
        cf = m.closed_form(g["t_c"], g["t_e"])
        fused_d, x_c_d, x_e_d = m(v_c, v_e, g["t_c"], g["t_e"], dedup=True)
    assert torch.allclose(cf, g["fused"], atol=1e-6, rtol=0)
    assert torch.equal(fused_d, fused) and torch.equal(x_c_d, x_c) and torch.equal(x_e_d, x_e)


def test_fusion_key_inventory(golden_dir):
    inv = torch.load(os.path.join(golden_dir, "fusion_keys.pt"))
    m = fusion_ref.Fus_CrossViT(TokenStub(384, 3), TokenStub(384, 3))
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == inv["keys"]
    assert sum(p.numel() for p in m.parameters()) == inv["n_params"] == 1185798


def test_vit_restatement_matches_torchvision():
    """Independent oracle for the absent timm ViT: same function under the 1:1 key map."""
    tv = pytest.importorskip("torchvision.models.vision_transformer")
    torch.manual_seed(0)
    ours = vit_ref.VisionTransformerMoCo(img_size=32, patch_size=16, num_classes=5, embed_dim=64, depth=3, num_heads=2,
                                         mlp_ratio=4)
    with torch.no_grad():
        for p in ours.parameters():
            p.add_(torch.randn_like(p) * 0.02)
    ref = tv.VisionTransformer(image_size=32, patch_size=16, num_layers=3, num_heads=2, hidden_dim=64, mlp_dim=256,
                               num_classes=5)
    sd = {"conv_proj.weight": ours.patch_embed.proj.weight, "conv_proj.bias": ours.patch_embed.proj.bias,
          "class_token": ours.cls_token, "encoder.pos_embedding": ours.pos_embed,
          "encoder.ln.weight": ours.norm.weight, "encoder.ln.bias": ours.norm.bias,
          "heads.head.weight": ours.head.weight, "heads.head.bias": ours.head.bias}
    for i, blk in enumerate(ours.blocks):
        pre = "encoder.layers.encoder_layer_%d." % i
        sd.update({pre + "ln_1.weight": blk.norm1.weight, pre + "ln_1.bias": blk.norm1.bias,
                   pre + "self_attention.in_proj_weight": blk.attn.qkv.weight,
                   pre + "self_attention.in_proj_bias": blk.attn.qkv.bias,
                   pre + "self_attention.out_proj.weight": blk.attn.proj.weight,
                   pre + "self_attention.out_proj.bias": blk.attn.proj.bias,
                   pre + "ln_2.weight": blk.norm2.weight, pre + "ln_2.bias": blk.norm2.bias,
                   pre + "mlp.0.weight": blk.mlp.fc1.weight, pre + "mlp.0.bias": blk.mlp.fc1.bias,
                   pre + "mlp.3.weight": blk.mlp.fc2.weight, pre + "mlp.3.bias": blk.mlp.fc2.bias})
    ref.load_state_dict({k: v.detach() for k, v in sd.items()}, strict=True)
    ref.eval()
    x = torch.randn(4, 3, 32, 32)
    with torch.no_grad():
        a, b = ours(x), ref(x)
    assert torch.allclose(a, b, atol=2e-5, rtol=1e-5), float((a - b).abs().max())


def test_vit_small_inventory():
    m = vit_ref.vit_small(num_classes=3)
    assert sum(p.numel() for p in m.parameters()) == 21666819
    assert not m.pos_embed.requires_grad
    x = torch.randn(1, 3, 224, 224)
    with torch.no_grad():
        tok = m.features3D(x)
        assert tok.shape == (1, 197, 384)
        assert torch.allclose(m(x), m.head(tok[:, 0]))  # forward == head o features3D[:,0]  


@pytest.fixture(scope="module")
def moco_golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "moco_ref.pt"))


def _regen_queue(g):
    """The reference builds the queue with torch.randn under the recorded seed *after* constructing both encoders and the
    MLPs; re-running the constructor is not possible without the reference, so the fixture stores sampled columns and
    the test rebuilds only what it needs from them."""
    return g["queue_cols"]


def test_moco_logits_restatement(moco_golden):
    g = moco_golden
    cols = g["queue_cols"]  # queue[:, ::4096]
    logits, labels, qn, kn = moco_ref.infonce_logits(g["q_raw"], g["k_raw"], cols, g["T"])
    # column 0 = l_pos/T ; columns 1.. = q @ queue[:, ::4096] / T == reference logits[:, 1::4096]
    assert torch.allclose(logits[:, 0], g["logits_head"][:, 0], atol=1e-6, rtol=1e-5)
    assert torch.allclose(logits[:, 1:], g["logits_sampled"], atol=1e-6, rtol=1e-5)
    assert torch.equal(labels, g["labels"])
    # enqueue: queue[:, 0:B] = normalised keys transposed, ptr advanced by B
    assert torch.allclose(g["enqueued"], kn.t(), atol=1e-7, rtol=0)
    assert int(g["queue_ptr_after"]) == g["q_raw"].shape[0]


def test_moco_enqueue_restatement(moco_golden):
    g = moco_golden
    K, dim, B = 4096, g["dim"], g["k_raw"].shape[0]
    queue = torch.zeros(dim, K)
    ptr = torch.zeros(1, dtype=torch.long)
    kn = F.normalize(g["k_raw"], dim=1)
    for step in range(3):
        moco_ref.dequeue_and_enqueue(queue, ptr, kn, K)
    assert int(ptr) == 3 * B
    assert torch.equal(queue[:, :B], g["enqueued"]) or torch.allclose(queue[:, :B], g["enqueued"], atol=1e-7)
    assert torch.equal(queue[:, 2 * B:3 * B], queue[:, :B])


def test_ema_bit_exact_python_and_c(moco_golden):
    g = moco_golden
    # python restatement
    pk = [t.clone() for t in g["ema_before"]]
    holders = [torch.nn.Parameter(t) for t in pk]
    moco_ref.ema_update(holders, [torch.nn.Parameter(t) for t in g["ema_q"]], g["m"])
    for a, b in zip(holders, g["ema_after"]):
        assert torch.equal(a.data, b)
    # C restatement (three separately rounded fp32 ops, separately rounded m and 1-m)
    lib_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libema_ref.so")
    if not os.path.exists(lib_path):
        import subprocess
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(os.path.dirname(lib_path))])
    lib = ctypes.CDLL(lib_path)
    lib.ema_ref.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_float, ctypes.c_float]
    for k0, q, k1 in zip(g["ema_before"], g["ema_q"], g["ema_after"]):
        k = np.ascontiguousarray(k0.numpy().copy())
        qq = np.ascontiguousarray(q.numpy())
        lib.ema_ref(k.ctypes.data, qq.ctypes.data, k.size, g["m"], 1.0 - g["m"])
        assert np.array_equal(k, k1.numpy())


def test_cosine_momentum_schedule():
    # MAIN_PRE:626-629
    assert abs(moco_ref.cosine_momentum(0.0, 100, 0.99) - 0.99) < 1e-12
    assert abs(moco_ref.cosine_momentum(100.0, 100, 0.99) - 1.0) < 1e-12
    assert 0.99 < moco_ref.cosine_momentum(50.0, 100, 0.99) < 1.0


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree (authoring container)")
def test_whole_model_moco_restatement_equals_the_real_reference_class():
    """oracle/moco_ref.MoCoViT (used by the GPU parity tests as the MoCo step oracle) against the REAL MoCo_ViT.forward
    of BLD on the same tiny encoder, seed and inputs: logits, labels, queue and pointer are identical (the reference's
    batch shuffle permutes which rows share nothing at one rank, so results agree exactly)."""
    from functools import partial
    from types import SimpleNamespace

    import torch.distributed as dist
    from oracle import ref_loader
    _, _, bld_py = ref_loader.load_reference()
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29541")
        dist.init_process_group("gloo", rank=0, world_size=1)
    saved_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # BLD:121,194 hard-code .cuda()
    try:
        factory = partial(vit_ref.VisionTransformerMoCo, img_size=32, patch_size=16, embed_dim=64, depth=2, num_heads=2)
        torch.manual_seed(77)
        real = bld_py.MoCo_ViT(factory, SimpleNamespace(arch="vit_tiny"), 32, 48, 0.2)
        mine = moco_ref.MoCoViT(factory, 32, 48, 0.2)
        mine.load_state_dict(real.state_dict(), strict=True)
        with torch.no_grad():
            for a, b in zip(real.base_encoder.parameters(), mine.base_encoder.parameters()):
                d = torch.randn_like(a) * 0.01
                a.add_(d)
                b.add_(d)
        real.eval(), mine.eval()  # running statistics: the shuffle of BLD:171 cannot change anything
        im_q, im_k = torch.randn(8, 3, 32, 32), torch.randn(8, 3, 32, 32)
        for step in range(2):
            lr, yr = real(im_q, im_k, 0.99)
            lm, ym = mine(im_q, im_k, 0.99)
            assert torch.allclose(lr, lm, atol=1e-6), float((lr - lm).abs().max())
            assert torch.equal(yr, ym) and int(real.queue_ptr) == int(mine.queue_ptr) == 8 * (step + 1)
            assert torch.allclose(real.queue, mine.queue, atol=1e-7)
            for a, b in zip(real.momentum_encoder.parameters(), mine.momentum_encoder.parameters()):
                assert torch.equal(a, b)  # EMA bit-exact
    finally:
        torch.Tensor.cuda = saved_cuda


def test_pos_embed_against_an_independent_hand_written_sincos():
    """dropin/vits.py and oracle/vit_ref.py share their sin-cos code, and the torchvision cross-check copies
    pos_embed across - so the layout is pinned here by a scalar loop written from the MoCo-v3 definition alone
    (facebookresearch/moco-v3 vits.py, build_2d_sincos_position_embedding): `grid_w, grid_h = meshgrid(arange(w),
    arange(h))` with torch's default 'ij' indexing makes grid_w vary along the FIRST axis, and both grids are flattened
    row-major.  Token t of the flattened Conv2d output (patch row t // grid, patch column t % grid) therefore gets
    a = t // grid in the "w" blocks and b = t % grid in the "h" blocks:
        [sin(a w_d), cos(a w_d), sin(b w_d), cos(b w_d)],  w_d = 10000^(-d / (C/4)),
    i.e. upstream's "w" runs over patch ROWS - a quirk that is invisible on square grids but fixes which axis lands in
    which channel block; the class token gets a zero row."""
    import math

    import vits
    C, grid = 384, 14
    quarter = C // 4
    expect = torch.zeros(1 + grid * grid, C, dtype=torch.float64)
    for t in range(grid * grid):
        a, b = t // grid, t % grid
        for d in range(quarter):
            omega = 1.0 / (10000.0 ** (d / quarter))
            expect[1 + t, d] = math.sin(a * omega)
            expect[1 + t, quarter + d] = math.cos(a * omega)
            expect[1 + t, 2 * quarter + d] = math.sin(b * omega)
            expect[1 + t, 3 * quarter + d] = math.cos(b * omega)
    for mod in (vits.vit_small(), vit_ref.vit_small()):
        pe = mod.pos_embed.detach()[0].double()
        assert pe.shape == expect.shape and not mod.pos_embed.requires_grad
        assert (pe - expect).abs().max().item() < 2e-6
    # a 24 x 24 grid (384 x 384 inputs) follows the same rule
    pe = vits.vit_small(img_size=384).pos_embed.detach()[0].double()
    a, b, d = 17, 5, 3
    omega = 1.0 / (10000.0 ** (d / quarter))
    row = 1 + a * 24 + b
    assert abs(float(pe[row, d]) - math.sin(a * omega)) < 2e-6 and abs(float(pe[row, 2 * quarter + d]) - math.sin(b * omega)) < 2e-6


def test_schedules_follow_the_reference_formulas():
    """MAIN_PRE:608-629 and MAIN_CA:1043-1055, evaluated by hand at a few points."""
    import math

    from mfvit import schedules as S
    lr = S.base_lr(1.5e-4, 1024)                       # MAIN_PRE:286-288: lr * batch / 4
    assert abs(lr - 1.5e-4 * 256) < 1e-12 and S.base_lr(0.03, 256, cos=False) == 0.03
    assert S.pretrain_lr(0.0, lr, 100, 10) == 0.0 and abs(S.pretrain_lr(5.0, lr, 100, 10) - lr / 2) < 1e-12
    assert abs(S.pretrain_lr(10.0, lr, 100, 10) - lr) < 1e-12
    assert abs(S.pretrain_lr(55.0, lr, 100, 10) - lr * 0.5 * (1 + math.cos(math.pi * 45 / 90))) < 1e-12
    assert abs(S.pretrain_lr(100.0, lr, 100, 10)) < 1e-12
    assert abs(S.pretrain_lr(61.0, 1.0, 100, cos=False, schedule=(30, 60)) - 0.01) < 1e-12
    assert abs(S.moco_momentum(0.0, 100, 0.99) - 0.99) < 1e-12 and abs(S.moco_momentum(100.0, 100, 0.99) - 1.0) < 1e-12
    assert abs(S.moco_momentum(25.0, 100, 0.99) - moco_ref.cosine_momentum(25.0, 100, 0.99)) < 1e-15
    assert abs(S.finetune_lr(25, 0.1, 50, cos=True) - 0.05) < 1e-12 and S.finetune_lr(45, 0.1, 50, schedule=(20, 40)) == pytest.approx(0.001)
