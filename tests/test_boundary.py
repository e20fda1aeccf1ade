"""CPU: the drop-in boundary - the C-ABI library exports exactly what include/mfvit.h declares, and the Python module
surface matches what the reference scripts touch.  No compute calls (no GPU here)."""
import importlib
import os
import re
from functools import partial
from types import SimpleNamespace

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FUS_MOD = ("model.crossvit_2vits_2additionaloutputs_changenormlayer_location_removeextralclayer_"
           "changemodelinputlocation_std002_sum")


def _header_functions():
    src = open(os.path.join(ROOT, "include", "mfvit.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mfv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mfvit import _lib
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libmfvit.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES out of sync with include/mfvit.h"
    assert lib.mfv_abi_version() == 1
    assert b"unsupported shape" in lib.mfv_strerror(-1)


def test_ctypes_struct_sizes_match_header_layout():
    from mfvit import _lib
    import ctypes
    assert ctypes.sizeof(_lib.GemmArgs) == 7 * 8 + 13 * 8 + 8 * 4 + 8 + 4 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.FusionParams) == 13 * 2 * 8
    assert ctypes.sizeof(_lib.EmaChunk) == 24
    # mfv_vit_plan: 10 i64 + 4 ptr + 20 i64 + 2 ptr + 11 ptr + 4 i32 + 4 ptr + (1 + 2 + 2 + 7) ptr
    assert ctypes.sizeof(_lib.VitPlan) == (10 + 4 + 20 + 2 + 11) * 8 + 16 + 4 * 8 + (1 + 2 + 2 + 7) * 8


def test_cpu_forward_fails_loudly():
    import vits_returnftrs as vits
    from mfvit import MfvError
    m = vits.vit_small(num_classes=3)
    with pytest.raises(MfvError):
        m(torch.randn(1, 3, 224, 224))
    with pytest.raises(MfvError):
        m.blocks[0](torch.randn(1, 197, 384))


def test_vit_module_surface_and_init_parity():
    import vits
    import vits_returnftrs
    from oracle import vit_ref
    for name in ("vit_small", "vit_base", "vit_small_ori", "vit_base_ori", "vit_conv_small", "vit_conv_base"):
        assert name in vits.__dict__ and name in vits_returnftrs.__dict__  # MAIN_CA:56-57,289
    torch.manual_seed(7)
    ours = vits_returnftrs.vit_small()
    torch.manual_seed(7)
    ref = vit_ref.vit_small()
    sd_o, sd_r = ours.state_dict(), ref.state_dict()
    assert list(sd_o.keys()) == list(sd_r.keys())
    for k in sd_o:
        assert torch.equal(sd_o[k], sd_r[k]), k  # identical RNG consumption -> identical init
    assert [n for n, _ in ours.named_parameters()] == [n for n, _ in ref.named_parameters()]
    assert not ours.pos_embed.requires_grad and ours.head.in_features == 384
    # MAIN_CA:298-310: freeze all but head, replace head
    for n, p in ours.named_parameters():
        if n not in ("head.weight", "head.bias"):
            p.requires_grad = False
    ours.head = nn.Linear(ours.head.in_features, 3)
    assert [n for n, p in ours.named_parameters() if p.requires_grad] == ["head.weight", "head.bias"]
    # BLD:217-222: del head, assign an MLP
    hidden = ours.head.weight.shape[1]
    del ours.head
    ours.head = nn.Sequential(nn.Linear(hidden, 8, bias=False), nn.BatchNorm1d(8))
    assert "head.0.weight" in ours.state_dict()
    sg = vits.vit_small(stop_grad_conv1=True)
    assert not sg.patch_embed.proj.weight.requires_grad and not sg.patch_embed.proj.bias.requires_grad
    assert vits.vit_small_ori().num_heads == 12 and vits.vit_small().num_heads == 6
    with pytest.raises(NotImplementedError):
        vits.vit_conv_small()


def test_fusion_module_surface(golden_dir):
    import vits_returnftrs as vits
    fm = importlib.import_module(FUS_MOD)
    inv = torch.load(os.path.join(golden_dir, "fusion_keys.pt"))
    a, b = vits.vit_small(num_classes=3), vits.vit_small(num_classes=3)
    f = fm.Fus_CrossViT(a, b)
    assert {k: tuple(v.shape) for k, v in f.state_dict().items()} == inv["keys"]       # the reference's 22 keys
    assert sum(p.numel() for p in f.parameters()) == inv["n_params"] == 1185798        # backbones NOT included
    assert f.vit_features_cxr.__self__ is a and f.vit_features_enh.__self__ is b       # bound methods (FUS:80,83)
    # same seed -> same init as the reference restatement (self.apply order)
    from oracle import fusion_ref
    torch.manual_seed(3)
    f1 = fm.Fus_CrossViT(a, b)
    torch.manual_seed(3)
    f2 = fusion_ref.Fus_CrossViT(a, b)
    for (k1, v1), (k2, v2) in zip(f1.state_dict().items(), f2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    mod = importlib.import_module("model.module")
    for n in ("PreNorm", "CrossAttention", "Attention", "FeedForward"):
        assert hasattr(mod, n)  # FUS:6 imports these names


def test_moco_module_surface(golden_dir):
    import vits
    bm = importlib.import_module("moco.builder_vit_mocov3structure_mocov2loss")
    inv = torch.load(os.path.join(golden_dir, "moco_keys.pt"))
    m = bm.MoCo_ViT(partial(vits.vit_small, stop_grad_conv1=True), SimpleNamespace(arch="vit_small"), 256, 4096, 0.2)
    assert sorted(m.state_dict().keys()) == inv["keys"]
    assert sum(p.numel() for p in m.base_encoder.parameters()) == inv["n_base"] == 41080704
    assert sum(p.numel() for p in m.predictor.parameters()) == inv["n_pred"] == 2105344
    assert m.K == 65536 and m.queue.shape == (256, 65536) and m.queue_ptr.dtype == torch.long
    assert all(not p.requires_grad for p in m.momentum_encoder.parameters())
    for pb, pm in zip(m.base_encoder.parameters(), m.momentum_encoder.parameters()):
        assert torch.equal(pb, pm)
    assert hasattr(bm, "concat_all_gather") and hasattr(bm, "MoCo_ResNet")
    # survives SyncBatchNorm conversion (MAIN_PRE:297)
    m2 = nn.SyncBatchNorm.convert_sync_batchnorm(m)
    assert any(isinstance(x, nn.SyncBatchNorm) for x in m2.modules())


def test_layout_is_aligned_and_complete():
    import vits
    from mfvit.engine import ViTLayout
    m = vits.vit_small(num_classes=3)
    lay = ViTLayout(224, 16, 384, 12, 6, 1536)
    names = [n for n, _ in m.named_parameters() if not n.startswith("head.")]
    assert [n for n, _, _ in lay.entries] == names
    assert all(o % 8 == 0 for _, _, o in lay.entries) and lay.P % 8 == 0
    assert sum(p.numel() for n, p in m.named_parameters() if not n.startswith("head.")) <= lay.P
    for n, shape, _ in lay.entries:
        assert tuple(m.get_parameter(n).shape) == shape
    # blocks are equally strided (the native runtime addresses block l at off_block0 + l*block_stride)
    for i in range(12):
        assert lay.offset["blocks.%d.norm1.weight" % i] == lay.off_block0 + i * lay.block_stride


def test_checkpoint_flow_pretrain_to_probe_to_fusion(tmp_path):
    """The reference hands weights from script to script through state dicts: MoCo pretraining -> `module.base_encoder.`
    prefix stripped, head dropped (MAIN_LPFT:327-337, missing keys must be exactly head.weight/head.bias) -> fine-tuned
    `model_best.pth.tar` -> strict load into both MF-ViT CA branches (MAIN_CA:357,385).  The drop-in keeps every key."""
    import vits
    bm = importlib.import_module("moco.builder_vit_mocov3structure_mocov2loss")
    moco = bm.MoCo_ViT(partial(vits.vit_small, stop_grad_conv1=True), SimpleNamespace(arch="vit_small"), 256, 4096, 0.2)
    ckpt = {"state_dict": {"module." + k: v for k, v in moco.state_dict().items()}}  # DDP-wrapped, as MAIN_PRE saves it
    path = os.path.join(tmp_path, "checkpoint_smallest_loss.pth.tar")
    torch.save(ckpt, path)
    # linear-probe script: rename, drop the projector head, non-strict load
    linear_keyword = "head"
    sd = torch.load(path, map_location="cpu")["state_dict"]
    for k in list(sd.keys()):
        if k.startswith("module.base_encoder") and not k.startswith("module.base_encoder.%s" % linear_keyword):
            sd[k[len("module.base_encoder."):]] = sd[k]
        del sd[k]
    probe = vits.vit_small(num_classes=3)
    msg = probe.load_state_dict(sd, strict=False)
    assert set(msg.missing_keys) == {"head.weight", "head.bias"} and not msg.unexpected_keys
    for n, p in probe.named_parameters():
        if not n.startswith("head."):
            assert torch.equal(p, moco.base_encoder.state_dict()[n]), n
    # fusion script: strict load of the fine-tuned single-branch checkpoint into a fresh branch with a 3-class head
    best = os.path.join(tmp_path, "model_best.pth.tar")
    torch.save({"state_dict": probe.state_dict()}, best)
    branch = vits.vit_small()
    branch.head = nn.Linear(branch.head.in_features, 3)  # MAIN_CA:309
    branch.load_state_dict(torch.load(best, map_location="cpu")["state_dict"])  # strict, MAIN_CA:357
    assert torch.equal(branch.head.weight, probe.head.weight)


def test_checkpoint_helpers_follow_the_reference_hand_offs(tmp_path):
    """mfvit.checkpoint restates MAIN_LPFT:318-340 (rename + non-strict load), MAIN_CA:343-385 (strict branch load) and
    MAIN_CA:1002-1011 (file naming) as functions."""
    import vits
    from mfvit import checkpoint as ck
    bm = importlib.import_module("moco.builder_vit_mocov3structure_mocov2loss")
    moco = bm.MoCo_ViT(partial(vits.vit_small, stop_grad_conv1=True), SimpleNamespace(arch="vit_small"), 256, 4096, 0.2)
    pre = ck.save_checkpoint(str(tmp_path), {"state_dict": {"module." + k: v for k, v in moco.state_dict().items()}},
                             is_best=False, filename="checkpoint_smallest_loss.pth.tar")
    assert pre.endswith("checkpoint_smallest_loss.pth.tar")
    probe = vits.vit_small(num_classes=3)
    ck.load_pretrained_backbone(probe, pre)
    assert torch.equal(probe.blocks[3].attn.qkv.weight, moco.base_encoder.blocks[3].attn.qkv.weight)
    assert ck.sanity_check_frozen(probe.state_dict(), pre)
    with torch.no_grad():
        probe.blocks[0].mlp.fc1.bias.add_(1.0)
    with pytest.raises(AssertionError):
        ck.sanity_check_frozen(probe.state_dict(), pre)
    with pytest.raises(RuntimeError):  # a checkpoint of another architecture
        ck.load_pretrained_backbone(vits.vit_small(num_classes=3), {"state_dict": {"module.base_encoder.cls_token": torch.zeros(1, 1, 384)}})
    best = ck.save_checkpoint(str(tmp_path), {"state_dict": {"module." + k: v for k, v in probe.state_dict().items()}}, is_best=True)
    assert os.path.basename(best) == "model_best.pth.tar"
    branch = vits.vit_small()
    branch.head = nn.Linear(branch.head.in_features, 3)
    ck.load_finetuned_branch(branch, best)
    assert torch.equal(branch.blocks[0].mlp.fc1.bias, probe.blocks[0].mlp.fc1.bias)
    sd = ck.strip_moco_prefix({"module.base_encoder.head.0.weight": 1, "module.momentum_encoder.x": 2, "module.queue": 3,
                               "module.base_encoder.norm.weight": 4})
    assert sd == {"norm.weight": 4}


def test_noprediction_builder_surface():
    """BLD_NOPRED differs from BLD in one line (keys skip the predictor); same classes, keys and buffers."""
    import vits
    a = importlib.import_module("moco.builder_vit_mocov3structure_mocov2loss")
    b = importlib.import_module("moco.builder_vit_mocov3structure_mocov2loss_noprediction_q")
    for n in ("MoCo", "MoCo_ViT", "MoCo_ResNet", "concat_all_gather"):
        assert hasattr(b, n)
    mk = lambda mod: mod.MoCo_ViT(partial(vits.vit_small, stop_grad_conv1=True), SimpleNamespace(arch="vit_small"), 256, 4096, 0.2)  # noqa: E731
    ma, mb = mk(a), mk(b)
    assert sorted(ma.state_dict().keys()) == sorted(mb.state_dict().keys())
    assert ma.predictor_on_keys and not mb.predictor_on_keys and isinstance(mb, a.MoCo)


def test_checkpoints_record_the_head_split_and_loaders_refuse_a_mismatch(tmp_path):
    """qkv.weight is [1152, 384] for 6 heads of 64 and for 12 heads of 32: a state dict cannot tell, so
    checkpoints written with model= carry the head count, a mismatch raises, and an unmarked checkpoint warns."""
    import vits
    from mfvit import checkpoint as ck
    six = vits.vit_small(num_classes=3)
    path = ck.save_checkpoint(str(tmp_path), {"state_dict": six.state_dict()}, is_best=True, model=six)
    assert torch.load(path)["mfvit_arch"]["num_heads"] == 6
    ck.load_finetuned_branch(vits.vit_small(num_classes=3), path)  # same split: silent
    with pytest.raises(RuntimeError, match="attention heads"):
        ck.load_finetuned_branch(vits.vit_small_ori(num_classes=3), path)
    unmarked = ck.save_checkpoint(str(tmp_path), {"state_dict": six.state_dict()}, is_best=False)
    with pytest.warns(UserWarning, match="12 heads of 32"):
        ck.load_finetuned_branch(vits.vit_small(num_classes=3), unmarked)


def test_shadow_freshness_follows_parameter_versions():
    """a `fresh` flag set by the optimizer step must not survive an in-place edit of the fp32 master
    (load_state_dict, p.copy_).  Host logic only: no kernels are launched here."""
    import vits
    from mfvit.engine import ViTEngine
    m = vits.vit_small(num_classes=3)
    eng = ViTEngine([m])
    eng._params = [eng._named_params(0)]  # what adopt() records, without touching a device
    eng.mark_shadow_fresh()
    assert eng.shadow_is_current()
    with torch.no_grad():
        m.blocks[0].mlp.fc1.bias.add_(1.0)  # what torch.optim / p.copy_ do: bumps Parameter._version
    assert not eng.shadow_is_current()
    eng.mark_shadow_fresh()
    m.load_state_dict(vits.vit_small(num_classes=3).state_dict())
    assert not eng.shadow_is_current() and eng._shadow_sig is None  # post-hook invalidated it
    eng.mark_shadow_fresh()
    eng.invalidate_shadow()  # what raw writers (EMA kernel, p.data edits) must call
    assert not eng.shadow_is_current()


def test_label_tensors_are_validated_before_the_kernel_sees_them():
    from mfvit import MfvError, ops
    logits = torch.zeros(4, 3)
    for bad in (torch.zeros(4, dtype=torch.int32), torch.zeros(5, dtype=torch.int64), torch.zeros(4, 1, dtype=torch.int64)):
        with pytest.raises(MfvError, match="int64"):
            ops._check_labels(bad, 4, logits.device, "ce_small")
    ops._check_labels(torch.zeros(4, dtype=torch.int64), 4, logits.device, "ce_small")


def test_trainer_default_is_the_reference_optimisation_set():
    """MAIN_CA:435-449: the optimizer is built over Fus_CrossViT.parameters() - 22 tensors; backbones and their heads are
    never stepped.  Checks the packed-buffer ranges the fused optimizer kernels are given."""
    import vits_returnftrs as vits
    from mfvit.trainer import FlatParams, MFViTCATrainer
    fm = importlib.import_module(FUS_MOD)
    cxr, enh = vits.vit_small(num_classes=3), vits.vit_small(num_classes=3)
    fus = fm.Fus_CrossViT(cxr, enh)
    n_fusion = sum((p.numel() + 3) // 4 * 4 for p in fus.parameters())
    for full in (False, True):
        tr = MFViTCATrainer(fus, cxr, enh, train_backbones=full)
        params, _ = fus._fusion_params(cxr, enh)
        tr._small = FlatParams(params, torch.device("cpu"))
        covered = sum(hi - lo for lo, hi in tr._small_ranges())
        assert covered == (tr._small.n if full else n_fusion)
    fus.mlp_head_cxr[0].bias.requires_grad = False  # requires_grad is honoured inside the packed buffer
    tr = MFViTCATrainer(fus, cxr, enh)
    tr._small = FlatParams(fus._fusion_params(cxr, enh)[0], torch.device("cpu"))
    assert sum(hi - lo for lo, hi in tr._small_ranges()) == n_fusion - 4
