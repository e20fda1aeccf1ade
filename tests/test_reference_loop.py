"""The reference's OWN training loop executed over the drop-in (SURVEY 8(b) "scripts run unmodified").

baseline/_ref/reference is an unmodified copy of the reference tree (staged by __graft_entry__.build() from
/root/reference, git-ignored, travels to the GPU box).  tools/overlay.py lays dropin/ over its moco_pretraining/moco/
directory exactly as INTEGRATION.md section 2 tells a maintainer to; the MAIN_CA script is then imported AS A MODULE -
every one of its imports (`vits_returnftrs`, `model.crossvit_..._sum`, `moco.loader`, `training_tools`, `aihc_utils`,
`config`) resolving inside that deployed tree - and its `train()` (MAIN_CA:793-926) is called twice with the same
synthetic loaders: once over the oracle modules (plain PyTorch fp32) and once over the drop-in (libmfvit.so kernels).
"""
import importlib.util
import os
import sys
import types

import pytest
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import e2e_common as E  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref", "reference")
MAIN_CA = ("main_vit_covid_test_val_single_img_type_5draws_rev_v2loss_v3structure_crossvit_2vits_2additionaloutputs_"
           "trainval_sum.py")


def _import_reference_script(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "multi-feature-vit_b200", "tools"))
    import overlay
    import_root = overlay.make_overlay(STAGED, str(tmp_path / "deploy"))
    for n in ("matplotlib", "matplotlib.pyplot"):  # imported by the script, never used by train(); absent in this image
        sys.modules.setdefault(n, types.ModuleType(n))
    # the deployed tree must win over the in-repo dropin/ directory for every module the script imports
    saved_path, saved_mods = list(sys.path), dict(sys.modules)
    for name in list(sys.modules):
        if name.split(".")[0] in ("vits", "vits_returnftrs", "model", "moco", "_path"):
            del sys.modules[name]
    sys.path.insert(0, import_root)
    try:
        spec = importlib.util.spec_from_file_location("reference_main_ca", os.path.join(str(tmp_path / "deploy"), MAIN_CA))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert os.path.realpath(mod.vits.__file__).startswith(os.path.realpath(import_root))
        assert os.path.realpath(mod.Fus_CrossViT.__init__.__code__.co_filename).startswith(os.path.realpath(import_root))
        return mod
    finally:
        sys.path[:] = saved_path
        for name in list(sys.modules):
            if name.split(".")[0] in ("vits", "vits_returnftrs", "model", "moco", "_path") and name not in saved_mods:
                sys.modules.pop(name)
        sys.modules.update({k: v for k, v in saved_mods.items()
                            if k.split(".")[0] in ("vits", "vits_returnftrs", "model", "moco", "_path")})


class _Writer:
    def __init__(self):
        self.rows = []

    def add_scalar(self, tag, value, step):
        self.rows.append((tag, float(value), step))


def _loaders(n_batches, B):
    """What MAIN_CA:573-578,664-669 hand to train(): per phase, a loader of ((view1, view2), target) per image type."""
    cxr, enh = {}, {}
    for pi, phase in enumerate(("train", "val")):
        cxr[phase], enh[phase] = [], []
        for i in range(n_batches):
            c, e, t = E.synthetic_pair(B, 224, rank=10 * pi + i)
            cxr[phase].append(((c, c), t.int()))  # the script casts with target.long() (MAIN_CA:859)
            enh[phase].append(((e, e), t.int()))
    return cxr, enh


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(STAGED), reason="baseline/_ref/reference not staged (run __graft_entry__.build() "
                                                      "where /root/reference exists)")
def test_reference_train_function_runs_unmodified_over_the_dropin(tmp_path):
    ref_main = _import_reference_script(tmp_path)
    torch.backends.cuda.matmul.allow_tf32 = False
    n_batches, B = 2, 8
    results = {}
    for which in ("oracle", "dropin"):
        (fus, cxr, enh), _ = E.build_mfvit_pair(seed=31)
        if which == "dropin":  # built the way the deployed script builds them (MAIN_CA:289-316,393), same weights
            mine = []
            for r in (cxr, enh):
                v = ref_main.vits.__dict__["vit_small"]()
                v.head = nn.Linear(v.head.in_features, 3)
                v.load_state_dict(r.state_dict(), strict=True)
                mine.append(v.cuda())
            f = ref_main.Fus_CrossViT(mine[0], mine[1])
            f.load_state_dict(fus.state_dict(), strict=True)
            fus, cxr, enh = f.cuda(), mine[0], mine[1]
            for obj in (fus, cxr):  # really the deployed tree's classes, not the in-repo dropin/ directory
                assert str(tmp_path / "deploy") in type(obj).__init__.__code__.co_filename
        optimizer = torch.optim.SGD(fus.parameters(), 0.01, momentum=0.9, weight_decay=0.0)  # MAIN_CA:435-449
        args = types.SimpleNamespace(semi_supervised=True)
        loaders, loaders_enh = _loaders(n_batches, B)
        num_imgs = {"train": n_batches * B, "val": n_batches * B}
        writer = _Writer()
        out = ref_main.train(loaders, loaders_enh, fus, cxr, enh, nn.CrossEntropyLoss().cuda(), optimizer, 0, args,
                             num_imgs, writer)
        epoch_loss, epoch_auc, epoch_acc = out[0], out[1], out[2]
        results[which] = (float(epoch_loss), float(epoch_auc), float(epoch_acc), dict((t, v) for t, v, _ in writer.rows),
                          {n: p.detach().clone() for n, p in fus.named_parameters()},
                          cxr.blocks[0].attn.qkv.weight.grad.detach().clone())
    (l0, a0, c0, w0, p0, g0), (l1, a1, c1, w1, p1, g1) = results["oracle"], results["dropin"]
    assert abs(l0 - l1) <= 2e-3, (l0, l1)                       # val loss of the epoch (MAIN_CA:907)
    assert abs(w0["train/loss"] - w1["train/loss"]) <= 2e-3
    assert abs(c0 - c1) <= 1.0 / (n_batches * B) + 1e-9 and abs(a0 - a1) <= 0.05
    for n in p0:                                                # the optimizer stepped the fusion's tensors identically
        assert (p0[n] - p1[n]).abs().max().item() <= 2e-4, n
    assert E.cos(g0, g1) >= 0.999                               # backbones received gradients (never stepped, fact 4)
