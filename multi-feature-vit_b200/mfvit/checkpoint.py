"""Checkpoint hand-offs between the reference's three scripts, as functions instead of inline code:

  MoCo pretraining  --checkpoint_*.pth.tar-->  linear probe / fine-tune  --model_best.pth.tar-->  MF-ViT CA

Key names are the reference's; nothing here is specific to the CUDA engine - the drop-in modules keep the
reference's state-dict layout, so author / Kaggle checkpoints load unchanged.

One thing a state dict cannot say: the self-attention head split.  qkv.weight is [1152, 384] for 6 heads of 64 (the
`vit_small` of this package, north_star) AND for 12 heads of 32 (upstream MoCo-v3 `vit_small`, `vit_small_ori` here), so a
checkpoint trained with one loads cleanly into the other and computes different numbers.  Checkpoints written through
`save_checkpoint(..., model=...)` carry the head count ("mfvit_arch") and the loaders below refuse a mismatch; for a
checkpoint without that record (author / Kaggle files) they warn, naming the head count in use."""
import os
import warnings

import torch


def arch_record(model):
    """What a state dict does not pin down."""
    return {"num_heads": int(model.num_heads), "embed_dim": int(model.embed_dim), "depth": int(model.depth),
            "img_size": int(model.img_size)}


def _check_arch(model, checkpoint, what):
    heads = getattr(model, "num_heads", None)
    if heads is None:
        return
    rec = checkpoint.get("mfvit_arch") if isinstance(checkpoint, dict) else None
    if rec is None:
        warnings.warn("%s: the checkpoint does not record its attention head count; loading into a model with %d heads "
                      "of %d.  Upstream MoCo-v3 vit_small checkpoints use 12 heads of 32 (build the model with "
                      "vits.vit_small_ori() or num_heads=12); the weight shapes are identical either way, so a wrong "
                      "choice is silent." % (what, heads, model.embed_dim // heads), stacklevel=3)
    elif int(rec.get("num_heads", heads)) != int(heads):
        raise RuntimeError("%s: checkpoint was written by a model with %d attention heads, this model has %d"
                           % (what, int(rec["num_heads"]), int(heads)))


def strip_moco_prefix(state_dict, linear_keyword="head", prefix="module.base_encoder."):
    """MAIN_LPFT:327-335: keep `module.base_encoder.*` except the projector that replaced the head, drop the rest
    (momentum encoder, predictor, queue), remove the prefix.  Returns a new dict."""
    out = {}
    stem = prefix[:-1]
    for k, v in state_dict.items():
        if k.startswith(stem) and not k.startswith(prefix + linear_keyword):
            out[k[len(prefix):]] = v
    return out


def load_pretrained_backbone(model, checkpoint, linear_keyword="head"):
    """MAIN_LPFT:318-340: non-strict load of a MoCo pretraining checkpoint (path or dict) into a fresh ViT; the only
    keys allowed to be missing are the classifier's."""
    if isinstance(checkpoint, (str, os.PathLike)):
        checkpoint = torch.load(checkpoint, map_location="cpu")
    _check_arch(model, checkpoint, "load_pretrained_backbone")
    sd = strip_moco_prefix(checkpoint["state_dict"] if "state_dict" in checkpoint else checkpoint, linear_keyword)
    msg = model.load_state_dict(sd, strict=False)
    missing = set(msg.missing_keys)
    if missing != {"%s.weight" % linear_keyword, "%s.bias" % linear_keyword} or msg.unexpected_keys:
        raise RuntimeError("pretrained checkpoint does not match the backbone: missing %s, unexpected %s"
                           % (sorted(missing), sorted(msg.unexpected_keys)))
    return msg


def load_finetuned_branch(model, checkpoint):
    """MAIN_CA:343-357 / 371-385: strict load of a single-branch `model_best.pth.tar` (head already replaced by the
    3-class Linear, MAIN_CA:309) into one MF-ViT CA branch."""
    if isinstance(checkpoint, (str, os.PathLike)):
        checkpoint = torch.load(checkpoint, map_location="cpu")
    _check_arch(model, checkpoint, "load_finetuned_branch")
    sd = checkpoint["state_dict"] if "state_dict" in checkpoint else checkpoint
    if any(k.startswith("module.") for k in sd):  # saved from a DataParallel / DDP wrapper
        sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}
    return model.load_state_dict(sd)


def save_checkpoint(checkpoint_folder, state, is_best, filename="last_checkpoint.pth.tar", model=None):
    """MAIN_CA:1002-1011: the best model is written as model_best.pth.tar, anything else under `filename`.  With
    `model` (a ViT of this package, or a MoCo wrapper holding one as base_encoder) the head count is recorded too."""
    os.makedirs(checkpoint_folder, exist_ok=True)
    vit = getattr(model, "base_encoder", model)
    if vit is not None and hasattr(vit, "num_heads"):
        state = dict(state, mfvit_arch=arch_record(vit))
    path = os.path.join(checkpoint_folder, "model_best.pth.tar" if is_best else filename)
    torch.save(state, path)
    return path


def sanity_check_frozen(state_dict, pretrained, linear_keyword="head", prefix="module.base_encoder."):
    """MAIN_LPFT's sanity_check: a linear probe must leave every backbone tensor bit-identical to the pretrained one."""
    if isinstance(pretrained, (str, os.PathLike)):
        pretrained = torch.load(pretrained, map_location="cpu")
    pre = pretrained["state_dict"] if "state_dict" in pretrained else pretrained
    for k, v in state_dict.items():
        if "%s.weight" % linear_keyword in k or "%s.bias" % linear_keyword in k:
            continue
        k_pre = prefix + (k[len("module."):] if k.startswith("module.") else k)
        if not torch.equal(v.cpu(), pre[k_pre].cpu()):
            raise AssertionError("%s is changed in linear classifier training." % k)
    return True
