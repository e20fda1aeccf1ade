"""One epoch of the MF-ViT CA loop, both phases (MAIN_CA:824-909), with the host round trips removed (SURVEY 8(f) rows 2
and 4): the `train` phase is MFViTCATrainer.step per batch, the `val` phase a forward + loss per batch; loss sum, argmax
hits and the raw summed logits accumulate on the device (mfv_epoch_metrics) and are read once per phase.

    for epoch in range(epochs):
        train_loss, train_auc, train_acc = run_phase("train", trainer, loaders["train"], metrics, num_imgs["train"])
        val_loss, val_auc, val_acc       = run_phase("val",   trainer, loaders["val"],   metrics, num_imgs["val"])
"""
import torch

from . import ops
from ._lib import MfvError


@torch.no_grad()
def evaluate_batch(trainer, img_cxr, img_enh, target):
    """Forward of both encoders + fusion + heads and the cross-entropy on fused + x_cxr + x_enh (MAIN_CA:862-873), no
    backward, no parameter update.  Returns (loss [1], (fused, x_cxr, x_enh)) as device tensors."""
    trainer._prepare(img_cxr.device)
    eng, lay = trainer.engine, trainer.engine.layout
    tok, _ = eng.forward([img_cxr, img_enh], save=False)
    fused, x = ops.fusion_fwd(tok, trainer._pstruct, img_cxr.shape[0], lay.S, lay.C, trainer.heads, trainer.NC)
    loss, _ = ops.ce_small(fused, x[0], x[1], target.long(), want_grad=False)
    return loss, (fused, x[0], x[1])


def run_phase(phase, trainer, loader, metrics, num_imgs=None):
    """phase 'train': forward + backward + optimizer step per batch; 'val' / 'test': forward only.  Returns
    (epoch_loss, epoch_auc, epoch_acc) exactly as MAIN_CA:905-907 defines them."""
    if phase not in ("train", "val", "test"):
        raise MfvError("unknown phase %r" % (phase,))
    metrics.reset()
    if phase == "train":
        if trainer.metrics is not metrics:
            raise MfvError("the trainer must have been built with metrics=<this EpochMetrics> (it is part of the "
                           "captured step)")
        for img_cxr, img_enh, target in loader:
            trainer.step(img_cxr, img_enh, target)
    else:
        for img_cxr, img_enh, target in loader:
            loss, (fused, x_cxr, x_enh) = evaluate_batch(trainer, img_cxr, img_enh, target)
            metrics.accumulate(fused, x_cxr, x_enh, target, loss)
    return metrics.result(num_imgs)
