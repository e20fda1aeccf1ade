"""Per-iteration schedules of the pretraining script (MAIN_PRE:608-629) and the fine-tuning scripts (MAIN_CA:1043-1055),
as pure functions of the fractional epoch.  The callers pass `epoch + i / iters_per_epoch` exactly as MAIN_PRE:512-527
does; the results go to MFViTCATrainer.set_lr / MoCoPretrainer.set_lr (a device scalar: no re-capture of the step) and to
the `m` argument of MoCo.forward."""
import math


def base_lr(lr, batch_size, cos=True):
    """MAIN_PRE:286-290: with --cos the learning rate is scaled by (global batch / 4); otherwise used as given."""
    return lr * batch_size / 4 if cos else lr


def pretrain_lr(epoch, lr, epochs, warmup_epochs=0, cos=True, schedule=()):
    """MAIN_PRE:608-623 adjust_learning_rate: linear warm-up over `warmup_epochs`, then a half-cycle cosine to zero at
    `epochs`; without --cos a step schedule (x0.1 at every milestone reached).  `epoch` may be fractional."""
    if cos:
        if epoch < warmup_epochs:
            return lr * epoch / warmup_epochs
        return lr * 0.5 * (1. + math.cos(math.pi * (epoch - warmup_epochs) / (epochs - warmup_epochs)))
    out = lr
    for milestone in schedule:
        out *= 0.1 if epoch >= milestone else 1.
    return out


def moco_momentum(epoch, epochs, moco_m):
    """MAIN_PRE:626-629 adjust_moco_momentum: m rises from moco_m to 1 along a half cosine."""
    return 1. - 0.5 * (1. + math.cos(math.pi * epoch / epochs)) * (1. - moco_m)


def finetune_lr(epoch, lr, epochs, cos=False, schedule=()):
    """MAIN_CA:1043-1055 / MAIN_LPFT adjust_learning_rate: per-epoch half cosine (no warm-up) or x0.1 milestones."""
    if cos:
        return lr * 0.5 * (1. + math.cos(math.pi * epoch / epochs))
    out = lr
    for milestone in schedule:
        out *= 0.1 if epoch >= milestone else 1.
    return out
