"""Sync-free single-branch linear probe / fine-tuning (MAIN_LPFT:647-826; BASELINE configs[0] shape).

    model = vits.vit_small(); model.head = nn.Linear(384, 3)                  # MAIN_LPFT:276-296
    # linear probe: everything but head.weight / head.bias frozen (MAIN_LPFT:283-286); --semi-supervised: nothing frozen
    trainer = ViTClassifierTrainer(model, lr=init_lr, momentum=0.9, weight_decay=wd, metrics=EpochMetrics(...))
    loss = trainer.step(images, target)                 # train phase of MAIN_LPFT:696-718 for one batch, no host sync
    loss, logits = trainer.evaluate(images, target)     # val phase / test() (MAIN_LPFT:765-826): forward + loss only
    epoch_loss, epoch_auc, epoch_acc = run_phase("train" | "val" | "test", trainer, loader, metrics, num_imgs)

The optimizer set is the reference's: `filter(lambda p: p.requires_grad, model.parameters())` (MAIN_LPFT:380-392) - two
tensors for a linear probe (asserted there), every tensor with --semi-supervised.  One step = mfv_vit_forward (G = 1) ->
head (mfv_linear_small_fwd) -> mfv_ce_small -> head backward -> [mfv_vit_backward when the backbone trains] -> fused
flat-buffer optimizer step; with a frozen backbone the encoder forward keeps no activations and no backward runs.
"""
import torch
import torch.nn as nn

from . import ops
from ._lib import MfvError
from .engine import engine_for
from .trainer import FlatParams


class ViTClassifierTrainer:
    def __init__(self, model, lr=0.1, momentum=0.9, weight_decay=0.0, optimizer="sgd", betas=(0.9, 0.999), eps=1e-8,
                 metrics=None):
        if optimizer not in ("sgd", "adam", "adamw"):
            raise MfvError("optimizer must be 'sgd', 'adam' or 'adamw', not %r" % (optimizer,))
        if not isinstance(getattr(model, "head", None), nn.Linear) or model.head.out_features > 32:
            raise MfvError("ViTClassifierTrainer needs model.head = nn.Linear(embed_dim, num_classes <= 32) (MAIN_LPFT:288)")
        self.model, self.engine = model, engine_for(model)
        self.lr, self.momentum, self.wd = float(lr), momentum, weight_decay
        self.optimizer, self.betas, self.eps = optimizer, tuple(betas), eps
        self.metrics = metrics
        self.NC = model.head.out_features
        self.steps = 0
        self._ready = False

    def _prepare(self, device):
        if self._ready:
            return
        eng = self.engine
        if not eng.is_adopted() or eng.device != device:
            eng.adopt(device)
        self.model.head.to(device)
        self._head = FlatParams([self.model.head.weight, self.model.head.bias], device)
        z = torch.zeros_like
        self._m_head, self._v_head = z(self._head.master), z(self._head.master)
        self._m_eng = z(eng.master)
        self._v_eng = z(eng.master) if self.optimizer != "sgd" else None
        self._lr_dev = torch.full((1,), self.lr, device=device, dtype=torch.float32)
        self._step_dev = torch.zeros(1, device=device, dtype=torch.int64)
        # contiguous trainable runs of the encoder's flat layout (empty for a linear probe)
        lay, runs, run = eng.layout, [], None
        for (name, shape, off), (_, p) in zip(lay.entries, eng._params[0]):
            if p.requires_grad:
                if run is not None and run[1] >= off - 8:
                    run[1] = off + p.numel()
                else:
                    if run is not None:
                        runs.append(run)
                    run = [off, off + p.numel()]
            elif run is not None:
                runs.append(run)
                run = None
        if run is not None:
            runs.append(run)
        self._runs = [(lo, min((hi + 3) // 4 * 4, lay.P)) for lo, hi in runs]
        self._head_trains = [self.model.head.weight.requires_grad, self.model.head.bias.requires_grad]
        self._shadow_complete = False
        self._dtok = {}
        self._ready = True

    def set_lr(self, lr):
        """adjust_learning_rate (MAIN_LPFT:870-882): takes effect at the next step."""
        self.lr = float(lr)
        if self._ready:
            self._lr_dev.fill_(self.lr)

    def _forward(self, img, save):
        eng, lay = self.engine, self.engine.layout
        tok, lease = eng.forward([img], save=save)
        B = img.shape[0]
        head = self._head
        logits = ops.linear_small_fwd(tok, lay.S * lay.C, head.views[0], head.views[1], B)
        return tok, lease, logits

    def _update(self, p, g, m, v, shadow, shadow16):
        if self.optimizer != "sgd":
            ops.adam_step_dev_(p, g, m, v, shadow, self._lr_dev, self.betas, self.eps, self.wd,
                               self.optimizer == "adamw", self._step_dev, shadow16=shadow16)
        else:
            ops.sgd_step_dev_(p, g, m, shadow, self._lr_dev, self.momentum, self.wd, self.steps == 0, shadow16=shadow16)

    def step(self, img, target):
        """optimizer.zero_grad(); output = model(images); loss = criterion(output, target); loss.backward();
        optimizer.step()   (MAIN_LPFT:704-718).  Returns the loss as a 1-element device tensor."""
        device = img.device
        if device.type != "cuda":
            raise MfvError("ViTClassifierTrainer runs only on CUDA sm_100a devices (no CPU fallback)")
        self._prepare(device)
        eng, lay = self.engine, self.engine.layout
        B = img.shape[0]
        target = target.long()
        train_enc = bool(self._runs)
        tok, lease, logits = self._forward(img, save=train_enc)
        loss, dlogits = ops.ce_small(logits, None, None, target)
        head = self._head
        ops.fill_(head.grad, 0.0)
        dtok = None
        if train_enc:
            dtok = self._dtok.get(B)
            if dtok is None:
                dtok = self._dtok[B] = torch.empty_like(tok)
            ops.fill_(dtok.view(-1), 0.0)
        ops.linear_small_bwd(tok, lay.S * lay.C, head.views[0], dlogits, dtok, lay.S * lay.C, head.gviews[0],
                             head.gviews[1], B)
        if self.optimizer != "sgd":
            self._step_dev.add_(1)
        if train_enc:
            grad = eng.backward(lease, dtok)
            adam = self.optimizer != "sgd"
            for lo, hi in self._runs:
                sl = slice(lo, hi)
                self._update(eng.master[0, sl], grad[0, sl], self._m_eng[0, sl], self._v_eng[0, sl] if adam else None,
                             eng.shadow[0, sl], eng.shadow16[0, sl] if eng.fwd_f16 else None)
            if not self._shadow_complete:
                eng.cast_shadow()
                self._shadow_complete = True
        eng.mark_shadow_fresh()  # rewritten by the step, or untouched (frozen backbone)
        off = 0
        for p, trains in zip(head.params, self._head_trains):
            n = (p.numel() + 3) // 4 * 4
            if trains:
                sl = slice(off, off + n)
                self._update(head.master[sl], head.grad[sl], self._m_head[sl],
                             self._v_head[sl] if self.optimizer != "sgd" else None, None, None)
            off += n
        if self.metrics is not None:
            self.metrics.accumulate(logits, None, None, target, loss)
        self._last = logits
        self.steps += 1
        return loss

    @torch.no_grad()
    def evaluate(self, img, target):
        """val phase of train() / test() (MAIN_LPFT:765-799): forward + loss, parameters untouched."""
        self._prepare(img.device)
        _, _, logits = self._forward(img, save=False)
        loss, _ = ops.ce_small(logits, None, None, target.long(), want_grad=False)
        return loss, logits

    def logits(self):
        return self._last


def run_phase(phase, trainer, loader, metrics, num_imgs=None):
    """One phase of MAIN_LPFT's train() (:681-763) or its test() (:765-826) over `loader` (batches of (images, target));
    returns (epoch_loss, epoch_auc, epoch_acc) as the reference defines them, read from the device once."""
    if phase not in ("train", "val", "test"):
        raise MfvError("unknown phase %r" % (phase,))
    metrics.reset()
    if phase == "train":
        if trainer.metrics is not metrics:
            raise MfvError("the trainer must have been built with metrics=<this EpochMetrics>")
        for img, target in loader:
            trainer.step(img, target)
    else:
        for img, target in loader:
            loss, logits = trainer.evaluate(img, target)
            metrics.accumulate(logits, None, None, target.long(), loss)
    return metrics.result(num_imgs)
