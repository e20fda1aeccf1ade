"""Paired GPU-side input pipeline and on-device epoch metrics for the MF-ViT CA loop.

The reference builds two independent `DataLoader(..., shuffle=True)` objects, one per image type (MAIN_CA:573-578,
664-669), so the CXR and the enhanced image of a step are not guaranteed to belong to the same patient, and every
sample is flipped / rotated / cropped / normalised by PIL workers into a float32 tensor that is then copied to the GPU.
Here one store holds both uint8 images of every sample (decoded and resized once), one permutation indexes both, the
uint8 batch (a quarter of the float32 bytes) goes over PCIe from pinned memory on a copy stream, and a single kernel
(`mfv_augment_u8`) applies image_transform.py:50-84's flip / rotation / crop / ToTensor / Normalize on the device,
bit-identical to the torchvision result.  The random draws follow torchvision's order per sample, from a seeded
generator, so a run is reproducible and the CPU restatement in oracle/augment_ref.py can be replayed against it.

    store  = PairedU8Store(cxr_u8, enh_u8, labels)                    # uint8 [N][H][W][3] x 2, int64 [N]
    loader = PairedDeviceLoader(store, batch_size=32, crop=224, degrees=1, training=True, device="cuda")
    for img_cxr, img_enh, target in loader:                           # float32 [B][3][crop][crop] on the device
        loss = trainer.step(img_cxr, img_enh, target)
"""
import math
import queue
import threading

import numpy as np
import torch

from . import ops
from ._lib import MfvError

# image_transform.py:4-19
STATS = {
    "CheXpert-v1.0-small": ([.5020, .5020, .5020], [float(np.round(np.sqrt(.085585), 4))] * 3),
    "CheXpert_Enh": ([.6086, .5204, .3384], [.134909, .088268, .035044]),
    "data": ([0.5045, 0.5045, 0.5045], [0.2462, 0.2462, 0.2462]),
    "Train_Mix": ([0.2243, 0.5507, 0.6865], [0.1026, 0.2995, 0.3300]),
}
N_PARAMS = 12  # MFV_AUG_PARAMS


def rotation_coefficients(angle_deg, w, h):
    """Pillow's inverse affine map for Image.rotate(angle, NEAREST) as 16.16 fixed-point integers; None = no rotation."""
    angle = angle_deg % 360.0
    if angle == 0:
        return None
    if angle == 180 or (angle in (90, 270) and w == h):
        raise MfvError("exact quarter turns are not produced by RandomRotation and are not supported")
    cx, cy = w / 2, h / 2
    rad = -math.radians(angle)
    cos, sin = round(math.cos(rad), 15), round(math.sin(rad), 15)
    m = [cos, sin, 0.0, -sin, cos, 0.0]
    m[2] = m[0] * -cx + m[1] * -cy + cx
    m[5] = m[3] * -cx + m[4] * -cy + cy
    fix = lambda v: int(math.floor(v * 65536.0 + 0.5))  # noqa: E731
    coef = (fix(m[0]), fix(m[1]), fix(m[2] + m[0] * 0.5 + m[1] * 0.5),
            fix(m[3]), fix(m[4]), fix(m[5] + m[3] * 0.5 + m[4] * 0.5))
    for x, y in ((0, 0), (w, 0), (0, h), (w, h)):  # the range Pillow itself requires of the fixed-point path
        if abs(x * m[0] + y * m[1] + m[2]) >= 32768.0 or abs(x * m[3] + y * m[4] + m[5]) >= 32768.0:
            raise MfvError("image too large for the 16.16 fixed-point rotation")
    return coef


def pack_params(samples, w, h, out=None):
    """samples: iterable of (flip, angle_deg, top, left) -> int32 [n][12] rows for mfv_augment_u8."""
    samples = list(samples)
    if out is None:
        out = torch.zeros(len(samples), N_PARAMS, dtype=torch.int32)
    rows = out.numpy()
    rows[:len(samples)] = 0
    for i, (flip, angle, top, left) in enumerate(samples):
        coef = rotation_coefficients(angle, w, h)
        rows[i, 0] = 1 if flip else 0
        if coef is not None:
            rows[i, 1] = 1
            rows[i, 2:8] = coef
        rows[i, 8], rows[i, 9] = top, left
    return out


def draw_train_params(n, h, w, crop, degrees, generator=None):
    """n samples of (flip, angle, top, left), consuming `generator` (default: the global one) exactly as the Compose of
    RandomHorizontalFlip, RandomRotation(degrees), RandomCrop((crop, crop)) does for one image after another."""
    deg = float(degrees)
    out = []
    for _ in range(n):
        flip = bool(torch.rand(1, generator=generator) < 0.5)
        angle = float(torch.empty(1).uniform_(-deg, deg, generator=generator).item())
        if crop == 0 or (h == crop and w == crop):
            top = left = 0
        else:
            top = int(torch.randint(0, h - crop + 1, size=(1,), generator=generator).item())
            left = int(torch.randint(0, w - crop + 1, size=(1,), generator=generator).item())
        out.append((flip, angle, top, left))
    return out


def draw_train_params_batch(n, h, w, crop, degrees, generator=None):
    """The same distributions drawn a batch at a time (flips, then angles, then window rows, then window columns): four
    generator calls per batch instead of up to four per sample, which is what keeps the loader's worker thread off the
    GIL.  (A multi-worker DataLoader has no reproducible per-sample stream to follow in the first place.)"""
    deg = float(degrees)
    flips = (torch.rand(n, generator=generator) < 0.5).tolist()
    angles = torch.empty(n, dtype=torch.float32).uniform_(-deg, deg, generator=generator).tolist()
    if crop == 0 or (h == crop and w == crop):
        tops = lefts = [0] * n
    else:
        tops = torch.randint(0, h - crop + 1, size=(n,), generator=generator).tolist()
        lefts = torch.randint(0, w - crop + 1, size=(n,), generator=generator).tolist()
    return list(zip(flips, angles, tops, lefts))


def eval_params(n, h, w, crop):
    top, left = (int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))) if crop else (0, 0)  # CenterCrop
    return [(False, 0.0, top, left)] * n


def shard_indices(n, epoch, seed, shuffle, rank=0, world_size=1, drop_last=False):
    """One permutation per epoch shared by both image types; contiguous-stride sharding with wrap-around padding, i.e.
    torch.utils.data.DistributedSampler's rule, so every rank sees the same number of samples."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        idx = torch.randperm(n, generator=g)
    else:
        idx = torch.arange(n)
    if world_size > 1:
        if drop_last and n % world_size:
            idx = idx[:n - n % world_size]
        else:
            pad = (-len(idx)) % world_size
            if pad:
                idx = torch.cat([idx, idx[:pad]])
        idx = idx[rank::world_size]
    return idx


class PairedU8Store:
    """Both uint8 images and the label of every sample, index-aligned, in pinned host memory."""

    def __init__(self, cxr_u8, enh_u8, labels, pin=True):
        cxr_u8, enh_u8 = torch.as_tensor(cxr_u8), torch.as_tensor(enh_u8)
        labels = torch.as_tensor(labels).long()
        for t in (cxr_u8, enh_u8):
            if t.dtype != torch.uint8 or t.dim() != 4 or t.shape[-1] != 3:
                raise MfvError("PairedU8Store wants uint8 [N][H][W][3] arrays")
        if not (len(cxr_u8) == len(enh_u8) == len(labels)):
            raise MfvError("CXR / enhanced / label counts differ: %d / %d / %d" % (len(cxr_u8), len(enh_u8), len(labels)))
        if cxr_u8.shape != enh_u8.shape:
            raise MfvError("both image types must be resized to the same H x W")
        pin = pin and torch.cuda.is_available()
        self.cxr = cxr_u8.contiguous().pin_memory() if pin else cxr_u8.contiguous()
        self.enh = enh_u8.contiguous().pin_memory() if pin else enh_u8.contiguous()
        self.labels = labels

    def __len__(self):
        return len(self.labels)

    @classmethod
    def from_csv(cls, folder_cxr, folder_enh, img_csv, img_size=224, maintain_ratio=False):
        """Decode + resize once.  Same list format as loader.py:Dataset_covid (space separated; fields[1] = root,
        fields[2] = file, fields[-2] = label) and the same decode (cv2.imread -> PIL) and Resize as
        image_transform.py:52-55.  The enhanced image of a sample sits under `folder_enh` with the same file name."""
        import os

        import cv2
        from PIL import Image
        cxr, enh, labels = [], [], []
        if maintain_ratio:
            raise MfvError("the uint8 store holds one H x W for every image: resize with maintain_ratio=False")
        size = (img_size, img_size)
        import torchvision.transforms as T
        resize = T.Resize(size)
        with open(img_csv) as f:
            for line in f:
                fields = line.strip("\n").split(" ")
                pair = []
                for folder in (folder_cxr, folder_enh):
                    img = cv2.imread(os.path.join(fields[1], folder, fields[2]))
                    if img is None:
                        raise MfvError("cannot read %s" % os.path.join(fields[1], folder, fields[2]))
                    pair.append(np.asarray(resize(Image.fromarray(img))))
                cxr.append(pair[0])
                enh.append(pair[1])
                labels.append(int(float(fields[-2])))
        return cls(np.stack(cxr), np.stack(enh), np.asarray(labels))


def _gather_rows(src, idx, dst):
    """dst[j] = src[idx[j]] for uint8 image stacks: whole rows moved as 8-byte words (index_select on uint8 goes byte by
    byte and is ~250x slower)."""
    row = src[0].numel()
    if row % 8 == 0:
        torch.index_select(src.view(len(src), row).view(torch.int64), 0, idx, out=dst.view(len(dst), row).view(torch.int64))
    else:
        np.take(src.numpy(), idx.numpy(), axis=0, out=dst.numpy())


class _Slot:
    def __init__(self, B, H, W, crop, device, pinned):
        mk = lambda *s, dt: torch.empty(*s, dtype=dt, pin_memory=pinned)  # noqa: E731
        self.h_cxr, self.h_enh = mk(B, H, W, 3, dt=torch.uint8), mk(B, H, W, 3, dt=torch.uint8)
        self.h_par, self.h_lab = mk(2, B, N_PARAMS, dt=torch.int32), mk(B, dt=torch.int64)
        self.d_cxr, self.d_enh = (torch.empty(B, H, W, 3, dtype=torch.uint8, device=device) for _ in range(2))
        self.d_par = torch.empty(2, B, N_PARAMS, dtype=torch.int32, device=device)
        self.d_lab = torch.empty(B, dtype=torch.int64, device=device)
        self.out = [torch.empty(B, 3, crop, crop, dtype=torch.float32, device=device) for _ in range(2)]
        self.copied = torch.cuda.Event()    # H2D of this slot finished: the pinned staging may be refilled
        self.ready = torch.cuda.Event()     # transform finished: the float32 batch may be consumed
        self.consumed = torch.cuda.Event()  # the consumer's work on this slot's outputs is done
        self.n = 0
        self.used = False


class PairedDeviceLoader:
    """Iterates index-aligned (img_cxr, img_enh, target) device batches; the next batches are gathered, copied and
    transformed on a copy stream while the caller works on the current one.  The yielded tensors are reused three
    batches later (after the work the caller enqueued on them has finished)."""

    def __init__(self, store, batch_size, crop=224, degrees=0, training=True, img_types=("data", "Train_Mix"),
                 device="cuda", shuffle=True, seed=0, rank=0, world_size=1, drop_last=False):
        self.store, self.B, self.crop, self.degrees, self.training = store, batch_size, crop, degrees, training
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise MfvError("PairedDeviceLoader runs its transforms on the GPU; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.shuffle, self.seed, self.rank, self.world, self.drop_last = shuffle, seed, rank, world_size, drop_last
        self.epoch = 0
        _, self.H, self.W, _ = store.cxr.shape
        if crop % 4 or crop > self.H or crop > self.W:
            raise MfvError("crop must be a multiple of 4 and fit the stored images")
        self.stats = []
        for t in img_types:
            mean, std = STATS[t] if isinstance(t, str) else t
            self.stats.append((torch.tensor(mean, dtype=torch.float32, device=self.device),
                               torch.tensor(std, dtype=torch.float32, device=self.device)))
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.slots = [_Slot(batch_size, self.H, self.W, crop, self.device, True) for _ in range(3)]
        self.h2d_bytes_per_batch = 2 * batch_size * self.H * self.W * 3 + 2 * batch_size * N_PARAMS * 4 + 8 * batch_size

    def set_epoch(self, epoch):
        self.epoch = epoch

    def __len__(self):
        n = len(shard_indices(len(self.store), 0, 0, False, self.rank, self.world, self.drop_last))
        return n // self.B if self.drop_last else (n + self.B - 1) // self.B

    def _draw(self, n, gen):
        """int32 [2][n][12]: each image type draws its own flip / angle / window, as two transform calls would."""
        par = torch.zeros(2, n, N_PARAMS, dtype=torch.int32)
        for t in range(2):
            samples = (draw_train_params_batch(n, self.H, self.W, self.crop, self.degrees, gen) if self.training
                       else eval_params(n, self.H, self.W, self.crop))
            pack_params(samples, self.W, self.H, out=par[t])
        return par

    def _stage(self, slot, idx, par):
        """Worker thread: host gather into pinned staging, H2D and the device transform, all on the copy stream."""
        n = len(idx)
        if slot.used:
            slot.copied.synchronize()  # the previous H2D out of this staging buffer is done
        _gather_rows(self.store.cxr, idx, slot.h_cxr[:n])
        _gather_rows(self.store.enh, idx, slot.h_enh[:n])
        slot.h_lab[:n] = self.store.labels[idx]
        slot.h_par[:, :n] = par
        with torch.cuda.stream(self.copy_stream):
            if slot.used:
                self.copy_stream.wait_event(slot.consumed)  # the consumer's kernels are done with this slot's outputs
            slot.d_cxr[:n].copy_(slot.h_cxr[:n], non_blocking=True)
            slot.d_enh[:n].copy_(slot.h_enh[:n], non_blocking=True)
            slot.d_par.copy_(slot.h_par, non_blocking=True)
            slot.d_lab[:n].copy_(slot.h_lab[:n], non_blocking=True)
            slot.copied.record(self.copy_stream)
            for t, (src, (mean, std)) in enumerate(zip((slot.d_cxr, slot.d_enh), self.stats)):
                ops.augment_u8(src[:n], slot.d_par[t, :n], mean, std, self.crop, out=slot.out[t][:n])
            slot.ready.record(self.copy_stream)
        slot.n, slot.used = n, True

    def __iter__(self):
        idx = shard_indices(len(self.store), self.epoch, self.seed, self.shuffle, self.rank, self.world, self.drop_last)
        gen = torch.Generator()
        gen.manual_seed(self.seed * 1000003 + self.epoch * 1009 + self.rank)
        batches = [idx[i:i + self.B] for i in range(0, len(idx), self.B)]
        if self.drop_last and batches and len(batches[-1]) < self.B:
            batches.pop()
        # A worker thread gathers, draws, copies and launches the transform ahead of the consumer (index_select and the
        # CUDA enqueues release the GIL): when the consumer asks for a batch it only waits on an event, so the host and
        # copy-engine side of batch i+1 overlaps the training step of batch i even if the loss is read every step.
        staged = queue.Queue()
        free = [threading.Semaphore(1) for _ in self.slots]
        stop = threading.Event()

        def worker():
            try:
                torch.cuda.set_device(self.device)
                # The interpreter-bound part (random draws, fixed-point coefficients) of batch i+1 is done right after
                # batch i is staged, i.e. while the consumer is busy with an earlier batch - not in the window after a
                # slot is handed back, where it would compete with the consumer's next launch for the GIL.
                par = self._draw(len(batches[0]), gen) if batches else None
                for i, b in enumerate(batches):
                    k = i % len(self.slots)
                    free[k].acquire()
                    if stop.is_set():
                        return
                    self._stage(self.slots[k], b, par)
                    staged.put(k)
                    if i + 1 < len(batches):
                        par = self._draw(len(batches[i + 1]), gen)
                staged.put(None)
            except BaseException as e:  # noqa: BLE001 - re-raised in the consumer
                staged.put(e)

        th = threading.Thread(target=worker, daemon=True)
        th.start()
        try:
            while True:
                k = staged.get()
                if k is None:
                    break
                if isinstance(k, BaseException):
                    raise k
                slot = self.slots[k]
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(slot.ready)
                yield slot.out[0][:slot.n], slot.out[1][:slot.n], slot.d_lab[:slot.n]
                # back from the consumer: whatever it enqueued on its stream has to finish before the slot is rewritten
                slot.consumed.record(torch.cuda.current_stream(self.device))
                free[k].release()
        finally:
            stop.set()
            for f in free:
                f.release()
            th.join()  # an abandoned pass must not keep staging into the slots the next pass will use
            cur = torch.cuda.current_stream(self.device)
            for slot in self.slots:  # an abandoned pass: later passes must still order after the consumer's work
                slot.consumed.record(cur)


def roc_auc_ovr_mean(vals, gts, num_classes):
    """Mean over classes of the one-vs-rest ROC AUC of the raw scores (MAIN_CA:897-903: label_binarize + roc_curve + auc),
    computed from mid-ranks: AUC_c = (sum of positive ranks - P(P+1)/2) / (P N)."""
    vals = np.asarray(vals, dtype=np.float64)
    gts = np.asarray(gts)
    aucs = []
    for c in range(num_classes):
        s = vals[:, c]
        pos = gts == c
        P, N = int(pos.sum()), int((~pos).sum())
        if P == 0 or N == 0:
            aucs.append(float("nan"))
            continue
        order = np.argsort(s, kind="mergesort")
        ranks = np.empty(len(s), dtype=np.float64)
        sorted_s = s[order]
        # mid-ranks for ties
        bounds = np.flatnonzero(np.concatenate(([True], sorted_s[1:] != sorted_s[:-1], [True])))
        for lo, hi in zip(bounds[:-1], bounds[1:]):
            ranks[order[lo:hi]] = 0.5 * (lo + hi - 1) + 1.0
        aucs.append((ranks[pos].sum() - P * (P + 1) / 2.0) / (P * N))
    return float(np.mean(aucs))


class EpochMetrics:
    """Loss / accuracy / AUC inputs accumulated on the device by `mfv_epoch_metrics`; one D2H per epoch (result())."""

    def __init__(self, capacity, num_classes=3, device="cuda"):
        self.capacity, self.NC = int(capacity), num_classes
        dev = torch.device(device)
        self.loss_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        self.counters = torch.zeros(3, dtype=torch.int64, device=dev)
        self.vals = torch.zeros(self.capacity, num_classes, dtype=torch.float32, device=dev)
        self.preds = torch.zeros(self.capacity, dtype=torch.int32, device=dev)
        self.gts = torch.zeros(self.capacity, dtype=torch.int32, device=dev)

    def accumulate(self, fused, x_cxr, x_enh, target, loss):
        ops.epoch_metrics_(fused, x_cxr, x_enh, target, loss, self.loss_sum, self.counters, self.vals, self.preds,
                           self.gts)

    def reset(self):
        self.loss_sum.zero_()
        self.counters.zero_()

    def result(self, num_imgs=None):
        """(epoch_loss, epoch_auc, epoch_acc) as MAIN_CA:905-907 computes them (divisor = num_imgs, default: rows seen)."""
        seen, hits, dropped = (int(v) for v in self.counters.cpu())
        if dropped:
            raise MfvError("EpochMetrics capacity %d too small: %d rows dropped" % (self.capacity, dropped))
        n = num_imgs if num_imgs else max(seen, 1)
        vals, gts = self.vals[:seen].cpu().numpy(), self.gts[:seen].cpu().numpy()
        auc = roc_auc_ovr_mean(vals, gts, self.NC) if seen else float("nan")
        return float(self.loss_sum.item()) / n, auc, hits / n
