"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's per-sample input transform for the MF-ViT CA training
loop, operating on the already-resized uint8 image (H x W x 3, the array `cv2.imread` + `Image.fromarray` hands to the
transform at loader.py:122-129):

    image_transform.get_transform_type (image_transform.py:50-84), training=True:
        RandomHorizontalFlip() -> RandomRotation(args.rotate) -> RandomCrop((crop, crop)) -> ToTensor() -> Normalize
    training=False:  CenterCrop((crop, crop)) -> ToTensor() -> Normalize

The transforms themselves live in torchvision / Pillow (third-party, not in /root/reference: torchvision 0.26, Pillow 12.2
in this image).  Restated from their published behaviour:
  * flip        = PIL transpose(FLIP_LEFT_RIGHT): out[y][x] = in[y][W-1-x]
  * rotation    = PIL Image.rotate(angle, NEAREST, expand=False, fill 0): identity copy when angle % 360 == 0, else the
                  inverse affine map about (W/2, H/2) with coefficients rounded to 15 decimals, evaluated at pixel centres
                  in 16.16 fixed point (libImaging Geometry.c `affine_fixed`), source pixels outside the image -> 0
  * crop        = window [top, top+crop) x [left, left+crop)
  * ToTensor    = uint8 -> float32, / 255 (IEEE division);  Normalize = (x - mean[c]) / std[c] in float32
and the random draws consume torch's global RNG in torchvision's order (flip: torch.rand(1); angle:
torch.empty(1).uniform_(-d, d); crop: two torch.randint, skipped when the window equals the image).

Pinned in tests/test_augment.py against the real torchvision transforms (same image, same seed) and against the committed
vectors tests/golden/augment_ref.pt, which oracle/gen_golden.py produced from torchvision itself.
"""
import math

import numpy as np
import torch

# image_transform.py:4-19 (the statistics each image type is normalised with)
STATS = {
    "CheXpert-v1.0-small": ([.5020, .5020, .5020], [float(np.round(np.sqrt(.085585), 4))] * 3),
    "CheXpert_Enh": ([.6086, .5204, .3384], [.134909, .088268, .035044]),
    "data": ([0.5045, 0.5045, 0.5045], [0.2462, 0.2462, 0.2462]),
    "Train_Mix": ([0.2243, 0.5507, 0.6865], [0.1026, 0.2995, 0.3300]),
}


def _fix(v):
    return int(math.floor(v * 65536.0 + 0.5))


def rotation_fixed(angle_deg, w, h):
    """16.16 fixed-point inverse map (a0..a5) PIL uses for rotate(angle_deg); None = identity fast path."""
    angle = angle_deg % 360.0
    if angle == 0:
        return None
    if angle == 180 or (angle in (90, 270) and w == h):
        raise ValueError("exact quarter turns take PIL's transpose path; not part of RandomRotation(+-1)")
    cx, cy = w / 2, h / 2
    a = -math.radians(angle)
    m = [round(math.cos(a), 15), round(math.sin(a), 15), 0.0, round(-math.sin(a), 15), round(math.cos(a), 15), 0.0]
    m[2] = m[0] * -cx + m[1] * -cy + m[2]
    m[5] = m[3] * -cx + m[4] * -cy + m[5]
    m[2] += cx
    m[5] += cy
    return (_fix(m[0]), _fix(m[1]), _fix(m[2] + m[0] * 0.5 + m[1] * 0.5),
            _fix(m[3]), _fix(m[4]), _fix(m[5] + m[3] * 0.5 + m[4] * 0.5))


def apply_u8(img, flip, fixed, top, left, crop):
    """img uint8 [H][W][3] -> uint8 [crop][crop][3]: flip, rotate (fixed = rotation_fixed(...) or None), crop."""
    h, w, _ = img.shape
    if flip:
        img = img[:, ::-1]
    if fixed is not None:
        a0, a1, a2, a3, a4, a5 = fixed
        ys, xs = np.meshgrid(np.arange(h, dtype=np.int64), np.arange(w, dtype=np.int64), indexing="ij")
        xin = (a2 + ys * a1 + xs * a0) >> 16
        yin = (a5 + ys * a4 + xs * a3) >> 16
        ok = (xin >= 0) & (xin < w) & (yin >= 0) & (yin < h)
        out = np.zeros_like(img)
        out[ok] = img[yin[ok], xin[ok]]
        img = out
    return np.ascontiguousarray(img[top:top + crop, left:left + crop])


def to_tensor_normalize(u8_hwc, mean, std):
    """ToTensor + Normalize, float32 [3][H][W]."""
    t = torch.from_numpy(np.ascontiguousarray(u8_hwc)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    mean = torch.as_tensor(mean, dtype=torch.float32).view(-1, 1, 1)
    std = torch.as_tensor(std, dtype=torch.float32).view(-1, 1, 1)
    return t.sub_(mean).div_(std)


def draw_train(h, w, crop, degrees):
    """One sample's random parameters from torch's global RNG, in the order the torchvision Compose consumes it."""
    flip = bool(torch.rand(1) < 0.5)
    angle = float(torch.empty(1).uniform_(-float(degrees), float(degrees)).item())
    if crop == 0 or (h == crop and w == crop):
        top = left = 0
    else:
        top = int(torch.randint(0, h - crop + 1, size=(1,)).item())
        left = int(torch.randint(0, w - crop + 1, size=(1,)).item())
    return flip, angle, top, left


def draw_train_batch(n, h, w, crop, degrees):
    """The loader's batched draw order (mfvit.data.draw_train_params_batch): same distributions, drawn per batch."""
    flips = [bool(v) for v in (torch.rand(n) < 0.5)]
    angles = [float(v) for v in torch.empty(n).uniform_(-float(degrees), float(degrees))]
    if crop == 0 or (h == crop and w == crop):
        tops = lefts = [0] * n
    else:
        tops = [int(v) for v in torch.randint(0, h - crop + 1, size=(n,))]
        lefts = [int(v) for v in torch.randint(0, w - crop + 1, size=(n,))]
    return list(zip(flips, angles, tops, lefts))


def center_crop_offsets(h, w, crop):
    return int(round((h - crop) / 2.0)), int(round((w - crop) / 2.0))


def transform_train(img, crop, degrees, mean, std):
    h, w, _ = img.shape
    flip, angle, top, left = draw_train(h, w, crop, degrees)
    c = crop if crop else h
    return to_tensor_normalize(apply_u8(img, flip, rotation_fixed(angle, w, h), top, left, c), mean, std)


def transform_eval(img, crop, mean, std):
    h, w, _ = img.shape
    top, left = center_crop_offsets(h, w, crop) if crop else (0, 0)
    return to_tensor_normalize(apply_u8(img, False, None, top, left, crop if crop else h), mean, std)


def epoch_metrics(vals, gts):
    """MAIN_CA:895-909: argmax accuracy and the mean one-vs-rest ROC AUC of the raw summed logits.  AUC restated as the
    rank statistic (ties count half), which is what sklearn's roc_curve + auc integrate to."""
    vals = np.asarray(vals, dtype=np.float64)
    gts = np.asarray(gts)
    acc = float(np.sum(vals.argmax(1) == gts)) / len(gts)
    aucs = []
    for c in range(vals.shape[1]):
        pos, neg = vals[gts == c, c], vals[gts != c, c]
        if len(pos) == 0 or len(neg) == 0:
            aucs.append(float("nan"))
            continue
        gt = (pos[:, None] > neg[None, :]).sum() + 0.5 * (pos[:, None] == neg[None, :]).sum()
        aucs.append(float(gt) / (len(pos) * len(neg)))
    return acc, float(np.mean(aucs))
