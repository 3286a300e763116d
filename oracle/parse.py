"""Oracle (test infrastructure): restatement of ``BiSeNet.predict`` (bise.py:327-418) and the grouping rules."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import nets

MEAN = (0.485, 0.456, 0.406)   # bise.py:187
STD = (0.229, 0.224, 0.225)    # bise.py:188
ATTR_THRESHOLD = 5             # bise.py:185
MASK_THRESHOLD = 10            # bise.py:186


def preprocess(crops_u8_nhwc: np.ndarray) -> torch.Tensor:
    """u8 [F,h,w,3] -> normalised f32 [F,3,512,512]: /255, bilinear (align_corners=False), (x-mean)/std. bise.py:387-392."""
    x = torch.from_numpy(np.ascontiguousarray(crops_u8_nhwc)).permute(0, 3, 1, 2).float()
    x = F.interpolate(x.div(255), (512, 512), mode="bilinear")
    mean = torch.tensor(MEAN).view(1, 3, 1, 1)
    std = torch.tensor(STD).view(1, 3, 1, 1)
    return (x - mean) / std


def labels_from_logits64(logits64: torch.Tensor, in_hw=(512, 512), out_hw=(256, 256)) -> np.ndarray:
    """Reference tail: bilinear(align_corners=True) to ``in_hw`` (bise.py:212) -> nearest to ``out_hw`` -> argmax (bise.py:394)."""
    o = F.interpolate(logits64, in_hw, None, "bilinear", True)
    return F.interpolate(o, out_hw, mode="nearest").argmax(1).numpy().astype(np.uint8)


def labels_from_logits64_sampled(logits64: np.ndarray, in_hw=(512, 512), out_hw=(256, 256)) -> np.ndarray:
    """The fused formulation the CUDA tail implements: evaluate the align_corners bilinear only at the
    pixels the nearest resize picks (src = floor(dst * in/out)), float32 in ATen's operation order."""
    Fh, Fw = logits64.shape[2:]
    def nearest(n_out, n_in):   # ATen nearest: min(floor(dst * (float)in/out), in-1), all in float32
        s = np.float32(n_in) / np.float32(n_out)
        return np.minimum(np.floor(np.arange(n_out, dtype=np.float32) * s).astype(np.int64), n_in - 1)

    ys, xs = nearest(out_hw[0], in_hw[0]), nearest(out_hw[1], in_hw[1])

    def taps(idx, n_in, n_out):
        scale = np.float32((n_in - 1) / (n_out - 1)) if n_out > 1 else np.float32(0)
        r = (scale * idx.astype(np.float32)).astype(np.float32)
        i0 = r.astype(np.int64)
        i1 = np.minimum(i0 + 1, n_in - 1)
        l1 = (r - i0.astype(np.float32)).astype(np.float32)
        return i0, i1, np.float32(1) - l1, l1

    y0, y1, hy0, hy1 = taps(ys, Fh, in_hw[0])
    x0, x1, hx0, hx1 = taps(xs, Fw, in_hw[1])
    L = logits64.astype(np.float32)
    top = hx0 * L[:, :, y0][:, :, :, x0] + hx1 * L[:, :, y0][:, :, :, x1]
    bot = hx0 * L[:, :, y1][:, :, :, x0] + hx1 * L[:, :, y1][:, :, :, x1]
    val = hy0[None, None, :, None] * top + hy1[None, None, :, None] * bot
    return val.argmax(1).astype(np.uint8)


def histogram(labels: np.ndarray) -> np.ndarray:
    """Per-face pixel count of each of the 19 classes, int32 [F,19]."""
    return np.stack([np.bincount(l.ravel(), minlength=19)[:19] for l in labels]).astype(np.int32) \
        if len(labels) else np.zeros((0, 19), np.int32)


def group(labels: np.ndarray, attr_groups, mask_groups):
    """``group_by_attributes`` (bise.py:214-267) + ``group_by_masks`` (bise.py:269-325) + empty-group drop (bise.py:407-416)."""
    hist = histogram(labels)
    attr_out = mask_out = None
    if attr_groups is not None:
        attr_out = {}
        for k, v in attr_groups.items():
            ok = np.ones(len(labels), dtype=bool)
            for a in v:
                cnt = hist[:, abs(a)]
                ok &= (cnt > ATTR_THRESHOLD) if a > 0 else (cnt <= ATTR_THRESHOLD)
            idx = [int(i) for i in np.nonzero(ok)[0]]
            if idx:
                attr_out[k] = idx
    if mask_groups is not None:
        mask_out = {}
        for k, v in mask_groups.items():
            m = np.isin(labels, np.array(v))
            idx = [i for i in range(len(labels)) if m[i].sum() > MASK_THRESHOLD]
            if idx:
                mask_out[k] = (idx, (m[idx] * 255).astype(np.uint8))
    return attr_out, mask_out


def predict(crops_u8_nhwc, sd, attr_groups=None, mask_groups=None, max_batch_size=8):
    """Whole ``BiSeNet.predict``: returns (labels u8[F,h,w], attr_groups, mask_groups)."""
    x = preprocess(crops_u8_nhwc)
    h, w = crops_u8_nhwc.shape[1:3]
    labels = [labels_from_logits64(nets.bisenet_logits64(sub, sd), (512, 512), (h, w))
              for sub in torch.split(x, max_batch_size)]
    labels = np.concatenate(labels)
    return (labels, *group(labels, attr_groups, mask_groups))
