"""Oracle (test infrastructure): restatement of ``RRDBNet.predict`` (rrdb.py:83-146)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import nets


def enhance_image(img_f32_chw: torch.Tensor, sd, nb=23) -> torch.Tensor:
    """One image: net(img/255) -> bicubic x0.25 -> clamp(0,1)*255 -> round.  rrdb.py:142-144."""
    x4 = nets.rrdbnet_forward(img_f32_chw.unsqueeze(0).div(255), sd, nb=nb)
    x1 = F.interpolate(x4, None, 0.25, "bicubic")
    return x1.clamp(0, 1).mul(255).round()[0]


def bicubic_quarter_stencil(x4: torch.Tensor) -> torch.Tensor:
    """The x0.25 bicubic resize written as the fixed separable 4-tap stencil [-3/32,19/32,19/32,-3/32], stride 4
    (what the CUDA ``conv_last`` epilogue implements; equals ``F.interpolate(x4, None, 0.25, 'bicubic')`` to ~2e-7)."""
    w = torch.tensor([-3 / 32, 19 / 32, 19 / 32, -3 / 32], dtype=x4.dtype)
    k = (w[:, None] * w[None, :]).expand(x4.shape[1], 1, 4, 4).contiguous()
    return F.conv2d(x4, k, stride=4, groups=x4.shape[1])


def should_enhance(landmarks, indices, i, h0, w0, min_face_factor) -> bool:
    """Per-image gate, rrdb.py:125-140 (face area normalised by image 0's size — reference quirk kept)."""
    if landmarks is None or indices is None:
        return True
    sel = landmarks[[idx == i for idx in indices]]
    if len(sel) == 0:
        return False
    w, h = (sel[:, 4] - sel[:, 0]).T
    return bool((w * h / (h0 * w0)).mean() <= min_face_factor)


def predict(images_f32_nchw: torch.Tensor, sd, landmarks=None, indices=None, min_face_factor=0.001, nb=23):
    """``RRDBNet.predict``: enhances gated images in place and returns the batch."""
    h0, w0 = images_f32_nchw[0].shape[1:]
    for i in range(len(images_f32_nchw)):
        if should_enhance(landmarks, indices, i, h0, w0, min_face_factor):
            images_f32_nchw[i] = enhance_image(images_f32_nchw[i], sd, nb)
    return images_f32_nchw
