"""Host-side helpers on either side of the hot path (file I/O, batching, landmark files).

These mirror ``face_crop_plus/utils.py`` in behaviour (same function names and return conventions) but are plain
host code for decoding/encoding and landmark files; ``as_batch`` runs on the GPU through ``fcp_as_batch``.
"""
from __future__ import annotations

import json
import os
import warnings

import numpy as np

from .landmarks import STANDARD_LANDMARKS_5  # noqa: F401  (re-exported like utils.py:13)

# which points of an N-point annotation are averaged into (left eye, right eye, nose, left mouth, right mouth)
_SLICES_5 = {
    5: [(0, 1), (1, 2), (2, 3), (3, 4), (4, 5)],
    12: [(10, 11), (11, 12), (2, 3), (3, 4), (4, 5)],
    17: [(2, 5), (7, 10), (10, 11), (13, 14), (16, 17)],
    21: [(6, 9), (9, 12), (14, 15), (17, 18), (19, 20)],
    29: [(4, 9), (13, 18), (19, 20), (22, 23), (27, 28)],
    49: [(19, 25), (25, 31), (13, 14), (31, 32), (37, 38)],
    68: [(36, 42), (42, 48), (30, 31), (48, 49), (54, 55)],
    98: [(60, 68), (68, 76), (54, 55), (76, 77), (82, 83)],
    106: [(66, 75), (75, 84), (54, 55), (85, 86), (91, 92)],
}


def get_landmark_slices_5(num_landmarks: int) -> list[slice]:
    """utils.py:90-132."""
    if num_landmarks not in _SLICES_5:
        raise ValueError(f"Invalid number of landmarks: {num_landmarks}")
    return [slice(a, b) for a, b in _SLICES_5[num_landmarks]]


def get_ldm_slices(num_tgt_landmarks: int, num_src_landmarks: int) -> list[slice]:
    """utils.py:134-168: only 5 target landmarks are supported."""
    if num_tgt_landmarks != 5:
        raise ValueError(f"The number of target landmarks is not supported: {num_tgt_landmarks}")
    return get_landmark_slices_5(num_src_landmarks)


def parse_landmarks_file(file_path: str, **kwargs) -> tuple[np.ndarray, np.ndarray]:
    """(landmarks f32 [F,K,2], file names [F]) from .json / .csv / whitespace-separated text (utils.py:21-88)."""
    if file_path.endswith(".json"):
        with open(file_path) as f:
            data = json.load(f)
        names = np.array(list(data.keys()))
        lms = np.array(list(data.values()), dtype=np.float32)
    else:
        if file_path.endswith(".csv"):
            kwargs.setdefault("delimiter", ",")
            kwargs.setdefault("skip_header", 1)          # utils.py:70-71: a .csv carries a header row
        names = np.genfromtxt(file_path, usecols=0, dtype=str, **kwargs)
        lms = np.genfromtxt(file_path, dtype=np.float32, **kwargs)[:, 1:]
    return lms.reshape(len(lms), -1, 2), names


def read_images(file_names, input_dir: str):
    """RGB uint8 arrays of the readable files + the names that survived (utils.py:228-271)."""
    import cv2
    images, kept = [], []
    for name in file_names:
        path = os.path.join(input_dir, name)
        bgr = cv2.imread(path)
        if bgr is None:
            warnings.warn(f"Could not read the image {path}")
            continue
        images.append(cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB))
        kept.append(name)
    return images, np.array(file_names)[[i for i, n in enumerate(file_names) if n in set(kept)]] if kept else np.array([], dtype=str)


def batch_plan(h: int, w: int, size) -> tuple[int, int, float, list[int]]:
    """Geometry of utils.py:317-331 for one image: (new_w, new_h, unscale, [top, bottom, left, right])."""
    size = (size, size) if isinstance(size, int) else tuple(size)
    rw, rh = size[0] / w, size[1] / h
    if rw < rh:
        new_w, new_h, unscale = size[0], int(h * rw), rw
        return new_w, new_h, unscale, [(size[1] - new_h) // 2, (size[1] - new_h + 1) // 2, 0, 0]
    new_w, new_h, unscale = int(w * rh), size[1], rh
    return new_w, new_h, unscale, [0, 0, (size[0] - new_w) // 2, (size[0] - new_w + 1) // 2]


def as_batch(images, size=512, padding_mode: str = "constant", ctx=None):
    """Aspect-preserving resize to fit ``size`` (w, h) + centred padding; returns (batch u8 [N,H,W,3], unscales, paddings
    [N,4] = top,bottom,left,right) exactly like utils.py:273-342 (INTER_AREA when shrinking, INTER_CUBIC otherwise).

    Runs on the GPU (``fcp_as_batch``: OpenCV's INTER_AREA / INTER_CUBIC / copyMakeBorder arithmetic restated in
    ``csrc/ingest.cu``); ``ctx`` is the :class:`_abi.Context` to use (default: the context of cuda:0)."""
    if ctx is None:
        from .models import get_context
        ctx = get_context("cuda:0")
    return ctx.as_batch(images, size, padding_mode)
