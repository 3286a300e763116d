"""Oracle (test infrastructure): numpy restatement of RetinaFace post-processing.

priors  _layers.py:41-62      decode  retinaface.py:146-212,455-461
filter  retinaface.py:214-304 strategy retinaface.py:306-408
All arithmetic is float32 in the reference's operation order so results are bit-comparable.
"""
from __future__ import annotations

from math import ceil

import numpy as np

F32 = np.float32
STEPS = (8, 16, 32)
MIN_SIZES = ((16, 32), (64, 128), (256, 512))


def priors(height: int, width: int) -> np.ndarray:
    """(A,4) float32 (cx, cy, w, h); python-double maths rounded once to f32 like ``torch.tensor`` (_layers.py:49-62)."""
    out = []
    for step, sizes in zip(STEPS, MIN_SIZES):
        fh, fw = ceil(height / step), ceil(width / step)
        ii, jj = np.meshgrid(np.arange(fh), np.arange(fw), indexing="ij")
        cx = (jj + 0.5) * step / width          # float64, same expression order as the reference
        cy = (ii + 0.5) * step / height
        lvl = np.empty((fh, fw, 2, 4), dtype=np.float64)
        for a, ms in enumerate(sizes):
            lvl[:, :, a, 0], lvl[:, :, a, 1] = cx, cy
            lvl[:, :, a, 2], lvl[:, :, a, 3] = ms / width, ms / height
        out.append(lvl.reshape(-1, 4))
    return np.concatenate(out).astype(F32)


def softmax_face_score(cls_raw: np.ndarray) -> np.ndarray:
    """P(face) = softmax(cls)[..., 1] in float32 (retinaface.py:144,458)."""
    x = cls_raw.astype(F32)
    m = x.max(-1, keepdims=True)
    e = np.exp(x - m)
    return (e[..., 1] / e.sum(-1)).astype(F32)


def decode(box_raw, ldm_raw, pri, height, width):
    """Decoded pixel boxes (…,4) x1y1x2y2 and landmarks (…,10).  retinaface.py:169-176,204-210,459-461."""
    v0, v1 = F32(0.1), F32(0.2)
    loc, pre = box_raw.astype(F32), ldm_raw.astype(F32)
    cxcy = pri[:, :2] + loc[..., :2] * v0 * pri[:, 2:]
    wh = pri[:, 2:] * np.exp(loc[..., 2:] * v1)
    x1y1 = cxcy - wh / F32(2)
    x2y2 = wh + x1y1
    scale = np.array([width, height], dtype=F32)
    boxes = np.concatenate([x1y1 * scale, x2y2 * scale], -1)
    lm = [(pri[:, :2] + pre[..., 2 * k:2 * k + 2] * v0 * pri[:, 2:]) * scale for k in range(5)]
    return boxes.astype(F32), np.concatenate(lm, -1).astype(F32)


def nms_image(scores, boxes, nms_threshold=0.4):
    """Greedy (+1 px) IoU NMS on one image's candidates; returns kept candidate positions, best first.

    Equivalent to the reference's repeated-filter loop (retinaface.py:274-292): a box survives iff
    ``ovr <= thr`` against every previously kept box.  Ties in score are broken lowest-position-first
    (the reference's ``argsort(descending=True)`` is unstable, so test data keeps scores distinct).
    """
    thr = F32(nms_threshold)
    order = np.argsort(-scores.astype(F32), kind="stable")
    area = (boxes[:, 2] - boxes[:, 0] + F32(1)) * (boxes[:, 3] - boxes[:, 1] + F32(1))
    keep = []
    alive = np.ones(len(order), dtype=bool)
    for oi, j in enumerate(order):
        if not alive[oi]:
            continue
        keep.append(int(j))
        rest = order[oi + 1:]
        xx1 = np.maximum(boxes[j, 0], boxes[rest, 0]); yy1 = np.maximum(boxes[j, 1], boxes[rest, 1])
        xx2 = np.minimum(boxes[j, 2], boxes[rest, 2]); yy2 = np.minimum(boxes[j, 3], boxes[rest, 3])
        w = np.maximum(F32(0), xx2 - xx1 + F32(1)); h = np.maximum(F32(0), yy2 - yy1 + F32(1))
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (area[j] + area[rest] - inter)
        alive[oi + 1:] &= (ovr <= thr)
    return keep


def filter_and_select(scores, boxes, ldms, vis_threshold=0.6, nms_threshold=0.4, strategy="all"):
    """Threshold + per-image NMS + strategy.  scores (N,A), boxes (N,A,4), ldms (N,A,10).

    Returns (landmarks f32[F,5,2], indices list[int], anchor_ids list[int], boxes f32[F,4]).
    """
    if strategy not in ("all", "best", "largest"):
        raise ValueError(f"Unsupported startegy: {strategy}")
    out_l, out_i, out_a, out_b = [], [], [], []
    for i in range(scores.shape[0]):
        cand = np.nonzero(scores[i] > F32(vis_threshold))[0]
        if len(cand) == 0:
            continue
        keep = [int(cand[k]) for k in nms_image(scores[i, cand], boxes[i, cand], nms_threshold)]
        if strategy == "best":
            keep = keep[:1]
        elif strategy == "largest":
            b = boxes[i, keep]
            area = (b[:, 2] - b[:, 0] + F32(1)) * (b[:, 3] - b[:, 1] + F32(1))
            keep = [keep[int(np.argmax(area))]]                      # first max wins (torch.argmax)
        out_a += keep
        out_i += [i] * len(keep)
        out_l.append(ldms[i, keep])
        out_b.append(boxes[i, keep])
    if not out_a:
        return np.zeros((0, 5, 2), F32), [], [], np.zeros((0, 4), F32)
    return np.concatenate(out_l).reshape(-1, 5, 2).astype(F32), out_i, out_a, np.concatenate(out_b).astype(F32)


def detect_post(cls_raw, box_raw, ldm_raw, height, width, vis_threshold=0.6, nms_threshold=0.4, strategy="all"):
    """Raw head outputs -> what ``RetinaFace.predict`` returns (+ anchor ids and boxes for diagnostics)."""
    pri = priors(height, width)
    boxes, ldms = decode(box_raw, ldm_raw, pri, height, width)
    return filter_and_select(softmax_face_score(cls_raw), boxes, ldms, vis_threshold, nms_threshold, strategy)
