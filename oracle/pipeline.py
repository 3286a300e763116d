"""Oracle (test infrastructure): the hot path of ``Cropper.process_batch`` (cropper.py:815-847), in memory.

detect (RetinaFace.predict) -> landmark un-pad -> [enhance] -> crop_align -> BiSeNet.predict, for a batch that is
already resized/padded (``as_batch`` is the identity for the 1024x1024 configs).
"""
from __future__ import annotations

import numpy as np
import torch

from face_crop_plus_b200.landmarks import landmarks_target

from . import align, detpost, enhance, nets, parse


def detect(images_u8_nhwc: np.ndarray, det_sd, vis_threshold=0.6, nms_threshold=0.4, strategy="largest"):
    """``RetinaFace.predict`` (retinaface.py:410-470) on a uint8 NHWC batch (as_tensor: utils.py:224)."""
    n, h, w, _ = images_u8_nhwc.shape
    x = torch.from_numpy(np.ascontiguousarray(images_u8_nhwc)).permute(0, 3, 1, 2).float()
    cls, box, ldm = nets.retinaface_heads_raw(nets.retinaface_preprocess(x), det_sd)
    return detpost.detect_post(cls.numpy(), box.numpy(), ldm.numpy(), h, w, vis_threshold, nms_threshold, strategy)


def process_batch(images_u8_nhwc, det_sd, par_sd=None, enh_sd=None, paddings=None, *, output_size=(256, 256),
                  face_factor=0.65, strategy="largest", padding="constant", allow_skew=False, det_threshold=0.6,
                  enh_threshold=None, attr_groups=None, mask_groups=None, batch_size=8, rrdb_blocks=23):
    """Returns dict(landmarks, indices, crops, matrices, labels, attr_groups, mask_groups)."""
    torch.set_grad_enabled(False)
    images = np.ascontiguousarray(images_u8_nhwc)
    landmarks, indices, anchors, boxes = detect(images, det_sd, det_threshold, 0.4, strategy)
    out = dict(landmarks=landmarks, indices=indices, anchors=anchors, boxes=boxes, crops=np.array([]),
               matrices=None, labels=None, attr_groups=None, mask_groups=None)
    if paddings is not None and len(indices):
        landmarks = landmarks - np.asarray(paddings)[indices][:, None, [2, 0]]       # cropper.py:822
        out["landmarks"] = landmarks
    if len(landmarks) == 0:
        return out                                                                    # cropper.py:824-826
    if enh_sd is not None and enh_threshold is not None:
        x = torch.from_numpy(images).permute(0, 3, 1, 2).float()                      # cropper.py:835-839
        x = enhance.predict(x, enh_sd, landmarks, indices, enh_threshold, rrdb_blocks)
        images = x.permute(0, 2, 3, 1).numpy().astype(np.uint8)
    tgt = landmarks_target(output_size, face_factor)
    crops, mats, valid = align.crop_align(images, paddings, indices, landmarks, tgt, output_size, padding, allow_skew)
    out.update(crops=crops, matrices=mats, valid=valid)
    if par_sd is not None and len(crops):
        labels, ag, mg = parse.predict(crops, par_sd, attr_groups, mask_groups, batch_size)
        out.update(labels=labels, attr_groups=ag, mask_groups=mg)
    return out
