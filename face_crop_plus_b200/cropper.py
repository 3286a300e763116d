"""Drop-in ``Cropper`` whose per-batch hot path runs on the B200 through libfcpb200.so.

Keeps the reference's public surface (cropper.py:139-156 constructor, ``process_dir`` :852, ``process_batch`` :748,
``crop_align`` :441, ``save_group(s)`` :554/:611, attributes ``det_model/enh_model/par_model/landmarks_target``).
What differs is where the work happens: the detect -> un-pad -> align -> parse sequence of ``process_batch``
(cropper.py:815-847) is ONE call into the CUDA library with a uint8 batch — no float32 NCHW round trips, no
per-face Python loop.
"""
from __future__ import annotations

import os
from collections import defaultdict
from functools import partial
from multiprocessing.pool import ThreadPool

import numpy as np
import torch

from . import _abi
from .landmarks import landmarks_target
from .models import BiSeNet, RetinaFace, RRDBNet, _lock, bind_stream, get_context
from .utils import as_batch, get_ldm_slices, parse_landmarks_file, read_images


class Cropper:
    def __init__(self, output_size=256, output_format=None, resize_size=1024, face_factor=0.65, strategy="largest",
                 padding="constant", allow_skew=False, landmarks=None, attr_groups=None, mask_groups=None,
                 det_threshold=0.6, enh_threshold=None, batch_size=8, num_processes=1, device="cuda:0",
                 state_dicts: dict | None = None):
        as_pair = lambda v: (v, v) if isinstance(v, int) else ((v[0], v[0]) if len(v) == 1 else tuple(v))
        self.output_size, self.resize_size = as_pair(output_size), as_pair(resize_size)
        self.output_format, self.face_factor, self.strategy = output_format, face_factor, strategy
        self.padding, self.allow_skew = padding, allow_skew
        self.landmarks = parse_landmarks_file(landmarks) if isinstance(landmarks, str) else landmarks
        self.attr_groups, self.mask_groups = attr_groups, mask_groups
        self.det_threshold, self.enh_threshold = det_threshold, enh_threshold
        self.batch_size, self.num_processes = batch_size, num_processes
        self.device = torch.device(device) if isinstance(device, str) else device
        self.num_std_landmarks = 5
        #: optional {"det"|"enh"|"par": state_dict} overriding the torch.hub cache lookup (offline use)
        self.state_dicts = state_dicts or {}
        if self.padding not in _abi.BORDERS:
            raise AttributeError(f"module 'cv2' has no attribute 'BORDER_{self.padding.upper()}'")
        self._init_models()
        self._init_landmarks_target()

    # ------------------------------------------------------------------------------------------------ set-up
    def _init_models(self):
        """cropper.py:346-390.  Safe to call again from pool workers: the models are created once and shared."""
        if getattr(self, "_models_ready", False):
            return
        self.det_model = self.enh_model = self.par_model = None
        self.ctx = get_context(self.device)
        if self.det_threshold is not None and self.landmarks is None:
            self.det_model = RetinaFace(self.strategy, self.det_threshold).load(self.device, self.state_dicts.get("det"))
        if self.enh_threshold is not None:
            self.enh_model = RRDBNet(self.enh_threshold).load(self.device, self.state_dicts.get("enh"))
        if self.attr_groups is not None or self.mask_groups is not None:
            self.par_model = BiSeNet(self.attr_groups, self.mask_groups, self.batch_size).load(self.device, self.state_dicts.get("par"))
        self._models_ready = True

    def _init_landmarks_target(self):
        self.landmarks_target = landmarks_target(self.output_size, self.face_factor, self.num_std_landmarks)

    # ------------------------------------------------------------------------------------------------- align
    def crop_align(self, images, padding, indices, landmarks_source):
        """``Cropper.crop_align`` (cropper.py:441-552): u8 [F,h,w,3] crops (``np.array([])`` when nothing survives)."""
        if len(indices) == 0:
            return np.array([])
        imgs = list(images) if isinstance(images, (list, tuple)) else np.ascontiguousarray(images)
        with _lock:
            bind_stream(self.ctx)
            crops, _, valid = self.ctx.align(imgs, padding, indices, landmarks_source, self.landmarks_target,
                                             self.output_size, self.padding, self.allow_skew)
        return crops[valid] if valid.any() else np.array([])       # un-estimable transforms are skipped (:529-531)

    # -------------------------------------------------------------------------------------------------- save
    def save_group(self, faces, file_names, output_dir):
        """cropper.py:554-609."""
        import cv2
        if len(faces) == 0:
            return
        os.makedirs(output_dir, exist_ok=True)
        seen = defaultdict(int)
        for face, file_name in zip(faces, file_names):
            stem, ext = os.path.splitext(file_name)
            if self.output_format is not None:
                ext = "." + self.output_format
            if self.strategy == "all":
                stem += f"_{seen[file_name]}"
                seen[file_name] += 1
            if face.ndim == 3:
                face = cv2.cvtColor(face, cv2.COLOR_RGB2BGR)
            cv2.imwrite(os.path.join(output_dir, stem + ext), face)

    def save_groups(self, faces, file_names, output_dir, attr_groups, mask_groups):
        """cropper.py:611-746: <output_dir>/<attr group>/<mask group>[ _mask]/<file>."""
        if attr_groups is None:
            attr_groups = {"": list(range(len(faces)))}
        if mask_groups is None:
            mask_groups = {"": (list(range(len(faces))), None)}
        for attr_name, attr_idx in attr_groups.items():
            for mask_name, (mask_idx, masks) in mask_groups.items():
                group = list(set(attr_idx) & set(mask_idx))
                out = os.path.join(output_dir, attr_name, mask_name)
                self.save_group([faces[i] for i in group], file_names[group], out)
                if masks is not None:
                    self.save_group(masks[[mask_idx.index(i) for i in group]], file_names[group], out + "_mask")

    # ------------------------------------------------------------------------------------------------- batch
    def process_batch(self, file_names, input_dir, output_dir):
        """cropper.py:748-850."""
        images, file_names = read_images(file_names, input_dir)
        if len(images) == 0:
            return
        groups, paddings = (None, None), None
        if self.landmarks is None and self.det_model is None:
            indices, landmarks = list(range(len(file_names))), None
        elif self.landmarks is not None:
            indices, ldm_rows = [], []
            for i, name in enumerate(file_names):
                rows = np.where(name == self.landmarks[1])[0]
                indices += [i] * len(rows)
                ldm_rows += rows.tolist()
            landmarks = self.landmarks[0][ldm_rows]
        else:
            # ingest straight into a device batch: resize + pad (fcp_as_batch) and the detect -> (enhance) -> align -> parse
            # call share it, the images never come back to the host between the stages
            batch = torch.empty((len(images), self.resize_size[1], self.resize_size[0], 3), dtype=torch.uint8,
                                device=self.device)
            with _lock:
                bind_stream(self.ctx)
                _, _, paddings = self.ctx.as_batch(images, self.resize_size, out=batch)
            return self._process_detected_batch(batch, paddings, file_names, output_dir)
        if landmarks is not None and len(landmarks) == 0:
            return
        if landmarks is not None and landmarks.shape[1] != self.num_std_landmarks:
            get_ldm_slices(self.num_std_landmarks, landmarks.shape[1])        # ValueError for unsupported counts, like the reference
            with _lock:
                landmarks = self.ctx.reduce_landmarks(landmarks)               # the slice means, on the device
        if self.enh_model is not None:
            # no detector: the images are the caller's ragged list (cropper.py:833-836 on a list of tensors); uint8 in/out
            images = self.enh_model.predict_u8(list(images), landmarks, indices)
        if landmarks is not None:
            images = self.crop_align(images, paddings, indices, landmarks)
        if self.par_model is not None and len(images):
            groups = self.par_model.predict_u8(np.stack(images) if isinstance(images, list) else images)
        self.save_groups(images, file_names[indices], output_dir, *groups)

    def _process_detected_batch(self, images, paddings, file_names, output_dir):
        """detect -> un-pad -> (enhance) -> align -> parse of cropper.py:815-847 as ONE library call on the uint8 batch."""
        with _lock:
            bind_stream(self.ctx)
            if self.par_model is not None:
                self.ctx.set_micro_batch(32, max(int(self.batch_size), 1))
            batch = images if hasattr(images, "data_ptr") else np.ascontiguousarray(images)
            # enhancement (cropper.py:833-836) is a stage of the same call: gate, RRDBNet and the warp source stay on the device
            self.ctx.set_enhance(self.enh_model.min_face_factor if self.enh_model is not None else None)
            try:
                out = self.ctx.pipeline(batch, paddings, self.landmarks_target, self.output_size,
                                        self.det_threshold, self.det_model.nms_threshold, self.strategy, self.padding,
                                        self.allow_skew, parse=self.par_model is not None)
            finally:
                self.ctx.set_enhance(None)
        if out["count"] == 0:
            return
        valid = out["valid"].astype(bool)
        faces, indices = out["crops"][valid], out["indices"][valid].tolist()
        groups = (None, None)
        if self.par_model is not None:
            groups = self.par_model.groups_from(out["labels"][valid], out["hist"][valid])
        self.save_groups(faces, file_names[indices], output_dir, *groups)

    def process_dir(self, input_dir, output_dir=None, desc="Processing"):
        """cropper.py:852-909."""
        if output_dir is None:
            output_dir = input_dir + "_faces"
        files = os.listdir(input_dir)
        batches = [files[i:i + self.batch_size] for i in range(0, len(files), self.batch_size)]
        if not batches:
            return
        worker = partial(self.process_batch, input_dir=input_dir, output_dir=output_dir)
        with ThreadPool(self.num_processes, self._init_models) as pool:
            it = pool.imap_unordered(worker, batches)
            if desc is not None:
                import tqdm
                it = tqdm.tqdm(it, total=len(batches), desc=desc)
            list(it)
