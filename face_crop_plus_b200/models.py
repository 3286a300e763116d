"""Host-side mirrors of the reference model classes, backed by libfcpb200.so.

Same constructor arguments, attributes, ``.load(device)`` and ``.predict(...)`` contracts as
``face_crop_plus.models.{RetinaFace,RRDBNet,BiSeNet}`` (models/retinaface.py:54,410; models/rrdb.py:37,83;
models/bise.py:122,327) so ``Cropper`` and user code can switch by changing the import.  No torch ``nn.Module``
is built: ``load`` feeds the reference-format state_dict to the CUDA library, ``predict`` calls its kernels.
"""
from __future__ import annotations

import threading
import warnings

import numpy as np
import torch

from . import _abi

URL_ROOT = "https://github.com/mantasu/face-crop-plus/releases/download/v1.0.0/"   # _layers.py:13
_contexts: dict[int, "_abi.Context"] = {}
_lock = threading.RLock()          # the C context is thread-compatible, not thread-safe (ThreadPool workers share it)


def device_index(device) -> int:
    """Maps a torch-style device to a CUDA ordinal.  There is no CPU path: 'cpu' selects cuda:0 with a warning."""
    dev = torch.device(device) if not isinstance(device, torch.device) else device
    if dev.type != "cuda":
        warnings.warn("face_crop_plus_b200 has no CPU path; using cuda:0", stacklevel=3)
        return 0
    return dev.index if dev.index is not None else torch.cuda.current_device()


def get_context(device) -> "_abi.Context":
    idx = device_index(device)
    with _lock:
        if idx not in _contexts:
            _contexts[idx] = _abi.Context(idx)
        return _contexts[idx]


def bind_stream(ctx: "_abi.Context") -> None:
    """Runs the library on torch's CURRENT stream of the context's device, so that tensors produced by torch ops
    (``.to(device)``, ``permute().contiguous()`` ...) are ordered before the kernels that read their ``data_ptr``.
    Call it (under ``_lock``) before every C-ABI call that may receive a CUDA tensor."""
    stream = torch.cuda.current_stream(ctx.device).cuda_stream
    if getattr(ctx, "_bound_stream", None) != stream:
        ctx.set_stream(stream)
        ctx._bound_stream = stream


class LoadMixin:
    """``LoadMixin`` of the reference (_layers.py:12-35): same cache location and URL, weights go to the CUDA library."""
    WEIGHTS_FILENAME: str | None = None
    MODEL_ID: int = -1

    def get_weights(self, device="cpu"):
        if self.WEIGHTS_FILENAME is None:
            raise ValueError("Please ensure 'WEIGHTS_FILENAME' is specified for the class that inherits this mixin.")
        return torch.hub.load_state_dict_from_url(URL_ROOT + self.WEIGHTS_FILENAME, map_location="cpu")

    def load(self, device="cuda:0", state_dict=None):
        """Loads the weights (``state_dict`` overrides the hub cache lookup) onto ``device`` and returns ``self``."""
        self.ctx = get_context(device)
        self.device = torch.device("cuda", self.ctx.device)
        sd = state_dict if state_dict is not None else self.get_weights()
        with _lock:
            self.ctx.load_state_dict(self.MODEL_ID, sd, getattr(self, "num_blocks", 23))
        return self


def _as_u8_nhwc(images) -> np.ndarray | torch.Tensor:
    """f32 [N,3,H,W] 0..255 (tensor / list of tensors / array) -> contiguous u8 [N,H,W,3] on the same device."""
    if isinstance(images, (list, tuple)):
        images = torch.stack([torch.as_tensor(i) for i in images])
    t = torch.as_tensor(images)
    if t.dtype != torch.uint8:
        t = t.round().clamp(0, 255).to(torch.uint8)
    return t.permute(0, 2, 3, 1).contiguous()


class RetinaFace(LoadMixin):
    WEIGHTS_FILENAME = "retinaface_detector.pth"
    MODEL_ID = _abi.MODEL_RETINAFACE

    def __init__(self, strategy: str = "all", vis: float = 0.6):
        self.strategy = strategy
        self.vis_threshold = vis
        self.nms_threshold = 0.4
        self.variance = [0.1, 0.2]

    def predict_u8(self, images_u8_nhwc):
        """Fast path: uint8 NHWC batch (numpy or CUDA tensor) -> (landmarks f32[F,5,2], indices list[int])."""
        if self.strategy not in _abi.STRATEGIES:
            raise ValueError(f"Unsupported startegy: {self.strategy}")          # retinaface.py:400 (sic)
        with _lock:
            bind_stream(self.ctx)
            out = self.ctx.detect(images_u8_nhwc, self.vis_threshold, self.nms_threshold, self.strategy)
        return out["landmarks"], out["indices"].tolist()

    @torch.no_grad()
    def predict(self, images):
        """``RetinaFace.predict`` (retinaface.py:410-470): images f32 [N,3,H,W] RGB 0..255."""
        return self.predict_u8(_as_u8_nhwc(images))


class RRDBNet(LoadMixin):
    WEIGHTS_FILENAME = "bsrgan_x4_enhancer.pth"
    MODEL_ID = _abi.MODEL_RRDBNET

    def __init__(self, min_face_factor: float = 0.001, num_blocks: int = 23):
        self.min_face_factor = min_face_factor
        self.num_blocks = num_blocks

    def gate(self, n: int, height: int, width: int, landmarks, indices) -> np.ndarray:
        """Which images get enhanced (rrdb.py:124-140; face area is normalised by image 0's size — reference quirk)."""
        g = np.zeros(n, np.uint8)
        for i in range(n):
            if landmarks is None or indices is None:
                g[i] = 1
                continue
            sel = landmarks[[idx == i for idx in indices]]
            if len(sel) == 0:
                continue
            w, h = (sel[:, 4] - sel[:, 0]).T
            g[i] = (w * h / (height * width)).mean() <= self.min_face_factor
        return g

    def predict_u8(self, images, landmarks=None, indices=None):
        """Fast path of :meth:`predict` on uint8 NHWC data (an [N,H,W,3] array / CUDA tensor, or a list of [h,w,3] arrays):
        the reference's result ``round(clamp(.)*255)`` is integral, so nothing is lost and no float32 NCHW copy is made.
        Enhances the gated images in place and returns the container."""
        is_list = isinstance(images, (list, tuple))
        n = len(images)
        if n == 0:
            return images
        h0, w0 = images[0].shape[:2]
        if landmarks is None or indices is None:
            g = np.ones(n, np.uint8)
        else:
            g = self.gate(n, h0, w0, np.asarray(landmarks, dtype=np.float32), list(indices))
        if not g.any():
            return images
        with _lock:
            bind_stream(self.ctx)
            if not is_list:
                self.ctx.enhance_u8(images, g)
            else:
                for i in np.nonzero(g)[0]:
                    one = np.ascontiguousarray(images[i])[None]
                    self.ctx.enhance_u8(one, None)
                    images[i] = one[0]
        return images

    @torch.no_grad()
    def predict(self, images, landmarks=None, indices=None):
        """``RRDBNet.predict`` (rrdb.py:83-146): enhances the gated images IN PLACE and returns the same container."""
        if isinstance(images, (list, tuple)):
            for i, img in enumerate(images):
                if landmarks is None or indices is None:
                    enhance = True                                       # rrdb.py:125-127: no landmarks -> every image
                else:
                    sub_l = landmarks[[idx == i for idx in indices]]
                    # normalised by images[0]'s size like the reference (rrdb.py:139)
                    enhance = bool(self.gate(1, *images[0].shape[1:], sub_l, [0] * len(sub_l))[0])
                if enhance:
                    batch = img.unsqueeze(0).contiguous()
                    with _lock:
                        bind_stream(self.ctx)
                        self.ctx.enhance(batch, None)
                    images[i] = batch[0]
            return images
        n, _, h, w = images.shape
        g = self.gate(n, h, w, landmarks, indices)
        if g.any():
            work = images if images.is_contiguous() else images.contiguous()
            with _lock:
                bind_stream(self.ctx)
                self.ctx.enhance(work.numpy() if (isinstance(work, torch.Tensor) and not work.is_cuda) else work, g)
            if work is not images:
                images.copy_(work)
        return images


class BiSeNet(LoadMixin):
    WEIGHTS_FILENAME = "bise_parser.pth"
    MODEL_ID = _abi.MODEL_BISENET

    def __init__(self, attr_groups=None, mask_groups=None, max_batch_size: int = 8):
        self.attr_groups = attr_groups
        self.mask_groups = mask_groups
        self.batch_size = max_batch_size
        self.attr_join_by_and = True
        self.attr_threshold = 5
        self.mask_threshold = 10
        self.mean = [0.485, 0.456, 0.406]
        self.std = [0.229, 0.224, 0.225]

    def groups_from(self, labels: np.ndarray, hist: np.ndarray):
        """Grouping rules of bise.py:214-325,407-416: membership flags and the 0/255 mask images come from the device
        (``fcp_group``: integer compares on the per-class histogram, one pass over the labels for all mask groups)."""
        attr_groups = mask_groups = None
        if len(hist) == 0:
            return ({} if self.attr_groups is not None else None), ({} if self.mask_groups is not None else None)
        with _lock:
            bind_stream(self.ctx)
            attr_m, mask_m, masks = self.ctx.group(np.ascontiguousarray(labels), np.ascontiguousarray(hist, dtype=np.int32),
                                                   self.attr_groups, self.mask_groups, self.attr_threshold, self.mask_threshold,
                                                   self.attr_join_by_and)
        if self.attr_groups is not None:
            attr_groups = {k: np.nonzero(attr_m[g])[0].tolist() for g, k in enumerate(self.attr_groups) if attr_m[g].any()}
        if self.mask_groups is not None:
            mask_groups = {}
            for m, k in enumerate(self.mask_groups):
                idx = np.nonzero(mask_m[m])[0].tolist()
                if idx:                                                        # empty groups are dropped (bise.py:407-416)
                    mask_groups[k] = (idx, masks[m][idx])
        return attr_groups, mask_groups

    def predict_u8(self, crops_u8_nhwc: np.ndarray):
        with _lock:
            bind_stream(self.ctx)
            self.ctx.set_micro_batch(32, max(int(self.batch_size), 1))
            labels, hist = self.ctx.parse(np.ascontiguousarray(crops_u8_nhwc))
        return self.groups_from(labels, hist)

    @torch.no_grad()
    def predict(self, images):
        """``BiSeNet.predict`` (bise.py:327-418): images f32 [F,3,h,w] 0..255 (tensor or list of tensors)."""
        u8 = _as_u8_nhwc(images)
        return self.predict_u8(u8.cpu().numpy() if isinstance(u8, torch.Tensor) else u8)
