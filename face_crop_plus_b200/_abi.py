"""ctypes binding of libfcpb200.so (declared in include/fcp_b200.h).

There is NO CPU fallback: importing works anywhere (so CPU-only tooling can inspect the package), but creating a
:class:`Context` without the built library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("FCP_B200_LIB", _HERE / "libfcpb200.so"))

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_CAPACITY = 0, 1, 2, 3, 4
MODEL_RETINAFACE, MODEL_BISENET, MODEL_RRDBNET = 0, 1, 2
STRATEGIES = {"all": 0, "best": 1, "largest": 2}
BORDERS = {"constant": 0, "replicate": 1, "reflect": 2, "wrap": 3, "reflect_101": 4}
ACTS = {"none": 0, "relu": 1, "lrelu": 2, "sigmoid": 3}

_p = C.c_void_p
_i = C.c_int
_f = C.c_float

#: every symbol include/fcp_b200.h declares -> (restype, argtypes)
SIGNATURES = {
    "fcp_create": (_i, [_i, C.POINTER(_p)]),
    "fcp_destroy": (None, [_p]),
    "fcp_last_error": (C.c_char_p, [_p]),
    "fcp_version": (C.c_char_p, []),
    "fcp_set_stream": (_i, [_p, _p]),
    "fcp_sync": (_i, [_p]),
    "fcp_launch_count": (C.c_int64, [_p]),
    "fcp_set_micro_batch": (_i, [_p, _i, _i]),
    "fcp_set_conv_impl": (_i, [_p, _i]),
    "fcp_profile": (_i, [_p, _i]),
    "fcp_profile_read": (_i, [_p, C.POINTER(C.c_double)]),
    "fcp_profile_stages": (_i, [_p, C.POINTER(C.c_double)]),
    "fcp_load_tensor": (_i, [_p, _i, C.c_char_p, _p, C.POINTER(C.c_int64), _i]),
    "fcp_finalize": (_i, [_p, _i, _i]),
    "fcp_detect": (_i, [_p, _p, _i, _i, _i, _f, _f, _i, _i, _p, _p, _p, _p, _p, _p]),
    "fcp_detect_heads": (_i, [_p, _p, _i, _i, _i, _p]),
    "fcp_detect_post": (_i, [_p, _p, _i, _i, _i, _f, _f, _i, _i, _p, _p, _p, _p, _p, _p]),
    "fcp_align": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _i, _p, _i, _i, _i, _i, _p, _p, _p]),
    "fcp_align_list": (_i, [_p, _p, _p, _p, _i, _p, _p, _p, _i, _p, _i, _i, _i, _i, _p, _p, _p]),
    "fcp_reduce_landmarks": (_i, [_p, _p, _i, _i, _p]),
    "fcp_set_cubic_mode": (_i, [_p, _i]),
    "fcp_as_batch": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "fcp_parse": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "fcp_parse_logits": (_i, [_p, _p, _i, _i, _i, _p]),
    "fcp_parse_tail": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "fcp_masks": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "fcp_group": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _i, _i, _i, _p, _i, _i, _p, _p, _p]),
    "fcp_enhance": (_i, [_p, _p, _i, _i, _i, _p]),
    "fcp_enhance_forward": (_i, [_p, _p, _i, _i, _i, _p]),
    "fcp_enhance_u8": (_i, [_p, _p, _i, _i, _i, _p]),
    "fcp_enhance_gate": (_i, [_p, _p, _p, _i, _i, _i, _i, _f, _p]),
    "fcp_set_enhance": (_i, [_p, _i, _f]),
    "fcp_comm_unique_id": (_i, [_p, _p]),
    "fcp_comm_init": (_i, [_p, _i, _i, _p]),
    "fcp_comm_destroy": (None, [_p]),
    "fcp_allgather_meta": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "fcp_set_gather": (_i, [_p, _p, _i, _i]),
    "fcp_pipeline": (_i, [_p, _p, _i, _i, _i, _p, _f, _f, _i, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "fcp_conv2d": (_i, [_p, _p, _i, _i, _i, _i, _p, _i, _i, _i, _i, _p, _p, _p, _i, _f, _i, _p]),
}

_lib = None


class FcpError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libfcpb200 error {code}: {message}")
        self.code = code


def load_library() -> C.CDLL:
    """Loads libfcpb200.so and declares the prototypes; raises if the library was not built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m face_crop_plus_b200.build` "
                               "(there is no CPU fallback)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _ptr(x):
    """Pointer of a numpy array, a torch tensor (host or CUDA), an int address, or None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"], "arrays crossing the C ABI must be C-contiguous"
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        assert x.is_contiguous(), "tensors crossing the C ABI must be contiguous"
        return x.data_ptr()
    return int(x)


class Context:
    """One libfcpb200 context = one CUDA device + one stream + the loaded models."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = _p()
        code = self.lib.fcp_create(int(device), C.byref(h))
        if code != OK:
            raise FcpError(code, f"fcp_create(device={device}) failed — a CUDA device is required, there is no CPU fallback")
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.fcp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, code: int, allow=()):
        if code != OK and code not in allow:
            raise FcpError(code, self.lib.fcp_last_error(self.h).decode())
        return code

    # ---- plumbing
    def version(self) -> str:
        return self.lib.fcp_version().decode()

    def set_stream(self, cuda_stream: int):
        self.check(self.lib.fcp_set_stream(self.h, cuda_stream))

    def sync(self):
        self.check(self.lib.fcp_sync(self.h))

    def launch_count(self) -> int:
        return int(self.lib.fcp_launch_count(self.h))

    def set_micro_batch(self, detect_images: int, parse_faces: int):
        self.check(self.lib.fcp_set_micro_batch(self.h, detect_images, parse_faces))

    def set_conv_impl(self, impl: int):
        self.check(self.lib.fcp_set_conv_impl(self.h, impl))

    def profile(self, enable: bool):
        self.check(self.lib.fcp_profile(self.h, int(enable)))

    def profile_read(self) -> dict:
        out = (C.c_double * 4)()
        self.check(self.lib.fcp_profile_read(self.h, out))
        return dict(conv_ms=out[0], conv_launches=int(out[1]), conv_flops=out[2], conv_bytes=out[3])

    STAGES = ("detect_net", "detect_post", "enhance", "align", "parse_net", "parse_tail", "gather", "ingest")

    def profile_stages(self) -> dict:
        """Milliseconds per stage accumulated since the last read (only while ``profile(True)``)."""
        out = (C.c_double * 8)()
        self.check(self.lib.fcp_profile_stages(self.h, out))
        return dict(zip(self.STAGES, [float(v) for v in out]))

    # ---- multi-GPU: the one collective (metadata all-gather over NCCL)
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self.check(self.lib.fcp_comm_unique_id(self.h, buf))
        return buf.raw

    def comm_init(self, rank: int, world: int, unique_id: bytes | None):
        self.check(self.lib.fcp_comm_init(self.h, int(rank), int(world), unique_id))
        self.comm_rank, self.comm_world = int(rank), int(world)

    def set_gather(self, out_records, cap: int = 0, index_base: int = 0):
        """``out_records``: CUDA float64 tensor [world, cap + 1, 20] (or None = off); see ``fcp_set_gather``."""
        self.check(self.lib.fcp_set_gather(self.h, _ptr(out_records), int(cap), int(index_base)))
        self._gather_keepalive = out_records

    def allgather_meta(self, landmarks, indices, matrices, valid, cap: int, index_base: int, out_records=None):
        """Explicit form of the collective: returns float64 [world, cap + 1, 20] (numpy unless ``out_records`` is given)."""
        f = len(indices)
        world = getattr(self, "comm_world", 1)
        lms = np.ascontiguousarray(landmarks, dtype=np.float32).reshape(f, 10) if not hasattr(landmarks, "data_ptr") else landmarks
        idx = np.ascontiguousarray(indices, dtype=np.int32) if not hasattr(indices, "data_ptr") else indices
        mats = None if matrices is None else (np.ascontiguousarray(matrices, dtype=np.float64) if not hasattr(matrices, "data_ptr") else matrices)
        val = None if valid is None else (np.ascontiguousarray(valid, dtype=np.uint8) if not hasattr(valid, "data_ptr") else valid)
        out = np.zeros((world, cap + 1, 20), np.float64) if out_records is None else out_records
        self.check(self.lib.fcp_allgather_meta(self.h, _ptr(lms), _ptr(idx), _ptr(mats), _ptr(val), f, int(cap), int(index_base), _ptr(out)))
        return out

    def load_state_dict(self, model: int, state_dict, rrdb_blocks: int = 23):
        """Feeds a reference-format state_dict (torch tensors or numpy arrays) and finalizes the model."""
        for key, value in state_dict.items():
            if key.endswith("num_batches_tracked"):
                continue
            arr = value.detach().cpu().numpy() if hasattr(value, "detach") else np.asarray(value)
            arr = np.ascontiguousarray(arr, dtype=np.float32)
            shape = (C.c_int64 * max(arr.ndim, 1))(*arr.shape)
            self.check(self.lib.fcp_load_tensor(self.h, model, key.encode(), arr.ctypes.data, shape, arr.ndim))
        self.check(self.lib.fcp_finalize(self.h, model, rrdb_blocks))

    # ---- detect
    def _det_outputs(self, cap):
        return (np.empty((cap, 5, 2), np.float32), np.empty(cap, np.int32), np.empty((cap, 4), np.float32),
                np.empty(cap, np.float32), np.empty(cap, np.int32), np.zeros(1, np.int32))

    def detect(self, images, vis_threshold=0.6, nms_threshold=0.4, strategy="all", max_faces=None, n=None, h=None, w=None):
        """images: u8 [n,h,w,3] numpy array or CUDA tensor.  Returns dict(landmarks, indices, boxes, scores, anchors)."""
        if n is None:
            n, h, w = images.shape[:3]
        cap = max_faces or (n if strategy != "all" else max(64 * n, 1024))
        while True:
            lm, idx, box, sc, an, cnt = self._det_outputs(cap)
            code = self.lib.fcp_detect(self.h, _ptr(images), n, h, w, vis_threshold, nms_threshold, STRATEGIES[strategy],
                                       cap, _ptr(lm), _ptr(idx), _ptr(box), _ptr(sc), _ptr(an), _ptr(cnt))
            self.check(code, allow=(ERR_CAPACITY,))
            if code == OK:
                k = int(cnt[0])
                return dict(landmarks=lm[:k], indices=idx[:k], boxes=box[:k], scores=sc[:k], anchors=an[:k])
            cap = int(cnt[0])

    def detect_heads(self, images):
        n, h, w = images.shape[:3]
        a = num_priors(h, w)
        out = np.empty((n, a, 16), np.float32)
        self.check(self.lib.fcp_detect_heads(self.h, _ptr(images), n, h, w, _ptr(out)))
        return out

    def detect_post(self, heads, h, w, vis_threshold=0.6, nms_threshold=0.4, strategy="all", max_faces=None):
        heads = np.ascontiguousarray(heads, dtype=np.float32)
        n = heads.shape[0]
        cap = max_faces or (n if strategy != "all" else max(64 * n, 1024))
        while True:
            lm, idx, box, sc, an, cnt = self._det_outputs(cap)
            code = self.lib.fcp_detect_post(self.h, _ptr(heads), n, h, w, vis_threshold, nms_threshold, STRATEGIES[strategy],
                                            cap, _ptr(lm), _ptr(idx), _ptr(box), _ptr(sc), _ptr(an), _ptr(cnt))
            self.check(code, allow=(ERR_CAPACITY,))
            if code == OK:
                k = int(cnt[0])
                return dict(landmarks=lm[:k], indices=idx[:k], boxes=box[:k], scores=sc[:k], anchors=an[:k])
            cap = int(cnt[0])

    # ---- align
    def align(self, images, paddings, indices, landmarks, target, out_size=(256, 256), border="constant", allow_skew=False):
        """images: u8 [n,h,w,3] array, or a list of u8 [h_i,w_i,3] arrays.  Returns (crops, matrices, valid)."""
        f = len(indices)
        ow, oh = int(out_size[0]), int(out_size[1])
        crops = np.empty((f, oh, ow, 3), np.uint8)
        mats = np.empty((f, 2, 3), np.float64)
        valid = np.zeros(f, np.uint8)
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        lms = np.ascontiguousarray(landmarks, dtype=np.float32).reshape(f, 5, 2)
        tgt = np.ascontiguousarray(target, dtype=np.float32)
        pad = None if paddings is None else np.ascontiguousarray(paddings, dtype=np.int32)
        bm = BORDERS[border] if isinstance(border, str) else int(border)
        if isinstance(images, (list, tuple)):
            imgs = [np.ascontiguousarray(im) for im in images]
            n = len(imgs)
            ptrs = (C.c_void_p * max(n, 1))(*[im.ctypes.data for im in imgs])
            hs = np.array([im.shape[0] for im in imgs], np.int32)
            ws = np.array([im.shape[1] for im in imgs], np.int32)
            self.check(self.lib.fcp_align_list(self.h, C.cast(ptrs, _p), _ptr(hs), _ptr(ws), n, _ptr(pad), _ptr(idx), _ptr(lms),
                                               f, _ptr(tgt), ow, oh, bm, int(allow_skew), _ptr(crops), _ptr(mats), _ptr(valid)))
        else:
            n, h, w = images.shape[:3]
            self.check(self.lib.fcp_align(self.h, _ptr(images), n, h, w, _ptr(pad), _ptr(idx), _ptr(lms), f, _ptr(tgt), ow, oh,
                                          bm, int(allow_skew), _ptr(crops), _ptr(mats), _ptr(valid)))
        return crops, mats, valid.astype(bool)

    def reduce_landmarks(self, landmarks):
        """f32 [F,K,2] -> f32 [F,5,2]: slice means of utils.py:90-168 / cropper.py:828-831 (ValueError for an unsupported K)."""
        lms = np.ascontiguousarray(landmarks, dtype=np.float32)
        f, k = lms.shape[:2]
        out = np.empty((f, 5, 2), np.float32)
        code = self.lib.fcp_reduce_landmarks(self.h, _ptr(lms), f, k, _ptr(out))
        if code == ERR_INVALID:
            raise ValueError((self.lib.fcp_last_error(self.h) or b"").decode())
        self.check(code)
        return out

    # ---- ingest
    def set_cubic_mode(self, floating_point: bool):
        """INTER_CUBIC arithmetic of :meth:`as_batch`: True (default) = what IPP-enabled cv2 builds compute, False = OpenCV's
        own fixed-point code (``fcp_set_cubic_mode``)."""
        self.check(self.lib.fcp_set_cubic_mode(self.h, int(bool(floating_point))))

    def as_batch(self, images, size=512, padding_mode="constant", out=None):
        """``utils.as_batch`` (utils.py:273-342) on the device: list of u8 [h_i,w_i,3] arrays (or CUDA uint8 tensors) ->
        (batch u8 [n,H,W,3], unscales f64 [n], paddings i64 [n,4]).  ``out``: optional preallocated batch (numpy array or
        CUDA uint8 tensor), returned in place of a new numpy array."""
        sw, sh = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))
        imgs = [im if hasattr(im, "data_ptr") else np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        n = len(imgs)
        for im in imgs:
            if im.ndim != 3 or im.shape[2] != 3:
                raise ValueError("as_batch expects HxWx3 uint8 images")
        ptrs = (C.c_void_p * max(n, 1))(*[_ptr(im) for im in imgs])
        hs = np.array([im.shape[0] for im in imgs], np.int32)
        ws = np.array([im.shape[1] for im in imgs], np.int32)
        batch = np.empty((n, sh, sw, 3), np.uint8) if out is None else out
        unscales, pads = np.zeros(n, np.float64), np.zeros((n, 4), np.int32)
        bm = BORDERS[padding_mode.lower()] if isinstance(padding_mode, str) else int(padding_mode)
        self.check(self.lib.fcp_as_batch(self.h, C.cast(ptrs, _p), _ptr(hs), _ptr(ws), n, sw, sh, bm, _ptr(batch), _ptr(unscales), _ptr(pads)))
        return batch, unscales, pads.astype(np.int64)

    # ---- parse
    def parse(self, crops):
        f, h, w = crops.shape[:3]
        labels = np.empty((f, h, w), np.uint8)
        hist = np.zeros((f, 19), np.int32)
        self.check(self.lib.fcp_parse(self.h, _ptr(crops), f, h, w, _ptr(labels), _ptr(hist)))
        return labels, hist

    def parse_logits(self, crops):
        f, h, w = crops.shape[:3]
        out = np.empty((f, 19, 64, 64), np.float32)
        self.check(self.lib.fcp_parse_logits(self.h, _ptr(crops), f, h, w, _ptr(out)))
        return out

    def parse_tail(self, logits, h, w):
        logits = np.ascontiguousarray(logits, dtype=np.float32)
        f = logits.shape[0]
        labels = np.empty((f, h, w), np.uint8)
        hist = np.zeros((f, 19), np.int32)
        self.check(self.lib.fcp_parse_tail(self.h, _ptr(logits), f, h, w, _ptr(labels), _ptr(hist)))
        return labels, hist

    def masks(self, labels, classes):
        labels = np.ascontiguousarray(labels, dtype=np.uint8)
        f, h, w = labels.shape
        lut = np.zeros(19, np.uint8)
        lut[[c for c in classes if 0 <= c < 19]] = 1
        out = np.empty((f, h, w), np.uint8)
        self.check(self.lib.fcp_masks(self.h, _ptr(labels), f, h, w, _ptr(lut), _ptr(out)))
        return out

    def group(self, labels, hist, attr_groups=None, mask_groups=None, attr_threshold=5, mask_threshold=10, join_and=True,
              want_masks=True):
        """Device-side ``group_by_attributes`` / ``group_by_masks`` (bise.py:214-325).  labels u8 [f,h,w] and hist i32 [f,19]
        may be numpy arrays or CUDA tensors.  Returns (attr_member bool [n_attr,f], mask_member bool [n_mask,f],
        masks u8 [n_mask,f,h,w] or None), rows in the dictionaries' key order."""
        f, h, w = labels.shape
        attr_lists = [list(v) for v in (attr_groups or {}).values()]
        mask_lists = [list(v) for v in (mask_groups or {}).values()]
        n_attr, n_mask = len(attr_lists), len(mask_lists)
        codes = np.array([a for v in attr_lists for a in v] or [0], np.int32)
        offs = np.cumsum([0] + [len(v) for v in attr_lists]).astype(np.int32)
        lut = np.zeros((max(n_mask, 1), 19), np.uint8)
        for m, v in enumerate(mask_lists):
            lut[m, [c for c in v if 0 <= c < 19]] = 1
        out_attr, out_mask = np.zeros((n_attr, f), np.uint8), np.zeros((n_mask, f), np.uint8)
        masks = np.empty((n_mask, f, h, w), np.uint8) if (want_masks and n_mask) else None
        self.check(self.lib.fcp_group(self.h, _ptr(labels), _ptr(hist), f, h, w, _ptr(codes), _ptr(offs), n_attr, int(attr_threshold),
                                      int(bool(join_and)), _ptr(lut), n_mask, int(mask_threshold), _ptr(out_attr), _ptr(out_mask),
                                      _ptr(masks)))
        return out_attr.astype(bool), out_mask.astype(bool), masks

    # ---- enhance
    def enhance(self, images_nchw, gate=None):
        """In place on a float32 [n,3,h,w] numpy array or CUDA tensor."""
        n, _, h, w = images_nchw.shape
        g = None if gate is None else np.ascontiguousarray(gate, dtype=np.uint8)
        self.check(self.lib.fcp_enhance(self.h, _ptr(images_nchw), n, h, w, _ptr(g)))
        return images_nchw

    def enhance_u8(self, images_nhwc, gate=None):
        """In place on a uint8 [n,h,w,3] numpy array or CUDA tensor (``fcp_enhance_u8``)."""
        n, h, w = images_nhwc.shape[:3]
        g = None if gate is None else np.ascontiguousarray(gate, dtype=np.uint8)
        self.check(self.lib.fcp_enhance_u8(self.h, _ptr(images_nhwc), n, h, w, _ptr(g)))
        return images_nhwc

    def enhance_gate(self, landmarks, indices, n: int, h: int, w: int, min_face_factor: float) -> np.ndarray:
        """Device evaluation of the gate of ``RRDBNet.predict`` (rrdb.py:124-141): uint8 [n]."""
        idx = np.ascontiguousarray(indices, dtype=np.int32)
        lms = np.ascontiguousarray(landmarks, dtype=np.float32).reshape(len(idx), 10)
        out = np.zeros(n, np.uint8)
        self.check(self.lib.fcp_enhance_gate(self.h, _ptr(lms), _ptr(idx), len(idx), n, h, w, float(min_face_factor), _ptr(out)))
        return out

    def set_enhance(self, min_face_factor: float | None):
        """Enhancement stage of :meth:`pipeline`: the ``min_face_factor`` of ``RRDBNet`` or None = off."""
        self.check(self.lib.fcp_set_enhance(self.h, int(min_face_factor is not None), float(min_face_factor or 0.0)))

    def enhance_forward(self, x_nchw):
        x = np.ascontiguousarray(x_nchw, dtype=np.float32)
        n, _, h, w = x.shape
        out = np.empty((n, 3, 4 * h, 4 * w), np.float32)
        self.check(self.lib.fcp_enhance_forward(self.h, _ptr(x), n, h, w, _ptr(out)))
        return out

    # ---- whole path
    def pipeline(self, images, paddings, target, out_size=(256, 256), vis_threshold=0.6, nms_threshold=0.4,
                 strategy="largest", border="constant", allow_skew=False, parse=True, max_faces=None, out=None,
                 n=None, h=None, w=None):
        """detect -> un-pad -> align -> parse in one call.  ``images``/``out`` buffers may be numpy (host) or CUDA tensors."""
        if n is None:
            n, h, w = images.shape[:3]
        ow, oh = int(out_size[0]), int(out_size[1])
        cap = max_faces or (n if strategy != "all" else max(64 * n, 1024))
        tgt = np.ascontiguousarray(target, dtype=np.float32)
        pad = None if paddings is None else np.ascontiguousarray(paddings, dtype=np.int32)
        bm = BORDERS[border] if isinstance(border, str) else int(border)
        while True:
            o = out or dict(landmarks=np.empty((cap, 5, 2), np.float32), indices=np.empty(cap, np.int32),
                            crops=np.empty((cap, oh, ow, 3), np.uint8), matrices=np.empty((cap, 2, 3), np.float64),
                            valid=np.zeros(cap, np.uint8), labels=np.empty((cap, oh, ow), np.uint8) if parse else None,
                            hist=np.zeros((cap, 19), np.int32) if parse else None)
            cnt = np.zeros(1, np.int32)
            code = self.lib.fcp_pipeline(self.h, _ptr(images), n, h, w, _ptr(pad), vis_threshold, nms_threshold,
                                         STRATEGIES[strategy], _ptr(tgt), ow, oh, bm, int(allow_skew), cap,
                                         _ptr(o["landmarks"]), _ptr(o["indices"]), _ptr(cnt), _ptr(o["crops"]),
                                         _ptr(o.get("matrices")), _ptr(o.get("valid")), _ptr(o.get("labels")), _ptr(o.get("hist")))
            self.check(code, allow=(ERR_CAPACITY,))
            if code == OK:
                k = int(cnt[0])
                res = {key: (val[:k] if val is not None else None) for key, val in o.items()}
                res["count"] = k
                return res
            if out is not None:
                raise FcpError(code, "caller-provided output buffers are too small")
            cap = int(cnt[0])

    # ---- kernel test hook
    def conv2d(self, x_nhwc, weight_oihw, stride=1, pad=0, scale=None, shift=None, residual=None, act="none", slope=0.0, impl=0):
        x = np.ascontiguousarray(x_nhwc, dtype=np.float32)
        wt = np.ascontiguousarray(weight_oihw, dtype=np.float32)
        n, h, w, cin = x.shape
        cout, _, k, _ = wt.shape
        ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        out = np.empty((n, ho, wo, cout), np.float32)
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
        scale, shift, residual = f32(scale), f32(shift), f32(residual)
        self.check(self.lib.fcp_conv2d(self.h, _ptr(x), n, h, w, cin, _ptr(wt), cout, k, stride, pad, _ptr(scale), _ptr(shift),
                                       _ptr(residual), ACTS[act], slope, impl, _ptr(out)))
        return out


def num_priors(h: int, w: int) -> int:
    """Number of RetinaFace priors for an h x w input (PriorBox, _layers.py:41-62)."""
    return sum(2 * -(-h // s) * -(-w // s) for s in (8, 16, 32))
