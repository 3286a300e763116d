"""Oracle (test infrastructure): numpy restatement of the batch-ingest step, ``utils.as_batch`` (utils.py:273-342).

``as_batch`` resizes every image so that it fits ``size`` keeping its aspect ratio (``cv2.resize`` with
INTER_AREA when the image is larger than the target, INTER_CUBIC otherwise, utils.py:320,334) and centres it
with ``cv2.copyMakeBorder`` (utils.py:335).  The arithmetic lives in OpenCV (un-vendored; opencv-python 4.13.0
here).  What is restated is OpenCV's own published algorithm (imgproc ``resize.cpp``):

* INTER_AREA, both scales >= 1: integer scales sum the block in int32 (2x2: ``(s+2)>>2``; otherwise
  ``cvRound(float(s) * (1.f/area))``); fractional scales build ``DecimateAlpha`` tables in float64, accumulate
  ``buf += S*alpha`` over x in float32 and ``sum (+)= beta*buf`` over y in float32 (no FMA), ``cvRound``.
  Pinned bit-for-bit against ``cv2.resize`` (IPP on and off) in tests/test_ingest_cpu.py.
* INTER_CUBIC: coefficient tables in float32 (A = -0.75) quantised to 11-bit fixed point, an exact int32
  horizontal pass, and a vertical pass that OpenCV's SIMD code does in float32
  (``S0*b0 + (S1*b1 + (S2*b2 + S3*b3))``, mul and add rounded separately, ``cvRound``) for the first
  ``8*floor(3*width/8)`` values of a row and in integers (``(v + 2^21) >> 22``) for the rest.
  Pinned bit-for-bit against ``cv2.resize`` with ``cv2.ipp.setUseIPP(False)``.  NOTE: the opencv-python wheel of
  this image routes 8-bit INTER_CUBIC through Intel IPP (closed source), whose result differs from OpenCV's own
  code by +-1 in ~4.5 % of the values; the tests bound that difference (max 1) instead of chasing it.
"""
from __future__ import annotations

import math

import numpy as np

from .align import BORDER_MODES, _border_index

COEF_BITS = 11          # INTER_RESIZE_COEF_BITS
COEF_SCALE = 1 << COEF_BITS


# ------------------------------------------------------------------------------------------------ INTER_AREA
def area_table(ssize: int, dsize: int, scale: float):
    """``computeResizeAreaTab``: list of (dst index, src index, float32 weight), grouped by dst index."""
    tab = []
    for d in range(dsize):
        f1 = d * scale
        f2 = f1 + scale
        cell = min(scale, ssize - f1)
        s1, s2 = math.ceil(f1), math.floor(f2)
        s2 = min(s2, ssize - 1)
        s1 = min(s1, s2)
        if s1 - f1 > 1e-3:
            tab.append((d, s1 - 1, np.float32((s1 - f1) / cell)))
        for s in range(s1, s2):
            tab.append((d, s, np.float32(1.0 / cell)))
        if f2 - s2 > 1e-3:
            tab.append((d, s2, np.float32(min(min(f2 - s2, 1.0), cell) / cell)))
    return tab


def resize_area(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """``cv2.resize(src, (dw, dh), interpolation=INTER_AREA)`` for u8 HxWxC when shrinking in both directions."""
    sh, sw, cn = src.shape
    scale_x, scale_y = 1.0 / (dw / sw), 1.0 / (dh / sh)          # cv::resize: inv_scale = dsize/ssize, scale = 1/inv_scale
    assert scale_x >= 1 and scale_y >= 1, "INTER_AREA restated for shrinking only (all as_batch ever asks for)"
    ix, iy = int(np.rint(scale_x)), int(np.rint(scale_y))
    eps = np.finfo(np.float64).eps
    if abs(scale_x - ix) < eps and abs(scale_y - iy) < eps:      # is_area_fast
        blk = src[:dh * iy, :dw * ix].reshape(dh, iy, dw, ix, cn).astype(np.int32).sum((1, 3))
        if ix == 2 and iy == 2:
            return ((blk + 2) >> 2).astype(np.uint8)
        sc = np.float32(1.0) / np.float32(ix * iy)
        return np.clip(np.rint(blk.astype(np.float32) * sc), 0, 255).astype(np.uint8)
    xt, yt = area_table(sw, dw, scale_x), area_table(sh, dh, scale_y)
    srcf = src.astype(np.float32)
    buf = np.zeros((sh, dw, cn), np.float32)
    for d, s, a in xt:                                            # in table order: buf = buf + S*alpha
        buf[:, d] = buf[:, d] + srcf[:, s] * a
    out = np.zeros((dh, dw, cn), np.uint8)
    acc, prev = None, None
    for d, s, b in yt:
        if d != prev:
            if prev is not None:
                out[prev] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
            acc, prev = buf[s] * b, d
        else:
            acc = acc + buf[s] * b
    out[prev] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return out


# ----------------------------------------------------------------------------------------------- INTER_CUBIC
def cubic_coeffs(x: np.float32) -> np.ndarray:
    """``interpolateCubic`` in float32, A = -0.75."""
    f = np.float32
    A, x, one = f(-0.75), f(x), f(1)
    c0 = ((A * (x + one) - f(5) * A) * (x + one) + f(8) * A) * (x + one) - f(4) * A
    c1 = ((A + f(2)) * x - (A + f(3))) * x * x + one
    c2 = ((A + f(2)) * (one - x) - (A + f(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], dtype=np.float32)


def cubic_table(ssize: int, dsize: int):
    """Per destination index: source offset of the second tap and the four 11-bit fixed-point weights."""
    scale = 1.0 / (dsize / ssize)
    ofs, coef = np.zeros(dsize, np.int64), np.zeros((dsize, 4), np.int32)
    for d in range(dsize):
        fx = np.float32((d + 0.5) * scale - 0.5)
        sx = int(np.floor(fx))
        fx = np.float32(fx - np.float32(sx))
        coef[d] = np.clip(np.rint(cubic_coeffs(fx) * np.float32(COEF_SCALE)), -32768, 32767).astype(np.int32)
        ofs[d] = sx
    return ofs, coef


def resize_cubic(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """``cv2.resize(src, (dw, dh), interpolation=INTER_CUBIC)`` for u8 HxWxC — OpenCV's own code path (IPP off)."""
    sh, sw, cn = src.shape
    if (dw, dh) == (sw, sh):
        return src.copy()                                         # cv::resize copies when the sizes agree
    xo, xa = cubic_table(sw, dw)
    yo, ya = cubic_table(sh, dh)
    xi = np.clip(xo[:, None] + np.arange(-1, 3)[None], 0, sw - 1)                      # taps clamp to the border
    hor = (src[:, xi, :].astype(np.int64) * xa[None, :, :, None]).sum(2)               # [sh, dw, cn] exact int32
    yi = np.clip(yo[:, None] + np.arange(-1, 3)[None], 0, sh - 1)
    rows = hor[yi]                                                                     # [dh, 4, dw, cn]
    vint = (rows * ya[:, :, None, None].astype(np.int64)).sum(1)
    out_int = np.clip((vint + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
    b = ya.astype(np.float32) * np.float32(1.0 / (COEF_SCALE * COEF_SCALE))
    rf = rows.astype(np.float32)
    t = rf[:, 3] * b[:, 3, None, None]
    t = rf[:, 2] * b[:, 2, None, None] + t
    t = rf[:, 1] * b[:, 1, None, None] + t
    t = rf[:, 0] * b[:, 0, None, None] + t
    out = np.clip(np.rint(t), 0, 255).astype(np.uint8).reshape(dh, dw * cn)
    simd = (dw * cn) // 8 * 8                                                          # v_int16 lanes = 8 (SSE baseline)
    out[:, simd:] = out_int.reshape(dh, dw * cn)[:, simd:]
    return out.reshape(dh, dw, cn)


def cubic_coeffs_f64(x: float) -> np.ndarray:
    """``interpolateCubic`` (A = -0.75) in float64."""
    A = -0.75
    c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
    c1 = ((A + 2) * x - (A + 3)) * x * x + 1
    c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
    return np.array([c0, c1, c2, 1.0 - c0 - c1 - c2], dtype=np.float64)


def resize_cubic_float(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """``cv2.resize(..., INTER_CUBIC)`` as the opencv-python x86 wheels compute it for 8-bit images: they hand the call to
    Intel IPP (closed source), whose result is the separable Keys cubic (a = -0.75, taps clamped to the border) evaluated in
    floating point and rounded to nearest - NOT OpenCV's 11-bit fixed-point code.  This restatement evaluates it in float64
    (horizontal pass, then vertical, products and sums rounded separately, round-half-even); measured against cv2 with IPP
    on it differs by one grey level in < 1e-5 of the bytes (IPP's internal float32 order is not public), where the
    fixed-point path differs in ~4.5 %.  tests/test_ingest_cpu.py::test_cubic_float_vs_ipp_build."""
    sh, sw, cn = src.shape
    if (dw, dh) == (sw, sh):
        return src.copy()

    def table(ssize, dsize):
        scale = 1.0 / (dsize / ssize)
        idx, cf = np.zeros((dsize, 4), np.int64), np.zeros((dsize, 4), np.float64)
        for d in range(dsize):
            fx = (d + 0.5) * scale - 0.5
            sx = math.floor(fx)
            cf[d] = cubic_coeffs_f64(fx - sx)
            idx[d] = np.clip(np.arange(sx - 1, sx + 3), 0, ssize - 1)
        return idx, cf
    xi, xc = table(sw, dw)
    yi, yc = table(sh, dh)
    s = src.astype(np.float64)
    hor = np.zeros((sh, dw, cn), np.float64)
    for k in range(4):
        hor = hor + s[:, xi[:, k]] * xc[None, :, k, None]
    out = np.zeros((dh, dw, cn), np.float64)
    for k in range(4):
        out = out + hor[yi[:, k]] * yc[:, k, None, None]
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------- border
def copy_make_border(img: np.ndarray, top: int, bottom: int, left: int, right: int, mode: str | int = "constant") -> np.ndarray:
    """``cv2.copyMakeBorder`` (constant value 0)."""
    m = BORDER_MODES[mode] if isinstance(mode, str) else int(mode)
    h, w = img.shape[:2]
    if m == 0:
        out = np.zeros((h + top + bottom, w + left + right) + img.shape[2:], img.dtype)
        out[top:top + h, left:left + w] = img
        return out
    ys = _border_index(np.arange(-top, h + bottom), h, m)
    xs = _border_index(np.arange(-left, w + right), w, m)
    return img[ys][:, xs]


def plan(h: int, w: int, size: tuple[int, int]):
    """(new_w, new_h, unscale, [top, bottom, left, right], interpolation) of utils.py:317-331."""
    interp = "area" if max(h, w) > max(size) else "cubic"
    rw, rh = size[0] / w, size[1] / h
    if rw < rh:
        nw, nh, unscale = size[0], int(h * rw), rw
        pad = [(size[1] - nh) // 2, (size[1] - nh + 1) // 2, 0, 0]
    else:
        nw, nh, unscale = int(w * rh), size[1], rh
        pad = [0, 0, (size[0] - nw) // 2, (size[0] - nw + 1) // 2]
    return nw, nh, unscale, pad, interp


def as_batch(images, size=512, padding_mode: str = "constant", cubic: str = "fixed"):
    """``as_batch`` (utils.py:273-342): (batch u8 [N,H,W,3], unscales f64 [N], paddings i64 [N,4]).  ``cubic``: "fixed" =
    OpenCV's own INTER_CUBIC code (bit-exact vs cv2 with IPP off), "float" = the IPP build's (see resize_cubic_float)."""
    size = (size, size) if isinstance(size, int) else tuple(size)
    batch, unscales, paddings = [], [], []
    for img in images:
        h, w = img.shape[:2]
        nw, nh, unscale, pad, interp = plan(h, w, size)
        if (nw, nh) == (w, h):
            res = img.copy()
        else:
            res = resize_area(img, nw, nh) if interp == "area" else (resize_cubic if cubic == "fixed" else resize_cubic_float)(img, nw, nh)
        batch.append(copy_make_border(res, *pad, mode=padding_mode))
        unscales.append(unscale)
        paddings.append(pad)
    return np.stack(batch), np.array(unscales, dtype=np.float64), np.array(paddings, dtype=np.int64)
