"""Oracle (test infrastructure): numpy restatement of the align step, ``Cropper.crop_align`` (cropper.py:441-552).

The arithmetic lives in OpenCV (un-vendored; opencv-python 4.13.0 in this image), which the
reference calls at cropper.py:515-527 (``estimateAffinePartial2D`` / ``estimateAffine2D`` with
``ransacReprojThreshold=inf``) and cropper.py:542-547 (``warpAffine``, INTER_LINEAR).  The published
algorithm restated here: a float64 least-squares fit, and OpenCV's
fixed-point bilinear remap (AB_BITS=10, INTER_BITS=5, 15-bit weights).  Pinned bit-for-bit against
``cv2`` itself in tests/test_oracle_golden.py and against reference crops in tests/golden/.
"""
from __future__ import annotations

import numpy as np

BORDER_MODES = {"constant": 0, "replicate": 1, "reflect": 2, "wrap": 3, "reflect_101": 4}


def solve_partial(src: np.ndarray, dst: np.ndarray):
    """4-dof similarity LS fit src->dst in float64; returns 2x3 or None (== estimateAffinePartial2D(..., inf)[0])."""
    s, d = src.astype(np.float64), dst.astype(np.float64)
    ms, md = s.mean(0), d.mean(0)
    xs, yd = s - ms, d - md
    den = (xs ** 2).sum()
    if den == 0:
        return None
    a = (xs[:, 0] * yd[:, 0] + xs[:, 1] * yd[:, 1]).sum() / den
    b = (xs[:, 0] * yd[:, 1] - xs[:, 1] * yd[:, 0]).sum() / den
    tx = md[0] - (a * ms[0] - b * ms[1])
    ty = md[1] - (b * ms[0] + a * ms[1])
    return np.array([[a, -b, tx], [b, a, ty]], dtype=np.float64)


def solve_affine(src: np.ndarray, dst: np.ndarray):
    """6-dof affine LS fit in float64 via centred normal equations (== estimateAffine2D(..., inf)[0]); None if degenerate."""
    s, d = src.astype(np.float64), dst.astype(np.float64)
    ms, md = s.mean(0), d.mean(0)
    xs, yd = s - ms, d - md
    sxx, sxy, syy = (xs[:, 0] ** 2).sum(), (xs[:, 0] * xs[:, 1]).sum(), (xs[:, 1] ** 2).sum()
    det = sxx * syy - sxy * sxy
    if det == 0 or not np.isfinite(det) or abs(det) < 1e-12 * max(sxx * syy, 1e-300):
        return None
    m = np.empty((2, 3))
    for r in range(2):
        bx, by = (xs[:, 0] * yd[:, r]).sum(), (xs[:, 1] * yd[:, r]).sum()
        m[r, 0] = (bx * syy - by * sxy) / det
        m[r, 1] = (by * sxx - bx * sxy) / det
        m[r, 2] = md[r] - m[r, 0] * ms[0] - m[r, 1] * ms[1]
    return m


def _border_index(p, n, mode):
    """OpenCV ``borderInterpolate`` for REPLICATE / REFLECT / WRAP / REFLECT_101 (vectorised)."""
    p = p.astype(np.int64)
    if mode == 1:
        return np.clip(p, 0, n - 1)
    if mode in (2, 4):
        if n == 1:
            return np.zeros_like(p)
        delta = 1 if mode == 4 else 0
        for _ in range(64):
            bad = (p < 0) | (p >= n)
            if not bad.any():
                break
            p = np.where(p < 0, -p - 1 + delta, p)
            p = np.where(p >= n, n - 1 - (p - n) - delta, p)
        return p
    if mode == 3:
        return np.mod(p, n)
    raise ValueError(mode)


def warp_affine(img: np.ndarray, M: np.ndarray, out_w: int, out_h: int, border: str | int = "constant") -> np.ndarray:
    """Bit-exact restatement of ``cv2.warpAffine(img, M, (out_w,out_h), flags=INTER_LINEAR, borderMode=...)`` for u8 HxWxC."""
    mode = BORDER_MODES[border] if isinstance(border, str) else int(border)
    H, W = img.shape[:2]
    M = np.asarray(M, dtype=np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    i00, i11 = M[1, 1] * D, M[0, 0] * D
    i01, i10 = M[0, 1] * (-D), M[1, 0] * (-D)
    i02 = -i00 * M[0, 2] - i01 * M[1, 2]
    i12 = -i10 * M[0, 2] - i11 * M[1, 2]
    sat = lambda v: np.clip(np.rint(v), -2147483648, 2147483647).astype(np.int64)
    x = np.arange(out_w, dtype=np.float64)
    y = np.arange(out_h, dtype=np.float64)
    adelta, bdelta = sat(i00 * x * 1024), sat(i10 * x * 1024)
    X0 = sat((i01 * y + i02) * 1024) + 16
    Y0 = sat((i11 * y + i12) * 1024) + 16
    X = (X0[:, None] + adelta[None, :]).astype(np.int32).astype(np.int64) >> 5   # int32 wrap like the C code
    Y = (Y0[:, None] + bdelta[None, :]).astype(np.int32).astype(np.int64) >> 5
    sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)     # saturate_cast<short>
    fx, fy = X & 31, Y & 31
    w = [32 * (32 - fx) * (32 - fy), 32 * fx * (32 - fy), 32 * (32 - fx) * fy, 32 * fx * fy]
    src = img.reshape(H, W, -1).astype(np.int64)
    acc = np.zeros((out_h, out_w, src.shape[2]), dtype=np.int64)
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        yy, xx = sy + dy, sx + dx
        if mode == 0:
            ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            p = src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)] * ok[..., None]
        else:
            p = src[_border_index(yy, H, mode), _border_index(xx, W, mode)]
        acc += w[k][..., None] * p
    out = ((acc + 16384) >> 15).astype(np.uint8)
    return out.reshape(out_h, out_w, *img.shape[2:])


def crop_align(images, padding, indices, landmarks_source, landmarks_target, output_size=(256, 256),
               border="constant", allow_skew=False):
    """``Cropper.crop_align`` (cropper.py:441-552).  Returns (crops u8[F',h,w,3], matrices f64[F,2,3], valid bool[F])."""
    solve = solve_affine if allow_skew else solve_partial
    crops, mats, valid = [], [], []
    for li, ii in enumerate(indices):
        M = solve(landmarks_source[li], landmarks_target)
        valid.append(M is not None)
        mats.append(M if M is not None else np.full((2, 3), np.nan))
        if M is None:
            continue                                                   # cropper.py:529-531
        img = images[ii]
        if padding is not None:
            t, b, l, r = (int(v) for v in padding[ii])
            img = img[t:img.shape[0] - b, l:img.shape[1] - r]          # cropper.py:536-539
        crops.append(warp_affine(img, M, output_size[0], output_size[1], border))
    crops = np.stack(crops) if crops else np.array([])
    return crops, np.array(mats, dtype=np.float64).reshape(-1, 2, 3), np.array(valid, dtype=bool)
