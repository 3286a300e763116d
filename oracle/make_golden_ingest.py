"""Generates tests/golden/ingest.npz by running the UNMODIFIED reference ``as_batch`` (utils.py:273-342, imported from
/root/reference/src) on a seeded ragged list of small images.  Build container only (``python -m oracle.make_golden_ingest``).

Two variants are stored per configuration: the reference as this image runs it by default (opencv-python's 8-bit
INTER_CUBIC goes through Intel IPP) and with ``cv2.ipp.setUseIPP(False)`` (OpenCV's own code, the arithmetic
``oracle/ingest.py`` and ``csrc/ingest.cu`` restate).  INTER_AREA and the borders are identical in both.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

# (h, w) of the source images: larger than the target (AREA: fractional, 2x, 3x), smaller (CUBIC), equal (copy)
SHAPES = [(150, 220), (128, 192), (192, 288), (40, 33), (64, 96), (31, 90), (200, 90), (64, 64)]
CONFIGS = [((96, 64), "constant"), ((96, 64), "reflect_101"), ((64, 64), "replicate"), ((80, 112), "wrap"), ((96, 64), "reflect")]


def images(seed: int = 77):
    from face_crop_plus_b200 import synth
    return [np.ascontiguousarray(synth.make_images(1, h, w, seed=seed + i)[0]) for i, (h, w) in enumerate(SHAPES)]


def main():
    import cv2
    sys.modules.setdefault("unidecode", types.ModuleType("unidecode"))
    sys.path.insert(0, "/root/reference/src")
    from face_crop_plus.utils import as_batch
    imgs = images()
    out = {"n_images": np.array(len(imgs))}
    for ci, (size, mode) in enumerate(CONFIGS):
        res = {}
        for tag, ipp in (("ipp", True), ("cv", False)):
            cv2.ipp.setUseIPP(ipp)
            res[tag], unscales, paddings = as_batch(imgs, size, mode)
        out[f"c{ci}_cv_batch"] = res["cv"]
        out[f"c{ci}_ipp_minus_cv"] = (res["ipp"].astype(np.int16) - res["cv"].astype(np.int16)).astype(np.int8)   # sparse, +-1
        out[f"c{ci}_unscales"] = np.asarray(unscales, np.float64)
        out[f"c{ci}_paddings"] = np.asarray(paddings, np.int64)
    cv2.ipp.setUseIPP(True)
    np.savez_compressed(REPO / "tests" / "golden" / "ingest.npz", **out)
    print("wrote tests/golden/ingest.npz", sum(v.nbytes for v in out.values()), "bytes raw")


if __name__ == "__main__":
    main()
