"""Ingest oracle (oracle/ingest.py) pinned against cv2 itself and against the reference's ``as_batch`` outputs
(tests/golden/ingest.npz, made by oracle/make_golden_ingest.py from /root/reference/src/face_crop_plus/utils.py:273-342)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import ingest
from oracle.make_golden_ingest import CONFIGS, images

GOLD = np.load(__file__.rsplit("/", 1)[0] + "/golden/ingest.npz")


@pytest.fixture(autouse=True)
def _restore_ipp():
    yield
    cv2.ipp.setUseIPP(True)


@pytest.mark.parametrize("ipp", [True, False])
@pytest.mark.parametrize("src,dst", [((1000, 1500), (682, 1024)), ((512, 512), (256, 256)), ((768, 1536), (256, 512)),
                                     ((700, 1025), (699, 1024)), ((333, 130), (256, 99)), ((300, 1100), (279, 1024))])
def test_area_bit_exact_vs_cv2(src, dst, ipp):
    cv2.ipp.setUseIPP(ipp)
    rng = np.random.default_rng(src[0] * 7 + dst[1])
    img = rng.integers(0, 256, (*src, 3), dtype=np.uint8)
    ref = cv2.resize(img, (dst[1], dst[0]), interpolation=cv2.INTER_AREA)
    assert np.array_equal(ingest.resize_area(img, dst[1], dst[0]), ref)


@pytest.mark.parametrize("src,dst", [((29, 37), (50, 64)), ((80, 100), (205, 257)), ((281, 500), (575, 1024)), ((600, 600), (256, 437)),
                                     ((33, 33), (300, 301)), ((9, 5), (37, 13))])
def test_cubic_bit_exact_vs_opencv_own_code(src, dst):
    cv2.ipp.setUseIPP(False)
    rng = np.random.default_rng(src[0] * 11 + dst[1])
    img = rng.integers(0, 256, (*src, 3), dtype=np.uint8)
    ref = cv2.resize(img, (dst[1], dst[0]), interpolation=cv2.INTER_CUBIC)
    assert np.array_equal(ingest.resize_cubic(img, dst[1], dst[0]), ref)


def test_cubic_simd_tail_split_is_pinned():
    """The float32 SIMD body / integer tail split (8 values) matters: narrow, tall images hit tail values where the two differ."""
    cv2.ipp.setUseIPP(False)
    rng = np.random.default_rng(7)
    for _ in range(300):
        sw, sh = int(rng.integers(3, 12)), int(rng.integers(20, 60))
        dw, dh = int(rng.integers(sw, 3 * sw)), int(rng.integers(2 * sh, 5 * sh))
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        assert np.array_equal(ingest.resize_cubic(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_CUBIC))


def test_cubic_vs_ipp_build_differs_by_at_most_one():
    cv2.ipp.setUseIPP(True)
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (120, 90, 3), dtype=np.uint8)
    ref = cv2.resize(img, (200, 267), interpolation=cv2.INTER_CUBIC)
    d = np.abs(ingest.resize_cubic(img, 200, 267).astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 0.08


def test_cubic_float_vs_ipp_build():
    """The float restatement of INTER_CUBIC against cv2 as this image runs it (IPP on): enlargements and the reductions a
    non-square target can ask for; at most one grey level off, in fewer than 5e-5 of the bytes (observed 7e-6)."""
    cv2.ipp.setUseIPP(True)
    rng = np.random.default_rng(8)
    bad = tot = 0
    for _ in range(24):
        sh, sw = int(rng.integers(8, 140)), int(rng.integers(8, 140))
        f = float(rng.uniform(0.7, 3.2))
        dw, dh = max(4, int(sw * f)), max(4, int(sh * f))
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_CUBIC)
        d = np.abs(ingest.resize_cubic_float(img, dw, dh).astype(int) - ref.astype(int))
        assert d.max() <= 1, (sh, sw, dw, dh)
        bad += int((d > 0).sum()); tot += d.size
    assert bad / tot < 5e-5, bad / tot


@pytest.mark.parametrize("mode", list(ingest.BORDER_MODES))
def test_copy_make_border_vs_cv2(mode):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (7, 5, 3), dtype=np.uint8)
    for pads in [(0, 0, 3, 4), (2, 3, 0, 0), (9, 8, 6, 7), (0, 0, 0, 0)]:
        ref = cv2.copyMakeBorder(img, *pads, borderType=getattr(cv2, "BORDER_" + mode.upper()))
        assert np.array_equal(ingest.copy_make_border(img, *pads, mode=mode), ref)


@pytest.mark.parametrize("ci", range(len(CONFIGS)))
def test_as_batch_vs_reference_golden(ci):
    size, mode = CONFIGS[ci]
    batch, unscales, paddings = ingest.as_batch(images(), size, mode)
    assert np.array_equal(batch, GOLD[f"c{ci}_cv_batch"])                       # reference with OpenCV's own resize code
    assert np.array_equal(unscales, GOLD[f"c{ci}_unscales"]) and np.array_equal(paddings, GOLD[f"c{ci}_paddings"])
    d = GOLD[f"c{ci}_ipp_minus_cv"]                                             # reference as this image runs it (IPP cubic)
    assert np.abs(d).max() <= 1


def test_random_shapes_property():
    """Randomised pin (seeded): any shrink through INTER_AREA and any resize through INTER_CUBIC equals cv2, plus the as_batch
    plan (sizes, paddings, interpolation choice) against cv2 called the way utils.py:317-335 calls it."""
    cv2.ipp.setUseIPP(False)
    rng = np.random.default_rng(2024)
    for _ in range(120):
        sh, sw = int(rng.integers(2, 90)), int(rng.integers(2, 90))
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        dh, dw = int(rng.integers(1, sh + 1)), int(rng.integers(1, sw + 1))
        assert np.array_equal(ingest.resize_area(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_AREA)), (sh, sw, dh, dw)
        dh, dw = int(rng.integers(1, 150)), int(rng.integers(1, 150))
        assert np.array_equal(ingest.resize_cubic(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_CUBIC)), (sh, sw, dh, dw)
    for _ in range(40):
        imgs = [rng.integers(0, 256, (int(rng.integers(8, 120)), int(rng.integers(8, 120)), 3), dtype=np.uint8) for _ in range(3)]
        size = (int(rng.integers(16, 100)), int(rng.integers(16, 100)))
        mode = list(ingest.BORDER_MODES)[int(rng.integers(0, 5))]
        batch, _, pads = ingest.as_batch(imgs, size, mode)
        for i, im in enumerate(imgs):
            h, w = im.shape[:2]
            nw, nh, _, pad, interp = ingest.plan(h, w, size)
            ref = cv2.resize(im, (nw, nh), interpolation=cv2.INTER_AREA if interp == "area" else cv2.INTER_CUBIC)
            ref = cv2.copyMakeBorder(ref, *pad, borderType=getattr(cv2, "BORDER_" + mode.upper()))
            assert np.array_equal(batch[i], ref) and pads[i].tolist() == pad
