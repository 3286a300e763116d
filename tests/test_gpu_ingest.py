"""fcp_as_batch (csrc/ingest.cu) vs the ingest oracle (bit-exact: integer / byte work) and vs the reference's ``as_batch``
outputs of tests/golden/ingest.npz; full-size cases through size-independent properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from face_crop_plus_b200 import _abi
    c = _abi.Context(0)
    yield c
    c.close()


def _rand_images(shapes, seed):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]


@pytest.mark.parametrize("mode", ["constant", "replicate", "reflect", "wrap", "reflect_101"])
def test_as_batch_matches_oracle_all_paths(ctx, mode):
    from oracle import ingest
    # fractional AREA, 2x2, 3x3 integer AREA, CUBIC enlargement (narrow: SIMD tail), copy, extreme aspect ratios
    shapes = [(150, 220), (128, 192), (192, 288), (40, 33), (64, 96), (31, 90), (200, 90), (9, 5), (97, 300), (64, 64)]
    imgs = _rand_images(shapes, 11)
    for size in [(96, 64), (64, 64), (80, 112)]:
        for cubic in ("float", "fixed"):                       # both INTER_CUBIC arithmetics, each bit-exact vs its restatement
            ctx.set_cubic_mode(cubic == "float")
            got = ctx.as_batch(imgs, size, mode)
            ref = ingest.as_batch(imgs, size, mode, cubic=cubic)
            assert np.array_equal(got[0], ref[0]), f"{size} {mode} {cubic}: {(got[0] != ref[0]).sum()} bytes differ"
            assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
    ctx.set_cubic_mode(True)


def test_as_batch_vs_reference_golden(ctx):
    from oracle.make_golden_ingest import CONFIGS, images
    gold = np.load(__file__.rsplit("/", 1)[0] + "/golden/ingest.npz")
    imgs = images()
    bad = tot = 0
    for ci, (size, mode) in enumerate(CONFIGS):
        ctx.set_cubic_mode(False)
        batch, unscales, paddings = ctx.as_batch(imgs, size, mode)
        assert np.array_equal(batch, gold[f"c{ci}_cv_batch"])                   # OpenCV's own arithmetic: bit-exact
        assert np.array_equal(unscales, gold[f"c{ci}_unscales"]) and np.array_equal(paddings, gold[f"c{ci}_paddings"])
        ctx.set_cubic_mode(True)                                                # default: the arithmetic of the IPP build
        batch, _, _ = ctx.as_batch(imgs, size, mode)
        ipp = gold[f"c{ci}_cv_batch"].astype(np.int16) + gold[f"c{ci}_ipp_minus_cv"]      # the reference as this image runs it
        d = np.abs(batch.astype(np.int16) - ipp)
        assert d.max() <= 1
        bad += int((d > 0).sum()); tot += d.size
    print(f"as_batch (float cubic) vs the reference's own outputs (IPP build of cv2): {bad} of {tot} bytes differ by one grey level")
    assert bad / tot < 5e-5


def test_as_batch_device_resident_in_and_out(ctx):
    import torch
    from oracle import ingest
    imgs = _rand_images([(300, 500), (700, 400), (256, 256)], 5)
    dev = [torch.from_numpy(im).cuda() for im in imgs]
    out = torch.empty((3, 256, 256, 3), dtype=torch.uint8, device="cuda")
    got, _, pads = ctx.as_batch(dev, 256, "constant", out=out)
    ref = ingest.as_batch(imgs, 256, "constant", cubic="float")
    assert got is out and np.array_equal(out.cpu().numpy(), ref[0]) and np.array_equal(pads, ref[2])


def test_full_size_properties(ctx):
    """BASELINE sizes: 1024x1024 is the identity (utils.py:320-334 degenerates to a copy), 2048x2048 -> 1024x1024 is the
    rounded 2x2 block mean, 3072 -> 1024 the cvRound of the 3x3 mean; mixed-resolution list keeps aspect + centring."""
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (1024, 1024, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (2048, 2048, 3), dtype=np.uint8)
    c = rng.integers(0, 256, (3072, 3072, 3), dtype=np.uint8)
    d = rng.integers(0, 256, (1648, 2464, 3), dtype=np.uint8)
    batch, unscales, pads = ctx.as_batch([a, b, c, d], 1024)
    assert np.array_equal(batch[0], a)
    blk = b.reshape(1024, 2, 1024, 2, 3).astype(np.int32).sum((1, 3))
    assert np.array_equal(batch[1], ((blk + 2) >> 2).astype(np.uint8))
    blk3 = c.reshape(1024, 3, 1024, 3, 3).astype(np.int32).sum((1, 3)).astype(np.float32) * (np.float32(1) / np.float32(9))
    assert np.array_equal(batch[2], np.rint(blk3).astype(np.uint8))
    nh = int(1648 * (1024 / 2464))
    assert pads[3].tolist() == [(1024 - nh) // 2, (1024 - nh + 1) // 2, 0, 0] and unscales[3] == 1024 / 2464
    assert not batch[3, :pads[3][0]].any() and not batch[3, pads[3][0] + nh:].any()
    inner = batch[3, pads[3][0]:pads[3][0] + nh].astype(np.float64)
    assert abs(inner.mean() - d.mean()) < 0.5                                   # area averaging preserves the mean


def test_as_batch_geometry_matches_host_plan(ctx):
    from face_crop_plus_b200 import utils
    shapes = [(100, 200), (300, 150), (64, 64), (77, 1000)]
    batch, unscales, pads = utils.as_batch(_rand_images(shapes, 2), (128, 96), ctx=ctx)
    assert batch.shape == (4, 96, 128, 3)
    for (h, w), u, p in zip(shapes, unscales, pads):
        _, _, unscale, pad = utils.batch_plan(h, w, (128, 96))
        assert u == unscale and p.tolist() == pad


def test_as_batch_errors(ctx):
    from face_crop_plus_b200 import _abi
    with pytest.raises(ValueError):
        ctx.as_batch([np.zeros((4, 4), np.uint8)], 8)
    with pytest.raises(_abi.FcpError):
        ctx.as_batch([np.zeros((1, 5000, 3), np.uint8)], 64)                    # collapses to zero rows, like cv2.resize raising
    batch, unscales, pads = ctx.as_batch([], 64)
    assert batch.shape == (0, 64, 64, 3)
