"""GPU parity tests: every stage of the hot path, called through the C ABI (ctypes), against the CPU oracle
and the reference-generated golden vectors.  Bars (BASELINE.json north_star): bit-exact NMS indices / anchor
selection / warped pixels given identical inputs / parsing argmax given identical logits; <= 1e-3 max-abs on
float32 landmarks, head outputs, logits and SR output."""
import zlib

import numpy as np
import pytest
import torch

from face_crop_plus_b200 import synth
from face_crop_plus_b200.landmarks import landmarks_target

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
CLASS_BIAS = 4.0          # the bias the golden vectors were generated with (tests/golden/meta.npz)
TOL = 1e-3                # float32 tolerance stated by north_star


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


@pytest.fixture(scope="module")
def ctx():
    from face_crop_plus_b200._abi import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def det_ctx(ctx):
    from face_crop_plus_b200 import _abi
    ctx.load_state_dict(_abi.MODEL_RETINAFACE, synth.make_state_dict("retinaface", 0, class_bias=CLASS_BIAS))
    return ctx


@pytest.fixture(scope="module")
def par_ctx(ctx):
    from face_crop_plus_b200 import _abi
    ctx.load_state_dict(_abi.MODEL_BISENET, synth.make_state_dict("bisenet", 0))
    return ctx


@pytest.fixture(scope="module")
def enh_ctx(ctx):
    from face_crop_plus_b200 import _abi
    ctx.load_state_dict(_abi.MODEL_RRDBNET, synth.make_state_dict("rrdbnet", 0))
    return ctx


# --------------------------------------------------------------------------------------------------- conv kernel
CONV_CASES = [  # (n, h, w, cin, cout, k, stride, pad, act, residual)
    (2, 17, 23, 64, 64, 3, 1, 1, "relu", False),
    (1, 32, 32, 256, 128, 1, 1, 0, "none", True),
    (2, 33, 31, 128, 256, 3, 2, 1, "relu", True),
    (1, 16, 16, 512, 1024, 1, 2, 0, "none", False),
    (1, 20, 24, 96, 32, 3, 1, 1, "lrelu", False),
    (1, 64, 64, 256, 19, 1, 1, 0, "none", False),
    (3, 9, 9, 32, 32, 1, 1, 0, "sigmoid", False),
    (1, 40, 40, 192, 64, 3, 1, 1, "none", False),
    (4, 64, 64, 64, 256, 1, 1, 0, "relu", True),      # 256 tiles > 148 CTAs: slab / residual reuse across a CTA's tiles
    (2, 30, 50, 64, 48, 3, 1, 1, "relu", True),       # Cout = 32 + 16: second TMA chunk clipped, residual zero-filled
    (1, 96, 96, 32, 160, 3, 2, 1, "lrelu", False),    # general epilogue form, stride-2 parity views, cout_pad 160 -> BN 32
    (1, 24, 40, 160, 32, 3, 1, 1, "lrelu", False),    # f16x3: 2.5 K-blocks per tap (trailing 32-channel half block)
    (2, 48, 48, 128, 128, 3, 1, 1, "relu", False),    # BN = 128, 18 K-blocks of 64 channels
    (1, 24, 40, 32, 96, 3, 1, 1, "lrelu", True),      # f16x3: K-blocks pair two taps (4.5 per tile); Cout 96 -> one padded 128-wide tile
    (2, 20, 20, 96, 64, 1, 1, 0, "relu", False),      # 1x1 with three 32-channel units: 1.5 K-blocks
    (1, 31, 45, 32, 32, 3, 2, 1, "none", False),      # unit pairs across taps with stride-2 parity views
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("impl", [0, 1, 2])
def test_conv2d_matches_torch(ctx, case, impl):
    n, h, w, cin, cout, k, stride, pad, act, use_res = case
    g = torch.Generator().manual_seed(hash(case) & 0xFFFF)
    x = torch.randn((n, cin, h, w), generator=g)
    wt = torch.randn((cout, cin, k, k), generator=g) * (2.0 / (cin * k * k)) ** 0.5
    scale = 1 + 0.1 * torch.randn(cout, generator=g)
    shift = 0.1 * torch.randn(cout, generator=g)
    ref = torch.nn.functional.conv2d(x.double(), wt.double(), None, stride, pad) * scale.double().view(1, -1, 1, 1) \
        + shift.double().view(1, -1, 1, 1)
    res = torch.randn(ref.shape, generator=g) if use_res else None
    if use_res:
        ref = ref + res.double()
    ref = {"relu": torch.relu, "none": lambda t: t, "lrelu": lambda t: torch.nn.functional.leaky_relu(t, 0.2),
           "sigmoid": torch.sigmoid}[act](ref)
    got = ctx.conv2d(x.permute(0, 2, 3, 1).contiguous().numpy(), wt.numpy(), stride, pad, scale.numpy(), shift.numpy(),
                     None if res is None else res.permute(0, 2, 3, 1).contiguous().numpy(), act, 0.2, impl)
    np.testing.assert_allclose(got, ref.permute(0, 2, 3, 1).numpy(), atol=2e-5, rtol=1e-5)


def test_conv2d_tensor_core_accuracy_large_k(ctx):
    """Both tensor-core split schemes (3xTF32, block-scaled 3xFP16) must be as accurate as the fp32 CUDA-core kernel
    (K = 9*512 = 4608, fp64 reference)."""
    g = torch.Generator().manual_seed(7)
    n, h, w, cin, cout = 2, 32, 32, 512, 256
    x = torch.randn((n, cin, h, w), generator=g).abs() + 0.5           # all-positive inputs: worst case for biased rounding
    wt = torch.randn((cout, cin, 3, 3), generator=g) * (2.0 / (cin * 9)) ** 0.5 + 0.01
    ref = torch.nn.functional.conv2d(x.double(), wt.double(), None, 1, 1).permute(0, 2, 3, 1).numpy()
    xin = x.permute(0, 2, 3, 1).contiguous().numpy()
    err = {}
    for impl in (0, 1, 2):
        got = ctx.conv2d(xin, wt.numpy(), 1, 1, impl=impl)
        err[impl] = np.abs(got - ref).max() / np.abs(ref).max()
    print("relative max error  cuda-core fp32: %.3e   tcgen05 3xTF32: %.3e   tcgen05 3xFP16 block-scaled: %.3e" % (err[0], err[1], err[2]))
    assert err[1] < 5e-6 and err[1] < 8 * max(err[0], 1e-7)
    assert err[2] < 5e-6 and err[2] < 8 * max(err[0], 1e-7)


def test_conv2d_fast_mode_accuracy_class(ctx):
    """impl 3 (opt-in, outside the parity bar): the block-scaled FP16 kernel without its correction terms.  Its error must
    sit in the single-pass class - 11 significant bits per operand, ~2^-11 relative to sum |a||w| at worst and far less
    after averaging over K - and well ABOVE the 3-term modes (i.e. the switch really does drop the correction terms)."""
    g = torch.Generator().manual_seed(13)
    n, h, w, cin, cout = 2, 32, 32, 256, 128
    x = torch.randn((n, cin, h, w), generator=g)
    wt = torch.randn((cout, cin, 3, 3), generator=g) * (2.0 / (cin * 9)) ** 0.5
    ref = torch.nn.functional.conv2d(x.double(), wt.double(), None, 1, 1).permute(0, 2, 3, 1).numpy()
    den = torch.nn.functional.conv2d(x.double().abs(), wt.double().abs(), None, 1, 1).permute(0, 2, 3, 1).numpy()
    xin = x.permute(0, 2, 3, 1).contiguous().numpy()
    err = {}
    for impl in (2, 3):
        got = ctx.conv2d(xin, wt.numpy(), 1, 1, impl=impl)
        err[impl] = float(np.max(np.abs(got - ref) / den))
    print("max |err| / sum|a||w|   3xFP16: %.3e   1xFP16 fast mode: %.3e" % (err[2], err[3]))
    assert err[3] < 2.0 ** -11
    assert err[3] > 20 * err[2]


def test_conv2d_cta_pair_variant_matches(ctx):
    """Opt-in cta_group::2 variant of the wide f16 tiles (FCP_TC_PAIR; measured slower, kept as a measured experiment):
    two CTAs of a cluster share one 256-row MMA and half a weight tile each.  Same results as one CTA per tile, bit for
    bit (the arithmetic per output element is identical), also with an odd number of pixel tiles (phantom tile)."""
    import os
    g = torch.Generator().manual_seed(21)
    for (n, h, w, cin, cout, k, pad) in [(3, 17, 15, 128, 256, 3, 1), (2, 32, 32, 256, 128, 1, 0)]:
        x = torch.randn((n, h, w, cin), generator=g).numpy()
        wt = (torch.randn((cout, cin, k, k), generator=g) * (2.0 / (cin * k * k)) ** 0.5).numpy()
        res = torch.randn((n, h, w, cout), generator=g).numpy()
        one = ctx.conv2d(x, wt, 1, pad, None, None, res, "relu", 0.0, 2)
        os.environ["FCP_TC_PAIR"] = "2"
        try:
            two = ctx.conv2d(x, wt, 1, pad, None, None, res, "relu", 0.0, 2)
        finally:
            del os.environ["FCP_TC_PAIR"]
        assert np.array_equal(one, two)


def test_conv2d_f16x3_dynamic_range(ctx):
    """The block scaling of the 3xFP16 mode: pixel rows spanning 2^-40 .. 2^40 (far outside fp16's 2^-24 .. 2^16), output
    channels whose weights span 2^-12 .. 1, exact zeros, and rows of zeros.  Error is measured per output element against
    sum |a||w| (the scale of its rounding noise) and must stay at the fp32 level everywhere - and not exceed the fp32
    CUDA-core kernel's."""
    g = torch.Generator().manual_seed(11)
    n, h, w, cin, cout = 1, 32, 32, 192, 64
    x = torch.randn((n, cin, h, w), generator=g)
    row_exp = torch.randint(-40, 41, (n, 1, h, w), generator=g).float()
    x = x * torch.exp2(row_exp)
    x[:, :, 3, :] = 0.0                                                  # rows of zeros
    x[:, ::7] = 0.0                                                      # exact zeros inside every row
    x[:, 5:9] *= 2.0 ** -14                                              # channels far below their row maximum
    wt = torch.randn((cout, cin, 3, 3), generator=g) * 0.05
    wt = wt * torch.exp2(-torch.randint(0, 13, (cout, 1, 1, 1), generator=g).float())
    ref = torch.nn.functional.conv2d(x.double(), wt.double(), None, 1, 1).permute(0, 2, 3, 1).numpy()
    den = torch.nn.functional.conv2d(x.double().abs(), wt.double().abs(), None, 1, 1).permute(0, 2, 3, 1).numpy()
    xin = x.permute(0, 2, 3, 1).contiguous().numpy()
    err = {}
    for impl in (0, 1, 2):
        got = ctx.conv2d(xin, wt.numpy(), 1, 1, impl=impl)
        assert np.isfinite(got).all()
        err[impl] = float(np.max(np.abs(got - ref) / np.maximum(den, 1e-300)))
    print("max |err| / sum|a||w|   cuda-core fp32: %.3e   3xTF32: %.3e   3xFP16 block-scaled: %.3e" % (err[0], err[1], err[2]))
    assert err[2] < 4e-7 and err[2] <= 2 * max(err[0], err[1])


# ---------------------------------------------------------------------------------------------------------- align
def test_align_bit_exact_vs_reference_crops(ctx, golden):
    from oracle import align
    g = golden["align"]
    imgs = synth.make_images(2, 300, 260, seed=int(g["images_seed"]))
    lms, idx, pads = g["landmarks"], g["indices"], g["paddings"]
    tgt = landmarks_target((256, 256), 0.65)
    for mode in align.BORDER_MODES:
        for skew in (False, True):
            crops, mats, valid = ctx.align(imgs, pads, idx, lms, tgt, (256, 256), mode, skew)
            assert valid.all()
            np.testing.assert_allclose(mats, g["matrices_affine" if skew else "matrices_partial"], atol=1e-8, rtol=0)
            assert [crc(c) for c in crops] == g[f"crc_{mode}_{int(skew)}"].tolist(), (mode, skew)
    crops, _, _ = ctx.align(imgs, None, idx, lms, landmarks_target((112, 160), 0.8), (112, 160))
    assert np.array_equal(crops, g["crops_112x160"])


def test_align_random_sizes_list_and_degenerate(ctx):
    from oracle import align
    rng = np.random.default_rng(5)
    tgt = landmarks_target((256, 256), 0.65)
    imgs, lms, idx = [], [], []
    for i in range(6):
        H, W = int(rng.integers(8, 500)), int(rng.integers(8, 500))
        imgs.append(rng.integers(0, 256, (H, W, 3), dtype=np.uint8))
        lms.append(synth.make_landmarks(1, max(min(H, W), 16), seed=70 + i)[0])
        idx.append(i)
    lms.append(np.full((5, 2), 3.0, np.float32))      # coincident points: no transform (cropper.py:529-531)
    idx.append(0)
    lms = np.stack(lms)
    for mode in align.BORDER_MODES:
        crops, mats, valid = ctx.align(imgs, None, idx, lms, tgt, (256, 256), mode)
        ref, rmats, rvalid = align.crop_align(imgs, None, idx, lms, tgt, (256, 256), mode)
        assert valid.tolist() == rvalid.tolist() == [True] * 6 + [False]
        assert np.array_equal(crops[valid], ref), mode
        np.testing.assert_allclose(mats[valid], rmats[rvalid], atol=1e-9, rtol=0)
        assert not crops[~valid].any()
    crops, mats, valid = ctx.align(imgs, None, [], np.zeros((0, 5, 2), np.float32), tgt)
    assert crops.shape == (0, 256, 256, 3)


def test_align_full_size_property(ctx):
    # 1024x1024 source, 64 faces: identity-like transforms must reproduce the source window exactly
    img = synth.make_images(1, 1024, 1024, seed=9)
    tgt = landmarks_target((256, 256), 0.65)
    shifts = np.array([[x, y] for x in range(0, 768, 96) for y in range(0, 768, 96)], np.float32)
    lms = tgt[None] + shifts[:, None, :]              # pure integer translation -> crop == img[y:y+256, x:x+256]
    crops, mats, valid = ctx.align(img, None, np.zeros(len(lms), np.int32), lms, tgt)
    assert valid.all()
    for (x, y), c in zip(shifts.astype(int), crops):
        assert np.array_equal(c, img[0, y:y + 256, x:x + 256])


# ------------------------------------------------------------------------------------------------- detection post
@pytest.mark.parametrize("strategy", ["all", "best", "largest"])
def test_detect_post_on_reference_heads(ctx, golden, strategy):
    g = golden["detect"]
    n, h, w, _ = (int(v) for v in g["shape"])
    heads = np.concatenate([g["cls_raw"], g["boxes_raw"], g["ldms_raw"]], -1)
    out = ctx.detect_post(heads, h, w, 0.6, 0.4, strategy)
    assert out["indices"].tolist() == g[f"indices_{strategy}"].tolist()          # bit-exact selection
    np.testing.assert_allclose(out["landmarks"], g[f"landmarks_{strategy}"], atol=1e-4, rtol=0)
    out = ctx.detect_post(heads, h, w, 1.0, 0.4, strategy)                        # nothing passes -> empty
    assert out["landmarks"].shape == (0, 5, 2) and len(out["indices"]) == 0


def test_detect_post_clustered_nms_matches_oracle(ctx):
    # dense clusters of overlapping boxes (many suppressions, > 4096 candidates on one image: exercises both sort paths)
    from oracle import detpost
    rng = np.random.default_rng(3)
    h = w = 512
    a = detpost.priors(h, w).shape[0]
    heads = np.zeros((3, a, 16), np.float32)
    heads[..., 0] = 4.0
    heads[..., 1] = rng.normal(-2, 3, (3, a))
    heads[1, :, 1] += 5                                                           # image 1: ~70% of priors pass
    heads[..., 2:6] = rng.normal(0, 1.5, (3, a, 4))
    heads[..., 6:] = rng.normal(0, 2, (3, a, 10))
    for strategy in ("all", "best", "largest"):
        l, i, anc, b = detpost.detect_post(heads[..., :2], heads[..., 2:6], heads[..., 6:], h, w, 0.6, 0.4, strategy)
        out = ctx.detect_post(heads, h, w, 0.6, 0.4, strategy)
        assert out["indices"].tolist() == i and out["anchors"].tolist() == anc
        np.testing.assert_allclose(out["landmarks"], l, atol=1e-3, rtol=0)
        np.testing.assert_allclose(out["boxes"], b, atol=1e-3, rtol=1e-6)
    assert len(i) == 3


def test_detect_post_capacity_error(ctx, golden):
    from face_crop_plus_b200._abi import FcpError
    g = golden["detect"]
    n, h, w, _ = (int(v) for v in g["shape"])
    heads = np.concatenate([g["cls_raw"], g["boxes_raw"], g["ldms_raw"]], -1)
    out = ctx.detect_post(heads, h, w, 0.6, 0.4, "all", max_faces=4)              # wrapper grows the capacity
    assert len(out["indices"]) == len(g["indices_all"])


# ------------------------------------------------------------------------------------------------ detector network
def test_detect_heads_match_reference(det_ctx, golden):
    g = golden["detect"]
    n, h, w, _ = (int(v) for v in g["shape"])
    imgs = synth.make_images(n, h, w, seed=int(g["images_seed"]))
    heads = det_ctx.detect_heads(imgs)
    ref = np.concatenate([g["cls_raw"], g["boxes_raw"], g["ldms_raw"]], -1)
    assert np.abs(heads - ref).max() < TOL, np.abs(heads - ref).max()


@pytest.mark.parametrize("strategy", ["all", "best", "largest"])
def test_detect_matches_reference_predict(det_ctx, golden, strategy):
    g = golden["detect"]
    n, h, w, _ = (int(v) for v in g["shape"])
    imgs = synth.make_images(n, h, w, seed=int(g["images_seed"]))
    det_ctx.set_micro_batch(2, 32)                     # 3 images -> micro-batches of 2 + 1 (ragged tail)
    out = det_ctx.detect(imgs, 0.6, 0.4, strategy)
    det_ctx.set_micro_batch(16, 32)
    assert out["indices"].tolist() == g[f"indices_{strategy}"].tolist()
    assert np.abs(out["landmarks"] - g[f"landmarks_{strategy}"]).max() < TOL


def test_detect_device_input_and_no_faces(det_ctx):
    imgs = synth.make_images(2, 256, 256, seed=31)
    host = det_ctx.detect(imgs, 0.6, 0.4, "all")
    dev = det_ctx.detect(torch.from_numpy(imgs).cuda(), 0.6, 0.4, "all")
    assert host["indices"].tolist() == dev["indices"].tolist() and np.array_equal(host["landmarks"], dev["landmarks"])
    none = det_ctx.detect(imgs, 1.0, 0.4, "largest")
    assert none["landmarks"].shape == (0, 5, 2)


# --------------------------------------------------------------------------------------------------------- parser
@pytest.mark.parametrize("tag", ["a", "b"])
def test_parse_tail_bit_exact_on_reference_logits(ctx, golden, tag):
    g = golden["parse"]
    n, h, w, _ = (int(v) for v in g[f"{tag}_shape"])
    labels, hist = ctx.parse_tail(g[f"{tag}_logits64"], h, w)
    assert np.array_equal(labels, g[f"{tag}_labels"])
    assert np.array_equal(hist, np.stack([np.bincount(l.ravel(), minlength=19) for l in g[f"{tag}_labels"]]))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_parse_matches_reference(par_ctx, golden, tag):
    g = golden["parse"]
    n, h, w, _ = (int(v) for v in g[f"{tag}_shape"])
    crops = synth.make_images(n, h, w, seed=int(g[f"{tag}_seed"]))
    logits = par_ctx.parse_logits(crops)
    assert np.abs(logits - g[f"{tag}_logits64"]).max() < TOL
    par_ctx.set_micro_batch(16, 2)                     # ragged face micro-batches
    labels, hist = par_ctx.parse(crops)
    par_ctx.set_micro_batch(16, 32)
    diff = labels != g[f"{tag}_labels"]
    if diff.any():   # argmax may flip only where the reference's own top-2 logits are within the float tolerance
        ref_lg = torch.from_numpy(g[f"{tag}_logits64"])
        up = torch.nn.functional.interpolate(torch.nn.functional.interpolate(ref_lg, (512, 512), None, "bilinear", True),
                                             (h, w), mode="nearest")
        top2 = up.topk(2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1]).numpy()
        assert margin[diff].max() < 2 * TOL and diff.mean() < 1e-4
    assert np.array_equal(hist, np.stack([np.bincount(l.ravel(), minlength=19) for l in labels]))


def test_masks_and_grouping(ctx, golden):
    g = golden["parse"]
    labels = g["a_labels"]
    for name, classes in {"eyes_and_eyebrows": [2, 3, 4, 5], "skin": [1], "lips": [11, 12, 13]}.items():
        m = ctx.masks(labels, classes)
        assert np.array_equal(m, (np.isin(labels, classes) * 255).astype(np.uint8))


# ------------------------------------------------------------------------------------------------------- enhancer
def test_enhance_matches_reference(enh_ctx, golden):
    g = golden["enhance"]
    x = torch.from_numpy(synth.make_images(2, 24, 32, seed=int(g["images_seed"]))).permute(0, 3, 1, 2).float().contiguous()
    y = enh_ctx.enhance_forward((x[:1] / 255).numpy())
    assert np.abs(y - g["forward0"]).max() < TOL
    imgs = x.clone().numpy()
    out = enh_ctx.enhance(imgs, gate=[1, 0])
    assert np.array_equal(out[1], x[1].numpy())                                    # gated off: untouched
    d = np.abs(out[0] - g["predict"][0])
    assert d.max() <= 1 and (d > 0).mean() < 0.01                                  # only x.5 rounding ties may flip
    dev = x.clone().cuda()
    enh_ctx.enhance(dev, gate=None)
    assert np.abs(dev.cpu().numpy() - g["predict_all"]).max() <= 1


# ----------------------------------------------------------------------------------------------------- whole path
def test_pipeline_matches_oracle(det_ctx, par_ctx):
    from oracle import pipeline
    imgs = synth.make_images(3, 256, 320, seed=2000)
    pads = np.array([[0, 0, 0, 0], [4, 2, 6, 0], [0, 0, 0, 0]], np.int32)
    det_sd = synth.make_state_dict("retinaface", 0, class_bias=CLASS_BIAS)
    par_sd = synth.make_state_dict("bisenet", 0)
    tgt = landmarks_target((256, 256), 0.65)
    for strategy in ("largest", "all"):
        ref = pipeline.process_batch(imgs, det_sd, par_sd, paddings=pads, strategy=strategy, det_threshold=0.9)
        out = det_ctx.pipeline(imgs, pads, tgt, (256, 256), 0.9, 0.4, strategy)
        assert out["indices"].tolist() == ref["indices"]
        assert np.abs(out["landmarks"] - ref["landmarks"]).max() < TOL
        # landmark noise (<= TOL px) is amplified by the source coordinates (~|M| * 320 px) in the translation column
        np.testing.assert_allclose(out["matrices"], ref["matrices"], atol=5e-3, rtol=0)
        # crops: identical wherever the fixed-point source coordinates agree; landmark noise of ~1e-4 px can move a
        # 1/32-px interpolation bin, so compare pixel values with a small tolerance and require near-total equality
        d = np.abs(out["crops"].astype(int) - ref["crops"].astype(int))
        lab = (out["labels"] != ref["labels"]).mean()
        print(f"pipeline vs oracle ({strategy}): landmarks {np.abs(out['landmarks'] - ref['landmarks']).max():.2e} px, crop px mismatch "
              f"{(d > 0).mean():.3e} (max {d.max()}), label mismatch {lab:.3e}")
        # observed on B200 (round 2): 1.3e-3 .. 1.4e-3 of the crop bytes differ (max 8), 1.0e-4 of the labels; bounds = ~2x observed
        assert (d > 0).mean() < 3e-3 and d.max() <= 16
        assert lab < 3e-4


def test_enhance_forward_64x64_and_linearity_property(enh_ctx):
    """RRDBNet.forward at 64x64 (351 convs, slab concat views, fused upsample path) vs the oracle, plus a size-independent
    property at 128x128: the network is translation-equivariant away from the borders (all-conv, zero padding)."""
    from oracle import nets
    sd = synth.make_state_dict("rrdbnet", 0)
    x = torch.from_numpy(synth.make_images(1, 64, 64, seed=21)).permute(0, 3, 1, 2).float().contiguous() / 255
    ref = nets.rrdbnet_forward(x, sd).numpy()
    got = enh_ctx.enhance_forward(x.numpy())
    assert np.abs(got - ref).max() < TOL
    big = torch.from_numpy(synth.make_images(1, 128, 160, seed=22)).permute(0, 3, 1, 2).float().contiguous() / 255
    full = enh_ctx.enhance_forward(big.numpy())
    crop = enh_ctx.enhance_forward(np.ascontiguousarray(big.numpy()[:, :, 16:112, 24:136]))
    # receptive field of the 351-conv stack is large; compare a centre window where the crop's borders cannot reach
    # only loosely (it decays geometrically with depth thanks to the x0.2 residual scaling)
    a = full[:, :, 4 * 16 + 160:4 * 112 - 160, 4 * 24 + 160:4 * 136 - 160]
    b = crop[:, :, 160:-160, 160:-160]
    assert a.shape == b.shape and np.abs(a - b).max() < 5e-3


# ------------------------------------------------------------------------------------------ full-size (1024x1024)
def test_detect_full_size_vs_oracle(ctx):
    """BASELINE.json's image size: 2 images of 1024x1024 (43 008 priors each) against the oracle."""
    from face_crop_plus_b200 import _abi
    from oracle import pipeline
    sd = synth.make_state_dict("retinaface", 0, class_bias=4.8)       # the bench's candidate density
    ctx.load_state_dict(_abi.MODEL_RETINAFACE, sd)
    imgs = synth.make_images(2, 1024, 1024, seed=1234)
    for strategy in ("all", "largest"):
        lms, idx, anchors, boxes = pipeline.detect(imgs, sd, 0.6, 0.4, strategy)
        out = ctx.detect(imgs, 0.6, 0.4, strategy)
        got = list(zip(out["indices"].tolist(), out["anchors"].tolist()))
        ref = list(zip(idx, anchors))
        assert sorted(got) == sorted(ref) and out["indices"].tolist() == idx       # same priors selected per image
        # within an image faces are ordered by score; 200+ random faces contain scores that differ by < 1e-6 (float32
        # spacing near 1.0 is 6e-8), so the order may differ only between such near-ties
        pos = {k: j for j, k in enumerate(got)}
        worst = max(float(np.abs(out["landmarks"][pos[k]] - lms[j]).max()) for j, k in enumerate(ref))
        ctx.set_conv_impl(0)                               # second witness: the CUDA-core fp32 kernel on the same device
        w0 = ctx.detect(imgs, 0.6, 0.4, strategy)
        ctx.set_conv_impl(2)
        p0 = {k: j for j, k in enumerate(zip(w0["indices"].tolist(), w0["anchors"].tolist()))}
        worst0 = max(float(np.abs(w0["landmarks"][p0[k]] - lms[j]).max()) for j, k in enumerate(ref) if k in p0)
        print(f"detect 1024x1024 ({strategy}): {len(ref)} faces, max landmark err vs oracle: tensor-core {worst:.3e} px, "
              f"cuda-core fp32 {worst0:.3e} px (torch CPU threads {torch.get_num_threads()})")
        for j, k in enumerate(ref):
            if pos[k] != j:
                assert abs(float(out["scores"][pos[k]]) - float(out["scores"][j])) < 1e-5
            assert np.abs(out["landmarks"][pos[k]] - lms[j]).max() < TOL
    assert len(idx) == 2
    ctx.load_state_dict(_abi.MODEL_RETINAFACE, synth.make_state_dict("retinaface", 0, class_bias=CLASS_BIAS))


def test_full_size_batch_consistency_between_conv_kernels(det_ctx, par_ctx):
    """Size-independent property at the bench shape (1024x1024, 24 images, ragged micro-batches): the two tensor-core
    paths (3xFP16 block-scaled, 3xTF32) and the CUDA-core (fp32) convolution path must select the same priors and agree
    to the float tolerance."""
    imgs = synth.make_images(24, 1024, 1024, seed=4000)
    tgt = landmarks_target((256, 256), 0.65)
    res = {}
    for impl in (2, 1, 0):
        det_ctx.set_conv_impl(impl)
        det_ctx.set_micro_batch(16 if impl else 5, 32 if impl else 7)
        res[impl] = det_ctx.pipeline(imgs, None, tgt, (256, 256), 0.6, 0.4, "largest")
    det_ctx.set_conv_impl(2)
    det_ctx.set_micro_batch(16, 32)
    for a, b in ((res[2], res[0]), (res[1], res[0]), (res[2], res[1])):
        assert a["count"] == b["count"] > 0 and a["indices"].tolist() == b["indices"].tolist()
        assert np.abs(a["landmarks"] - b["landmarks"]).max() < TOL
        assert (a["crops"] != b["crops"]).mean() < 0.02 and (a["labels"] != b["labels"]).mean() < 5e-3
        assert np.array_equal(a["hist"].sum(1), np.full(a["count"], 256 * 256))
