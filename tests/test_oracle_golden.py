"""CPU: pins the oracle restatement against outputs of the unmodified reference (tests/golden, made by
oracle/make_golden.py) and against cv2 itself where it is importable."""
import zlib

import numpy as np
import pytest
import torch

from face_crop_plus_b200 import synth
from face_crop_plus_b200.landmarks import landmarks_target
from oracle import align, detpost, enhance, nets, parse

torch.set_grad_enabled(False)
CLASS_BIAS = 4.0


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def test_weight_digests(golden):
    m = golden["meta"]
    assert float(m["class_bias"]) == CLASS_BIAS
    for model in synth.SPECS:
        kw = {"class_bias": CLASS_BIAS} if model == "retinaface" else {}
        assert synth.state_dict_digest(synth.make_state_dict(model, 0, **kw)) == str(m[f"digest_{model}"])


@pytest.fixture(scope="module")
def det_oracle(golden):
    g = golden["detect"]
    imgs = synth.make_images(3, 320, 384, seed=int(g["images_seed"]))
    sd = synth.make_state_dict("retinaface", 0, class_bias=CLASS_BIAS)
    x = torch.from_numpy(imgs).permute(0, 3, 1, 2).float()
    cls, box, ldm = nets.retinaface_heads_raw(nets.retinaface_preprocess(x), sd)
    return cls.numpy(), box.numpy(), ldm.numpy()


def test_retinaface_forward_matches_reference(golden, det_oracle):
    g = golden["detect"]
    cls, box, ldm = det_oracle
    # same torch ops on the same host class: expect (near) bit equality; tolerance covers thread-count effects
    np.testing.assert_allclose(cls, g["cls_raw"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(box, g["boxes_raw"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(ldm, g["ldms_raw"], atol=2e-4, rtol=0)
    np.testing.assert_allclose(detpost.softmax_face_score(cls), g["scores"], atol=2e-6, rtol=0)


def test_priors_match_reference_formula():
    # _layers.py:49-62 evaluated literally (python doubles -> float32) on a non-power-of-two size
    from itertools import product
    from math import ceil
    h, w = 250, 330
    anchors = []
    for k, step in enumerate((8, 16, 32)):
        for i, j in product(range(ceil(h / step)), range(ceil(w / step))):
            for ms in detpost.MIN_SIZES[k]:
                anchors.append(((j + 0.5) * step / w, (i + 0.5) * step / h, ms / w, ms / h))
    assert np.array_equal(detpost.priors(h, w), np.array(anchors, dtype=np.float32))
    assert detpost.priors(1024, 1024).shape == (43008, 4)


@pytest.mark.parametrize("strategy", ["all", "best", "largest"])
def test_detect_post_matches_reference_predict(golden, strategy):
    # post-processing pinned on the reference's own head outputs -> exact index parity, landmarks to rounding
    g = golden["detect"]
    _, h, w, _ = g["shape"]
    l, i, _, _ = detpost.detect_post(g["cls_raw"], g["boxes_raw"], g["ldms_raw"], int(h), int(w), 0.6, 0.4, strategy)
    assert i == g[f"indices_{strategy}"].tolist()
    np.testing.assert_allclose(l, g[f"landmarks_{strategy}"], atol=1e-4, rtol=0)
    assert len(i) > (20 if strategy == "all" else 2)


def test_detect_post_edge_cases(golden):
    g = golden["detect"]
    with pytest.raises(ValueError):
        detpost.detect_post(g["cls_raw"], g["boxes_raw"], g["ldms_raw"], 320, 384, strategy="bogus")
    l, i, a, b = detpost.detect_post(g["cls_raw"], g["boxes_raw"], g["ldms_raw"], 320, 384, vis_threshold=1.0)
    assert l.shape == (0, 5, 2) and i == [] and a == []


def test_align_matches_reference(golden):
    g = golden["align"]
    imgs = synth.make_images(2, 300, 260, seed=int(g["images_seed"]))
    lms, idx, pads = g["landmarks"], g["indices"].tolist(), g["paddings"]
    tgt = landmarks_target((256, 256), 0.65)
    assert np.array_equal(tgt, g["target_256_0.65"])
    assert np.array_equal(landmarks_target((112, 160), 0.8), g["target_112x160_0.8"])
    np.testing.assert_allclose(np.stack([align.solve_partial(l, tgt) for l in lms]), g["matrices_partial"], atol=1e-9)
    np.testing.assert_allclose(np.stack([align.solve_affine(l, tgt) for l in lms]), g["matrices_affine"], atol=1e-8)
    for mode in align.BORDER_MODES:
        for skew in (False, True):
            crops, mats, valid = align.crop_align(imgs, pads, idx, lms, tgt, (256, 256), mode, skew)
            assert valid.all() and [crc(c) for c in crops] == g[f"crc_{mode}_{int(skew)}"].tolist(), (mode, skew)
            if mode == "constant" and not skew:
                assert np.array_equal(crops, g["crops_constant_0"])
    crops, _, _ = align.crop_align(imgs, None, idx, lms, landmarks_target((112, 160), 0.8), (112, 160))
    assert np.array_equal(crops, g["crops_112x160"])


def test_align_degenerate_and_empty():
    tgt = landmarks_target((256, 256), 0.65)
    same = np.full((1, 5, 2), 7.0, np.float32)
    crops, mats, valid = align.crop_align(np.zeros((1, 32, 32, 3), np.uint8), None, [0], same, tgt)
    assert crops.size == 0 and not valid[0]                       # cropper.py:529-531: face skipped
    crops, _, _ = align.crop_align(np.zeros((1, 32, 32, 3), np.uint8), None, [], np.zeros((0, 5, 2), np.float32), tgt)
    assert crops.shape == (0,)                                    # np.array([]) like cropper.py:550


def test_align_against_cv2_directly():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    tgt = landmarks_target((256, 256), 0.65)
    for i in range(6):
        H, W = int(rng.integers(8, 400)), int(rng.integers(8, 400))
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        src = synth.make_landmarks(1, max(min(H, W), 16), seed=50 + i)[0]
        M = cv2.estimateAffinePartial2D(src, tgt, ransacReprojThreshold=np.inf)[0]
        np.testing.assert_allclose(align.solve_partial(src, tgt), M, atol=1e-9)
        for name in align.BORDER_MODES:
            ref = cv2.warpAffine(img, M, (256, 256), borderMode=getattr(cv2, "BORDER_" + name.upper()))
            assert np.array_equal(align.warp_affine(img, M, 256, 256, name), ref), (H, W, name)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_parse_matches_reference(golden, tag):
    from oracle.make_golden import ATTR_GROUPS, MASK_GROUPS
    g = golden["parse"]
    n, h, w, _ = (int(v) for v in g[f"{tag}_shape"])
    crops = synth.make_images(n, h, w, seed=int(g[f"{tag}_seed"]))
    sd = synth.make_state_dict("bisenet", 0)
    lg = nets.bisenet_logits64(parse.preprocess(crops), sd)
    np.testing.assert_allclose(lg.numpy(), g[f"{tag}_logits64"], atol=2e-4, rtol=0)
    # tail restatements on the reference's own logits: exact
    ref_lg = torch.from_numpy(g[f"{tag}_logits64"])
    assert np.array_equal(parse.labels_from_logits64(ref_lg, (512, 512), (h, w)), g[f"{tag}_labels"])
    assert np.array_equal(parse.labels_from_logits64_sampled(g[f"{tag}_logits64"], (512, 512), (h, w)), g[f"{tag}_labels"])
    ag, mg = parse.group(g[f"{tag}_labels"], ATTR_GROUPS, MASK_GROUPS)
    assert sorted(ag) == g[f"{tag}_attr_keys"].tolist() and sorted(mg) == g[f"{tag}_mask_keys"].tolist()
    for k, v in ag.items():
        assert v == g[f"{tag}_attr_{k}"].tolist()
    for k, (vi, vm) in mg.items():
        assert vi == g[f"{tag}_maskidx_{k}"].tolist() and np.array_equal(vm, g[f"{tag}_mask_{k}"])
    # end to end through the oracle's own logits: labels may differ only at near-ties
    labels, _, _ = parse.predict(crops, sd, None, None, 2)
    assert (labels != g[f"{tag}_labels"]).mean() < 1e-4


def test_enhance_matches_reference(golden):
    g = golden["enhance"]
    sd = synth.make_state_dict("rrdbnet", 0)
    x = torch.from_numpy(synth.make_images(2, 24, 32, seed=int(g["images_seed"]))).permute(0, 3, 1, 2).float()
    y = nets.rrdbnet_forward(x[:1] / 255, sd)
    np.testing.assert_allclose(y.numpy(), g["forward0"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(enhance.bicubic_quarter_stencil(y).numpy(),
                               torch.nn.functional.interpolate(y, None, 0.25, "bicubic").numpy(), atol=1e-6)
    out = enhance.predict(x.clone(), sd, g["landmarks"], g["indices"].tolist(), 0.02).numpy()
    assert np.array_equal(out[1], g["predict"][1]) and np.array_equal(out[1], x[1].numpy())   # gated off: untouched
    assert (np.abs(out[0] - g["predict"][0]) > 0).mean() < 0.01 and np.abs(out[0] - g["predict"][0]).max() <= 1
    out = enhance.predict(x.clone(), sd, None, None, 0.02).numpy()
    assert np.abs(out - g["predict_all"]).max() <= 1
