"""fcp_group (device-side group_by_attributes / group_by_masks, bise.py:214-325) vs the oracle's restatement: integer work,
bit-exact membership and mask images."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ATTR = {"glasses": [6], "no_accessories": [-6, -9, -15, -18], "skin_hair": [1, 17], "never": [8, 7], "bg_only": [-1, -17, 0]}
MASK = {"eyes_and_eyebrows": [2, 3, 4, 5], "skin": [1], "lips": [11, 12, 13], "dup": [1, 1, 17], "none": [8]}


@pytest.fixture(scope="module")
def ctx():
    from face_crop_plus_b200 import _abi
    c = _abi.Context(0)
    yield c
    c.close()


def _labels(f, h, w, seed):
    """Blocky label maps with a few rare classes so that the > 5 / <= 5 / > 10 thresholds are exercised on both sides."""
    rng = np.random.default_rng(seed)
    lab = rng.choice([0, 1, 17], size=(f, h // 8, w // 8), p=[0.4, 0.4, 0.2]).repeat(8, 1).repeat(8, 2).astype(np.uint8)
    for i in range(f):
        for cls in (2, 3, 6, 9, 11, 15, 18):
            n = int(rng.choice([0, 3, 5, 6, 10, 11, 40]))
            ys, xs = rng.integers(0, h, n), rng.integers(0, w, n)
            lab[i, ys, xs] = cls
    return lab


@pytest.mark.parametrize("f,h,w", [(7, 64, 64), (33, 256, 256), (1, 40, 24)])
def test_group_matches_oracle(ctx, f, h, w):
    from oracle import parse
    lab = _labels(f, h, w, f)
    hist = parse.histogram(lab)
    ref_attr, ref_mask = parse.group(lab, ATTR, MASK)
    attr_m, mask_m, masks = ctx.group(lab, hist, ATTR, MASK)
    for g, k in enumerate(ATTR):
        assert np.nonzero(attr_m[g])[0].tolist() == ref_attr.get(k, []), k
    for m, k in enumerate(MASK):
        idx = np.nonzero(mask_m[m])[0].tolist()
        assert idx == (ref_mask[k][0] if k in ref_mask else []), k
        if idx:
            assert np.array_equal(masks[m][idx], ref_mask[k][1]), k
        assert set(np.unique(masks[m])) <= {0, 255}


def test_group_or_join_and_device_inputs(ctx):
    import torch
    from oracle import parse
    lab = _labels(9, 64, 64, 3)
    hist = parse.histogram(lab)
    attr_m, mask_m, masks = ctx.group(torch.from_numpy(lab).cuda(), torch.from_numpy(hist).cuda(), ATTR, MASK, join_and=False)
    for g, v in enumerate(ATTR.values()):
        tests = np.stack([(hist[:, abs(a)] > 5) if a > 0 else (hist[:, abs(a)] <= 5) for a in v], 1)
        assert np.array_equal(attr_m[g], tests.any(1))
    assert np.array_equal(mask_m[1], hist[:, 1] > 10)


def test_bisenet_shim_groups_like_oracle(ctx):
    """The host mirror's BiSeNet.groups_from (what Cropper.process_batch uses) returns the reference's dictionaries."""
    from face_crop_plus_b200.models import BiSeNet
    from oracle import parse
    lab = _labels(12, 64, 64, 5)
    m = BiSeNet(ATTR, MASK)
    m.ctx = ctx
    attr, mask = m.groups_from(lab, parse.histogram(lab))
    ref_attr, ref_mask = parse.group(lab, ATTR, MASK)
    assert attr == ref_attr and list(mask) == list(ref_mask)
    for k in ref_mask:
        assert mask[k][0] == ref_mask[k][0] and np.array_equal(mask[k][1], ref_mask[k][1])
