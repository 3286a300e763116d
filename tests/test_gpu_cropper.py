"""GPU: the drop-in ``Cropper`` API (process_dir / process_batch / crop_align / model shims) against the reference's own
outputs (tests/golden/pipeline.npz, produced by the unmodified reference's Cropper.process_dir) and the oracle."""
import numpy as np
import pytest
import torch

from face_crop_plus_b200 import synth

pytestmark = pytest.mark.gpu
ATTR_GROUPS = {"glasses": [6], "no_accessories": [-6, -9, -15, -18], "skin_hair": [1, 17], "never": [8, 7]}
MASK_GROUPS = {"eyes_and_eyebrows": [2, 3, 4, 5], "skin": [1], "lips": [11, 12, 13]}


@pytest.fixture(scope="module")
def state_dicts():
    return {"det": synth.make_state_dict("retinaface", 0, class_bias=4.0), "par": synth.make_state_dict("bisenet", 0),
            "enh": synth.make_state_dict("rrdbnet", 0)}


def write_pngs(dirpath, images):
    import cv2
    dirpath.mkdir()
    for k, im in enumerate(images):
        cv2.imwrite(str(dirpath / f"img{k}.png"), cv2.cvtColor(im, cv2.COLOR_RGB2BGR))


@pytest.mark.parametrize("strategy,thr", [("largest", 0.6), ("all", 0.95)])
def test_process_dir_matches_reference_tree(tmp_path, golden, state_dicts, strategy, thr):
    import cv2
    from face_crop_plus_b200 import Cropper
    g = golden["pipeline"]
    write_pngs(tmp_path / "in", synth.make_images(4, 256, 320, seed=int(g["images_seed"])))
    cr = Cropper(output_size=128, output_format="png", resize_size=(320, 256), face_factor=0.65, strategy=strategy,
                 det_threshold=thr, enh_threshold=None, attr_groups=ATTR_GROUPS, mask_groups=MASK_GROUPS, batch_size=4,
                 num_processes=1, device="cuda:0", state_dicts=state_dicts)
    out = tmp_path / "out"
    cr.process_dir(str(tmp_path / "in"), str(out), desc=None)
    files = sorted(str(p.relative_to(out)) for p in out.rglob("*.png"))
    assert files == g[f"{strategy}_files"].tolist()                       # same directory tree and file names
    if strategy == "largest":
        for f in files:
            got, ref = cv2.imread(str(out / f), cv2.IMREAD_UNCHANGED), g[f"largest_data/{f}"]
            assert got.shape == ref.shape
            # float32 landmark noise (<=1e-3 px) may move a 1/32-px interpolation bin or a near-tie label
            assert (got != ref).mean() < 0.02, f


def test_landmarks_only_config_c1(tmp_path, state_dicts):
    """BASELINE.json configs[0]: 8 synthetic 256x256 images, landmarks supplied, strategy=largest."""
    import cv2
    from face_crop_plus_b200 import Cropper
    from face_crop_plus_b200.landmarks import landmarks_target
    from oracle import align
    imgs = synth.make_images(8, 256, 256, seed=300)
    write_pngs(tmp_path / "in", imgs)
    names = np.array([f"img{k}.png" for k in range(8)])
    lms = synth.make_landmarks(8, 256, seed=0)
    cr = Cropper(landmarks=(lms, names), det_threshold=None, output_format="png", batch_size=3, device="cuda:0")
    assert cr.det_model is None and cr.par_model is None
    cr.process_dir(str(tmp_path / "in"), str(tmp_path / "out"), desc=None)
    tgt = landmarks_target((256, 256), 0.65)
    for k in range(8):
        got = cv2.cvtColor(cv2.imread(str(tmp_path / "out" / f"img{k}.png")), cv2.COLOR_BGR2RGB)
        assert np.array_equal(got, align.warp_affine(imgs[k], align.solve_partial(lms[k], tgt), 256, 256))
    # 68-point annotations are reduced to 5 points by slice means (cropper.py:828-831)
    lms68 = np.repeat(lms[:, :1], 68, 1)
    for j, (a, b) in enumerate([(36, 42), (42, 48), (30, 31), (48, 49), (54, 55)]):
        lms68[:, a:b] = lms[:, j:j + 1]
    cr68 = Cropper(landmarks=(lms68, names), det_threshold=None, output_format="png", batch_size=8, device="cuda:0")
    cr68.process_dir(str(tmp_path / "in"), str(tmp_path / "out68"), desc=None)
    for k in range(8):   # the float32 mean of six equal values may differ from the value by an ulp -> compare with tolerance
        a, b = cv2.imread(str(tmp_path / "out68" / f"img{k}.png")), cv2.imread(str(tmp_path / "out" / f"img{k}.png"))
        assert (a != b).mean() < 0.01


def test_model_shims_follow_reference_contracts(state_dicts, golden):
    from face_crop_plus_b200.models import BiSeNet, RetinaFace, RRDBNet
    g = golden["detect"]
    n, h, w, _ = (int(v) for v in g["shape"])
    x = torch.from_numpy(synth.make_images(n, h, w, seed=int(g["images_seed"]))).permute(0, 3, 1, 2).float().cuda()
    det = RetinaFace("largest", 0.6).load("cuda:0", state_dicts["det"])
    lms, idx = det.predict(x)                                             # f32 NCHW 0..255 on the model device
    assert isinstance(lms, np.ndarray) and lms.dtype == np.float32 and idx == g["indices_largest"].tolist()
    assert np.abs(lms - g["landmarks_largest"]).max() < 1e-3
    det.strategy = "bogus"
    with pytest.raises(ValueError):
        det.predict(x)
    gp = golden["parse"]
    crops = synth.make_images(3, 256, 256, seed=int(gp["a_seed"]))
    par = BiSeNet(ATTR_GROUPS, MASK_GROUPS, 2).load("cuda:0", state_dicts["par"])
    ag, mg = par.predict(torch.from_numpy(crops).permute(0, 3, 1, 2).float())
    assert sorted(ag) == gp["a_attr_keys"].tolist() and sorted(mg) == gp["a_mask_keys"].tolist()
    for k, v in ag.items():
        assert v == gp[f"a_attr_{k}"].tolist()
    for k, (vi, vm) in mg.items():
        assert vi == gp[f"a_maskidx_{k}"].tolist() and (vm != gp[f"a_mask_{k}"]).mean() < 1e-4
    ge = golden["enhance"]
    xe = torch.from_numpy(synth.make_images(2, 24, 32, seed=int(ge["images_seed"]))).permute(0, 3, 1, 2).float().contiguous()
    enh = RRDBNet(0.02).load("cuda:0", state_dicts["enh"])
    out = enh.predict(xe, ge["landmarks"], ge["indices"].tolist())
    assert out is xe and np.abs(out.numpy() - ge["predict"]).max() <= 1


@pytest.mark.parametrize("k", [5, 12, 17, 21, 29, 49, 68, 98, 106])
def test_reduce_landmarks_bit_exact(k):
    """fcp_reduce_landmarks == the slice means of cropper.py:828-831 (float32 ``landmarks[:, s].mean(1)``), bit for bit."""
    from face_crop_plus_b200 import utils
    from face_crop_plus_b200.models import get_context
    ctx = get_context("cuda:0")
    rng = np.random.default_rng(k)
    lms = (rng.random((37, k, 2)) * 1000).astype(np.float32)
    ref = np.stack([lms[:, s].mean(1) for s in utils.get_ldm_slices(5, k)], 1)
    got = ctx.reduce_landmarks(lms)
    assert got.dtype == np.float32 and np.array_equal(got, ref)


def test_reduce_landmarks_rejects_unknown_counts():
    from face_crop_plus_b200.models import get_context
    with pytest.raises(ValueError):
        get_context("cuda:0").reduce_landmarks(np.zeros((2, 7, 2), np.float32))
    assert get_context("cuda:0").reduce_landmarks(np.zeros((0, 68, 2), np.float32)).shape == (0, 5, 2)
