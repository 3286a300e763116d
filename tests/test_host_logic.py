"""CPU: host-side logic of the drop-in mirror that needs no GPU."""
import numpy as np
import pytest

from face_crop_plus_b200 import synth, utils
from face_crop_plus_b200.landmarks import landmarks_target


def test_landmark_slices():
    assert [(s.start, s.stop) for s in utils.get_ldm_slices(5, 68)] == [(36, 42), (42, 48), (30, 31), (48, 49), (54, 55)]
    with pytest.raises(ValueError):
        utils.get_ldm_slices(5, 7)
    with pytest.raises(ValueError):
        utils.get_ldm_slices(4, 68)
    with pytest.raises(ValueError):
        landmarks_target((256, 256), 0.65, 4)


def test_batch_plan_geometry():
    """utils.py:317-331; the resize + padding itself runs on the GPU (tests/test_gpu_ingest.py)."""
    plans = [utils.batch_plan(h, w, (128, 96)) for h, w in [(100, 200), (300, 150), (64, 64)]]
    assert [p[3] for p in plans] == [[16, 16, 0, 0], [0, 0, 40, 40], [0, 0, 16, 16]]
    assert [p[:2] for p in plans] == [(128, 64), (48, 96), (96, 96)]
    np.testing.assert_allclose([p[2] for p in plans], [0.64, 0.32, 1.5])
    assert utils.batch_plan(64, 64, 64) == (64, 64, 1.0, [0, 0, 0, 0])          # identity at the native size


def test_enhance_gate_matches_oracle():
    from face_crop_plus_b200.models import RRDBNet
    from oracle import enhance
    rng = np.random.default_rng(0)
    lms = rng.uniform(0, 64, (7, 5, 2)).astype(np.float32)
    idx = [0, 0, 1, 3, 3, 3, 4]
    m = RRDBNet(0.05)
    g = m.gate(6, 64, 80, lms, idx)
    assert g.tolist() == [int(enhance.should_enhance(lms, idx, i, 64, 80, 0.05)) for i in range(6)]
    assert m.gate(3, 64, 80, None, None).tolist() == [1, 1, 1]


def test_parse_landmarks_file(tmp_path):
    p = tmp_path / "l.txt"
    p.write_text("a.jpg 1 2 3 4 5 6 7 8 9 10\nb.jpg 10 9 8 7 6 5 4 3 2 1\n")
    lms, names = utils.parse_landmarks_file(str(p))
    assert lms.shape == (2, 5, 2) and names.tolist() == ["a.jpg", "b.jpg"] and lms[1, 0].tolist() == [10, 9]


def test_rrdb_predict_list_without_landmarks_enhances_everything(monkeypatch):
    """rrdb.py:125-127: landmarks None (enhance-only mode of Cropper) -> every image of a list is enhanced."""
    import torch
    from face_crop_plus_b200 import models

    class FakeCtx:
        device = 0
        calls = 0

        def enhance(self, batch, gate):
            self.calls += 1
            batch += 1
            return batch

    monkeypatch.setattr(models, "bind_stream", lambda ctx: None)
    m = models.RRDBNet(0.001)
    m.ctx = FakeCtx()
    imgs = [torch.zeros((3, 4, 5)), torch.zeros((3, 6, 2))]
    out = m.predict(imgs, None, [0, 1])
    assert m.ctx.calls == 2 and all(float(o.min()) == 1.0 for o in out)
    out = m.predict([torch.zeros((3, 4, 5))], None, None)
    assert m.ctx.calls == 3
    # with landmarks: image 1 has no faces -> skipped; image 0 gated by its face factor (normalised by images[0])
    lms = np.zeros((1, 5, 2), np.float32)
    lms[0, 4] = [1.0, 1.0]
    m.ctx.calls = 0
    m.predict([torch.zeros((3, 100, 100)), torch.zeros((3, 50, 50))], lms, [0])
    assert m.ctx.calls == 1


def test_parse_landmarks_csv_skips_header(tmp_path):
    """utils.py:70-71: a .csv is read with delimiter ',' and its header row skipped."""
    p = tmp_path / "l.csv"
    p.write_text("name,x1,y1,x2,y2,x3,y3,x4,y4,x5,y5\na.jpg,1,2,3,4,5,6,7,8,9,10\nb.jpg,10,9,8,7,6,5,4,3,2,1\n")
    lms, names = utils.parse_landmarks_file(str(p))
    assert lms.shape == (2, 5, 2) and names.tolist() == ["a.jpg", "b.jpg"] and lms[0, 4].tolist() == [9, 10]


def test_bench_roofline_traffic_comes_from_a_committed_ncu_summary():
    """bench.py's roofline.traffic is read from the committed ncu launch-list summary of the newest build that has one."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    import bench
    t = bench.committed_traffic()
    assert t and t["kernel"] == "conv_tc_kernel" and t["launches"] > 100
    assert 0.5 < t["dram_bytes_per_launch"] / t["algorithmic_bytes_per_launch"] < 1.5
    assert (Path(bench.REPO) / "profiles" / "r2b_launches_dram.txt").exists()
