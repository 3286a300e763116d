"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference/src).

Run in the build container only (``python -m oracle.make_golden``); the GPU box has no /root/reference and
consumes the committed fixtures.  The reference finds our seeded synthetic state_dicts because we save them
under ``$TORCH_HOME/hub/checkpoints/<WEIGHTS_FILENAME>`` (``torch.hub.load_state_dict_from_url`` cache,
_layers.py:27-35).  ``unidecode`` (utils.py:9) is not installed, so a stub module is registered first — only
``clean_names`` uses it.
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
import zlib
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
GOLD = REPO / "tests" / "golden"
sys.path.insert(0, str(REPO))

from face_crop_plus_b200 import synth  # noqa: E402

DET_CLASS_BIAS = 4.0   # denser candidates than the bench default: the golden images are small
ATTR_GROUPS = {"glasses": [6], "no_accessories": [-6, -9, -15, -18], "skin_hair": [1, 17], "never": [8, 7]}
MASK_GROUPS = {"eyes_and_eyebrows": [2, 3, 4, 5], "skin": [1], "lips": [11, 12, 13]}


def import_reference(torch_home: str):
    os.environ["TORCH_HOME"] = torch_home
    sys.modules.setdefault("unidecode", types.ModuleType("unidecode"))
    sys.path.insert(0, "/root/reference/src")
    ckpt = Path(torch_home) / "hub" / "checkpoints"
    ckpt.mkdir(parents=True, exist_ok=True)
    for model, fname in synth.REFERENCE_FILENAMES.items():
        kw = {"class_bias": DET_CLASS_BIAS} if model == "retinaface" else {}
        torch.save(synth.make_state_dict(model, 0, **kw), ckpt / fname)
    import face_crop_plus  # noqa: F401
    from face_crop_plus.cropper import Cropper
    from face_crop_plus.models import BiSeNet, RetinaFace, RRDBNet
    return Cropper, RetinaFace, BiSeNet, RRDBNet


def main():
    import cv2
    torch.set_grad_enabled(False)
    torch.manual_seed(0)
    GOLD.mkdir(parents=True, exist_ok=True)
    with tempfile.TemporaryDirectory() as th:
        Cropper, RetinaFace, BiSeNet, RRDBNet = import_reference(th)

        # ---- spec check: our key/shape lists == the reference modules' state_dicts
        for model, cls in (("retinaface", RetinaFace), ("bisenet", BiSeNet), ("rrdbnet", RRDBNet)):
            ref = {k: tuple(v.shape) for k, v in cls().state_dict().items()}
            ours = {k: tuple(v.shape) for k, v in synth.make_state_dict(model, 0).items()}
            assert ref == ours, (model, set(ref) ^ set(ours))
        digests = {m: synth.state_dict_digest(synth.make_state_dict(m, 0, **({"class_bias": DET_CLASS_BIAS}
                                                                            if m == "retinaface" else {})))
                   for m in synth.SPECS}

        # ---- detect: RetinaFace.forward + predict for every strategy, non-square 320x384 batch
        imgs = synth.make_images(3, 320, 384, seed=1234)
        x = torch.from_numpy(imgs).permute(0, 3, 1, 2).float()
        det = {"images_seed": 1234, "shape": np.array(imgs.shape)}
        model = RetinaFace(strategy="all", vis=0.6).load("cpu")
        xin = x[:, [2, 1, 0]] - torch.tensor([104, 117, 123]).view(3, 1, 1)
        scores, boxes, ldms = model(xin)
        det.update(scores=scores[..., 1].numpy(), boxes_raw=boxes.numpy(), ldms_raw=ldms.numpy())
        # raw class logits are not exposed by forward(); recompute them through the reference modules
        fpn = model.fpn(model.body(xin))
        fts = [model.ssh1(fpn[0]), model.ssh2(fpn[1]), model.ssh3(fpn[2])]
        det["cls_raw"] = torch.cat([model.ClassHead[i](f) for i, f in enumerate(fts)], 1).numpy()
        for strat in ("all", "best", "largest"):
            model.strategy = strat
            l, i = model.predict(x.clone())
            det[f"landmarks_{strat}"], det[f"indices_{strat}"] = l, np.array(i, dtype=np.int64)
        model.vis_threshold = 0.999999
        l, i = RetinaFace(strategy="all", vis=1.0).load("cpu").predict(x.clone())
        assert len(i) == 0 and l.shape == (0, 5, 2), (l.shape, i)
        np.savez_compressed(GOLD / "detect.npz", **det)

        # ---- align: Cropper.crop_align, all border modes, with and without paddings, skew on/off
        ali = {}
        big = synth.make_images(2, 300, 260, seed=4321)
        lms = synth.make_landmarks(4, 200, seed=3)
        lms[3] = lms[3] * 0.25 + 20                        # a small face near the corner -> borders get sampled
        idx = [0, 0, 1, 1]
        pads = np.array([[10, 6, 4, 8], [0, 0, 0, 0]])
        ali.update(images_seed=4321, landmarks=lms, indices=np.array(idx), paddings=pads)
        for mode in ("constant", "replicate", "reflect", "wrap", "reflect_101"):
            for skew in (False, True):
                c = Cropper(det_threshold=None, enh_threshold=None, padding=mode, allow_skew=skew,
                            landmarks=(lms, np.array(["a"] * 4)), output_size=(256, 256), face_factor=0.65)
                crops = c.crop_align(big, pads, idx, lms)
                ali[f"crc_{mode}_{int(skew)}"] = np.array([zlib.crc32(a.tobytes()) for a in crops], dtype=np.int64)
                if mode == "constant" and not skew:
                    ali["crops_constant_0"] = crops
        c = Cropper(det_threshold=None, enh_threshold=None, landmarks=(lms, np.array(["a"] * 4)),
                    output_size=(112, 160), face_factor=0.8)
        ali["target_112x160_0.8"] = c.landmarks_target
        ali["crops_112x160"] = c.crop_align(big, None, idx, lms)
        ali["target_256_0.65"] = Cropper(det_threshold=None, enh_threshold=None,
                                         landmarks=(lms, np.array(["a"] * 4))).landmarks_target
        ali["matrices_partial"] = np.stack([cv2.estimateAffinePartial2D(l, ali["target_256_0.65"],
                                                                        ransacReprojThreshold=np.inf)[0] for l in lms])
        ali["matrices_affine"] = np.stack([cv2.estimateAffine2D(l, ali["target_256_0.65"],
                                                                ransacReprojThreshold=np.inf)[0] for l in lms])
        np.savez_compressed(GOLD / "align.npz", **ali)

        # ---- parse: BiSeNet.forward / predict on 3 crops of 256x256 and 2 of 200x240
        par = {}
        bis = BiSeNet(ATTR_GROUPS, MASK_GROUPS, max_batch_size=2).load("cpu")
        for tag, (n, h, w, seed) in {"a": (3, 256, 256, 77), "b": (2, 200, 240, 99)}.items():
            crops = synth.make_images(n, h, w, seed=seed)
            xc = torch.from_numpy(crops).permute(0, 3, 1, 2).float()
            mean = torch.tensor(bis.mean).view(1, 3, 1, 1)
            std = torch.tensor(bis.std).view(1, 3, 1, 1)
            xi = (torch.nn.functional.interpolate(xc.div(255), (512, 512), mode="bilinear") - mean) / std
            feat = bis.conv_out(bis.ffm(*bis.cp(xi)))
            o = bis(xi)
            labels = torch.nn.functional.interpolate(o, (h, w), mode="nearest").argmax(1)
            ag, mg = bis.predict(xc)
            par[f"{tag}_seed"] = seed
            par[f"{tag}_shape"] = np.array(crops.shape)
            par[f"{tag}_logits64"] = feat.numpy()
            par[f"{tag}_labels"] = labels.numpy().astype(np.uint8)
            for k, v in ag.items():
                par[f"{tag}_attr_{k}"] = np.array(v, dtype=np.int64)
            for k, (vi, vm) in mg.items():
                par[f"{tag}_maskidx_{k}"] = np.array(vi, dtype=np.int64)
                par[f"{tag}_mask_{k}"] = vm
            par[f"{tag}_attr_keys"] = np.array(sorted(ag.keys()))
            par[f"{tag}_mask_keys"] = np.array(sorted(mg.keys()))
        np.savez_compressed(GOLD / "parse.npz", **par)

        # ---- enhance: RRDBNet.forward on 1x3x24x32 and predict on a 2-image batch (one gated off)
        enh = {}
        rr = RRDBNet(min_face_factor=0.02).load("cpu")
        small = synth.make_images(2, 24, 32, seed=5)
        xs = torch.from_numpy(small).permute(0, 3, 1, 2).float()
        enh["images_seed"] = 5
        enh["forward0"] = rr(xs[:1] / 255).numpy()
        l2 = np.zeros((2, 5, 2), np.float32)
        l2[0, 4] = (3, 3)        # tiny face in image 0 -> factor 9/768=0.0117 <= 0.02 -> enhanced
        l2[1, 4] = (20, 20)      # big face in image 1 -> 400/768 -> untouched
        enh["landmarks"], enh["indices"] = l2, np.array([0, 1])
        enh["predict"] = rr.predict(xs.clone(), l2, [0, 1]).numpy()
        enh["predict_all"] = rr.predict(xs.clone(), None, None).numpy()
        np.savez_compressed(GOLD / "enhance.npz", **enh)

        # ---- whole path: Cropper.process_dir (detect + align + parse) on PNG files, outputs read back
        with tempfile.TemporaryDirectory() as d:
            src, dst = Path(d) / "in", Path(d) / "out"
            src.mkdir()
            pimgs = synth.make_images(4, 256, 320, seed=2000)
            for k, im in enumerate(pimgs):
                cv2.imwrite(str(src / f"img{k}.png"), cv2.cvtColor(im, cv2.COLOR_RGB2BGR))
            pipe = {"images_seed": 2000}
            for strat, thr, full in (("largest", 0.6, True), ("all", 0.95, False)):
                cr = Cropper(output_size=128, output_format="png", resize_size=(320, 256), face_factor=0.65,
                             strategy=strat, det_threshold=thr, enh_threshold=None, attr_groups=ATTR_GROUPS,
                             mask_groups=MASK_GROUPS, batch_size=4, num_processes=1, device="cpu")
                out = dst / strat
                cr.process_dir(str(src), str(out))
                files = sorted(str(p.relative_to(out)) for p in out.rglob("*.png"))
                data = [cv2.imread(str(out / f), cv2.IMREAD_UNCHANGED) for f in files]
                pipe[f"{strat}_files"] = np.array(files)
                pipe[f"{strat}_crc"] = np.array([zlib.crc32(np.ascontiguousarray(a).tobytes()) for a in data],
                                                dtype=np.int64)
                pipe[f"{strat}_ndim"] = np.array([a.ndim for a in data])
                if full:   # BGR as stored on disk; masks are single-channel
                    for f, a in zip(files, data):
                        pipe[f"{strat}_data/{f}"] = a
            np.savez_compressed(GOLD / "pipeline.npz", **pipe)

        np.savez(GOLD / "meta.npz", class_bias=DET_CLASS_BIAS, **{f"digest_{k}": v for k, v in digests.items()},
                 torch=torch.__version__, cv2=cv2.__version__)
    for f in sorted(GOLD.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
