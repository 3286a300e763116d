"""Diagnostic: is Cropper.process_dir deterministic across runs and across num_processes?"""
import hashlib, os, shutil, sys, tempfile
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import cv2
from face_crop_plus_b200 import Cropper, synth

base = synth.make_images(16, 1024, 1024, seed=1234)
root = Path(tempfile.mkdtemp(prefix="fcp_det_"))
src = root / "in"; src.mkdir()
for i in range(192):
    cv2.imwrite(str(src / f"img{i:05d}.jpg"), cv2.cvtColor(base[i % 16], cv2.COLOR_RGB2BGR), [cv2.IMWRITE_JPEG_QUALITY, 90])
sds = {"det": synth.make_state_dict("retinaface", 0, class_bias=4.8), "par": synth.make_state_dict("bisenet", 0)}
def run(procs, tag):
    cr = Cropper(output_size=256, output_format="png", resize_size=1024, strategy="largest", det_threshold=0.6,
                 mask_groups={"skin": [1]}, batch_size=64, num_processes=procs, device="cuda:0", state_dicts=sds)
    out = root / tag
    cr.process_dir(str(src), str(out), desc=None)
    res = {}
    for p in sorted(out.rglob("*.png")):
        res[str(p.relative_to(out))] = hashlib.md5(p.read_bytes()).hexdigest()
    return res
a = run(1, "a"); b = run(1, "b"); c = run(4, "c"); d = run(16, "d")
for name, r in (("P=1 again", b), ("P=4", c), ("P=16", d)):
    missing = sorted(set(a) ^ set(r))
    diff = [k for k in set(a) & set(r) if a[k] != r[k]]
    print(f"{name}: files {len(r)} (first run {len(a)}), only-in-one {missing[:6]}, content differs in {len(diff)} files {sorted(diff)[:4]}")
shutil.rmtree(root)
