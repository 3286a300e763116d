"""Diagnostic: run-to-run determinism of fcp_pipeline on one device-resident batch (same process, same inputs)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from face_crop_plus_b200 import _abi, synth
from face_crop_plus_b200.landmarks import landmarks_target

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
par_mb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = _abi.Context(0)
ctx.load_state_dict(_abi.MODEL_RETINAFACE, synth.make_state_dict("retinaface", 0, class_bias=4.8))
ctx.load_state_dict(_abi.MODEL_BISENET, synth.make_state_dict("bisenet", 0))
ctx.set_micro_batch(16, par_mb)
base = synth.make_images(16, 1024, 1024, seed=1234)
imgs = torch.from_numpy(np.concatenate([base] * (n // 16))).cuda()
tgt = landmarks_target((256, 256), 0.65)
runs = []
for r in range(6):
    out = ctx.pipeline(imgs, None, tgt, (256, 256), 0.6, 0.4, "largest")
    heads = ctx.detect_heads(imgs[:16])
    runs.append((out["landmarks"].copy(), out["crops"].copy(), out["labels"].copy(), heads))
for r in range(1, 6):
    same = [np.array_equal(runs[0][k], runs[r][k]) for k in range(4)]
    d = [float(np.abs(runs[0][k].astype(np.float64) - runs[r][k].astype(np.float64)).max()) for k in range(4)]
    print(f"run {r} vs 0: landmarks/crops/labels/heads identical {same}, max |diff| {d}", flush=True)
# duplicates inside the batch (image i and i+16 are the same picture)
lm = runs[0][0]
if n >= 32 and len(lm) == n:
    print("duplicates within the batch agree:", bool(np.array_equal(lm[:16], lm[16:32])), float(np.abs(lm[:16] - lm[16:32]).max()))
