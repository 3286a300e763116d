"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): a few convolutions through both
tensor-core modes and the CUDA-core kernel (residual + TMA epilogue, stride 2, BN = 32/64/128) and one tiny whole-path call."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from face_crop_plus_b200 import _abi, synth
from face_crop_plus_b200.landmarks import landmarks_target

ctx = _abi.Context(0)
rng = np.random.default_rng(0)
cases = [(1, 16, 16, 64, 64, 3, 1, 1, True), (1, 16, 24, 128, 128, 1, 1, 0, True), (1, 17, 15, 64, 32, 3, 2, 1, False), (1, 12, 12, 96, 32, 3, 1, 1, False)]
for (n, h, w, cin, cout, k, s, p, res) in cases:
    x = rng.standard_normal((n, h, w, cin)).astype(np.float32)
    wt = (rng.standard_normal((cout, cin, k, k)) * 0.05).astype(np.float32)
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    r = rng.standard_normal((n, ho, wo, cout)).astype(np.float32) if res else None
    outs = [ctx.conv2d(x, wt, s, p, None, None, r, "relu", 0.0, impl) for impl in (0, 1, 2)]
    print("conv", (n, h, w, cin, cout, k, s), "max |tc - ffma|", float(np.abs(outs[1] - outs[0]).max()), float(np.abs(outs[2] - outs[0]).max()), flush=True)
if "--pipeline" in sys.argv:
    ctx.load_state_dict(_abi.MODEL_RETINAFACE, synth.make_state_dict("retinaface", 0, class_bias=4.0))
    ctx.load_state_dict(_abi.MODEL_BISENET, synth.make_state_dict("bisenet", 0))
    imgs = synth.make_images(1, 128, 128, seed=2000)
    out = ctx.pipeline(imgs, None, landmarks_target((256, 256), 0.65), (256, 256), 0.6, 0.4, "largest")
    print("pipeline faces", out["count"], flush=True)
ctx.close()
