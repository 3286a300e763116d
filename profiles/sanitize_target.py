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
# round 2 additions: K-blocks pairing 32-channel units across taps + the padded 96-wide tile, the cta_group::2 variant of the
# wide tiles (an odd number of pixel tiles: the pair's phantom tile), and RRDBNet's source-major dense blocks (partial
# activation, partial sums accumulated in place through the residual path)
import os
x = rng.standard_normal((1, 12, 20, 32)).astype(np.float32)
wt = (rng.standard_normal((96, 32, 3, 3)) * 0.05).astype(np.float32)
r = rng.standard_normal((1, 12, 20, 96)).astype(np.float32)
outs = [ctx.conv2d(x, wt, 1, 1, None, None, r, "lrelu", 0.2, impl) for impl in (0, 2)]
print("conv 32->96 (unit pairs, BN=128 padded)", float(np.abs(outs[1] - outs[0]).max()), flush=True)
os.environ["FCP_TC_PAIR"] = "2"
for shape in ((1, 24, 40), (3, 17, 15)):
    x = rng.standard_normal(shape + (128,)).astype(np.float32)
    wt = (rng.standard_normal((256, 128, 3, 3)) * 0.03).astype(np.float32)
    r = rng.standard_normal(shape + (256,)).astype(np.float32)
    outs = [ctx.conv2d(x, wt, 1, 1, None, None, r, "relu", 0.0, impl) for impl in (0, 2)]
    print("conv 128->256 cta_group::2", shape, float(np.abs(outs[1] - outs[0]).max()), flush=True)
del os.environ["FCP_TC_PAIR"]
ctx.load_state_dict(_abi.MODEL_RRDBNET, synth.make_state_dict("rrdbnet", 0), rrdb_blocks=1)     # one RRDB = 3 dense blocks
xe = rng.random((1, 3, 24, 20)).astype(np.float32)
ye = ctx.enhance_forward(xe)
print("rrdbnet source-major", ye.shape, float(np.abs(ye).max()), flush=True)
if "--pipeline" in sys.argv:
    ctx.load_state_dict(_abi.MODEL_RETINAFACE, synth.make_state_dict("retinaface", 0, class_bias=4.0))
    ctx.load_state_dict(_abi.MODEL_BISENET, synth.make_state_dict("bisenet", 0))
    imgs = synth.make_images(1, 128, 128, seed=2000)
    out = ctx.pipeline(imgs, None, landmarks_target((256, 256), 0.65), (256, 256), 0.6, 0.4, "largest")
    print("pipeline faces", out["count"], flush=True)
ctx.close()
