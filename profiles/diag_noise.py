"""Diagnostic: landmark error of the three convolution paths against the oracle on the full-size detection test, twice
per path (determinism), with and without the early residual prefetch (FCP_TC_ABLATE=8 in a second process)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from face_crop_plus_b200 import _abi, synth
from oracle import pipeline

torch.set_grad_enabled(False)
sd = synth.make_state_dict("retinaface", 0, class_bias=4.8)
imgs = synth.make_images(2, 1024, 1024, seed=1234)
lms, idx, anchors, boxes = pipeline.detect(imgs, sd, 0.6, 0.4, "all")
ref = {(i, a): l for i, a, l in zip(idx, anchors, lms)}
ctx = _abi.Context(0)
ctx.load_state_dict(_abi.MODEL_RETINAFACE, sd)
prev = {}
for impl in (0, 1, 2, 2, 1):
    ctx.set_conv_impl(impl)
    out = ctx.detect(imgs, 0.6, 0.4, "all")
    got = {(i, a): l for i, a, l in zip(out["indices"].tolist(), out["anchors"].tolist(), out["landmarks"])}
    common = set(got) & set(ref)
    err = max(np.abs(got[k] - ref[k]).max() for k in common)
    heads = ctx.detect_heads(imgs)
    same = "" if impl not in prev else f" identical to previous run of impl {impl}: {np.array_equal(prev[impl], heads)}"
    prev[impl] = heads
    print(f"impl {impl}: faces {len(got)} (ref {len(ref)}, common {len(common)}), max landmark err {err:.3e} px, heads |max| {np.abs(heads).max():.3f}{same}", flush=True)
h0 = None
for impl in (0, 1, 2):
    ctx.set_conv_impl(impl)
    h = ctx.detect_heads(imgs)
    if h0 is None:
        h0 = h
    else:
        d = np.abs(h - h0)
        print(f"heads impl {impl} vs impl 0: max abs {d.max():.3e}, mean abs {d.mean():.3e}", flush=True)

# ---- does what ran before on the device change the result?  (a second context runs RRDBNet on a 1024x1024 image)
if len(sys.argv) > 1:
    ctx2 = _abi.Context(0)
    ctx2.load_state_dict(_abi.MODEL_RRDBNET, synth.make_state_dict("rrdbnet", 0))
    big = synth.make_images(1, int(sys.argv[1]), int(sys.argv[1]), seed=5)
    ctx2.enhance_u8(big, None)
    ctx2.close()
    order = [int(c) for c in (sys.argv[2] if len(sys.argv) > 2 else "210")]
    for impl in order:
        ctx.set_conv_impl(impl)
        out = ctx.detect(imgs, 0.6, 0.4, "all")
        got = {(i, a): l for i, a, l in zip(out["indices"].tolist(), out["anchors"].tolist(), out["landmarks"])}
        common = set(got) & set(ref)
        err = max(np.abs(got[k] - ref[k]).max() for k in common)
        heads = ctx.detect_heads(imgs)
        print(f"after RRDBNet {sys.argv[1]}^2 on another context: impl {impl}: max landmark err {err:.3e} px, heads identical to before: "
              f"{np.array_equal(prev[impl], heads)}, max |diff| {np.abs(prev[impl] - heads).max():.3e}", flush=True)
    ctx3 = _abi.Context(0)
    ctx3.load_state_dict(_abi.MODEL_RETINAFACE, sd)
    h3 = ctx3.detect_heads(imgs)
    print(f"fresh context, impl 2: heads identical to the first run: {np.array_equal(prev[2], h3)}, max |diff| {np.abs(prev[2] - h3).max():.3e}")
