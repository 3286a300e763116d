"""Diagnostic: run-to-run determinism of single convolutions through fcp_conv2d (impl 2 vs 1), bench-sized shapes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from face_crop_plus_b200 import _abi
ctx = _abi.Context(0)
rng = np.random.default_rng(0)
shapes = [  # n, h, w, cin, cout, k, stride, pad, res
    (16, 256, 256, 64, 64, 3, 1, 1, False), (8, 256, 256, 64, 256, 1, 1, 0, True), (16, 128, 128, 128, 128, 3, 1, 1, False),
    (16, 64, 64, 256, 256, 3, 1, 1, False), (16, 64, 64, 1024, 256, 1, 1, 0, False), (16, 64, 64, 256, 1024, 1, 1, 0, True),
    (16, 128, 128, 256, 32, 1, 1, 0, False), (16, 128, 128, 256, 64, 3, 1, 1, False), (16, 128, 128, 128, 512, 1, 1, 0, True)]
impls = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "21")]
for (n, h, w, cin, cout, k, s, p, res) in shapes:
    x = np.maximum(rng.standard_normal((n, h, w, cin)).astype(np.float32), 0)
    wt = (rng.standard_normal((cout, cin, k, k)) * (2.0 / (cin * k * k)) ** 0.5).astype(np.float32)
    ho, wo = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    r = rng.standard_normal((n, ho, wo, cout)).astype(np.float32) if res else None
    for impl in impls:
        outs = [ctx.conv2d(x, wt, s, p, None, None, r, "relu", 0.0, impl) for _ in range(4)]
        same = [bool(np.array_equal(outs[0], o)) for o in outs[1:]]
        d = max(float(np.abs(outs[0] - o).max()) for o in outs[1:])
        nbad = max(int((outs[0] != o).sum()) for o in outs[1:])
        print(f"impl {impl} k{k} {cin}->{cout} M={n*ho*wo} res={int(res)}: identical {same}, max |diff| {d:.3e}, differing values {nbad}", flush=True)
