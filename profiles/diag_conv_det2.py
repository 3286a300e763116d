"""Diagnostic: structure of the nondeterministic outputs of one convolution shape."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from face_crop_plus_b200 import _abi
ctx = _abi.Context(0)
rng = np.random.default_rng(0)
n, h, w, cin, cout, k, res = 16, 64, 64, 256, 1024, 1, int(sys.argv[1]) if len(sys.argv) > 1 else 1
x = np.maximum(rng.standard_normal((n, h, w, cin)).astype(np.float32), 0)
wt = (rng.standard_normal((cout, cin, k, k)) * (2.0 / cin) ** 0.5).astype(np.float32)
r = rng.standard_normal((n, h, w, cout)).astype(np.float32) if res else None
ref = ctx.conv2d(x, wt, 1, 0, None, None, r, "relu", 0.0, 1)
bad_runs = 0
for it in range(16):
    o = ctx.conv2d(x, wt, 1, 0, None, None, r, "relu", 0.0, 2)
    d = np.abs(o - ref)
    bad = d > 1e-3
    if bad.any():
        bad_runs += 1
        idx = np.argwhere(bad)
        print(f"run {it}: {bad.sum()} bad values; images {np.unique(idx[:,0])}, rows {np.unique(idx[:,1])}, cols {idx[:,2].min()}..{idx[:,2].max()} ({len(np.unique(idx[:,2]))} distinct), "
              f"channels {idx[:,3].min()}..{idx[:,3].max()} ({len(np.unique(idx[:,3]))} distinct), max err {d.max():.3f}", flush=True)
        # per (pixel) count
        pix = {}
        for a in idx: pix.setdefault((a[0], a[1], a[2]), []).append(a[3])
        for kx, v in list(pix.items())[:6]:
            print("   pixel", kx, "channels", min(v), "..", max(v), "count", len(v), "err sample", d[kx[0], kx[1], kx[2], v[0]], "out", o[kx[0], kx[1], kx[2], v[0]], "ref", ref[kx[0], kx[1], kx[2], v[0]],
                  "res", None if r is None else r[kx[0], kx[1], kx[2], v[0]])
print("bad runs", bad_runs, "of 16")
