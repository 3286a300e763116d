"""Times one fused conv launch through the C-ABI test hook on device-resident tensors (kernel time from fcp_profile).

    python profiles/conv_probe.py N H W CIN COUT K STRIDE [res] [impl]
"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from face_crop_plus_b200 import _abi  # noqa: E402

n, h, w, cin, cout, k, stride = (int(v) for v in sys.argv[1:8])
use_res = len(sys.argv) > 8 and sys.argv[8] == "res"
impl = int(sys.argv[9]) if len(sys.argv) > 9 else 1
pad = k // 2
ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
ctx = _abi.Context(0)
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((n, h, w, cin), device="cuda", generator=g)
wt = (torch.randn((cout, cin, k, k)) * (2.0 / (cin * k * k)) ** 0.5).contiguous()
res = torch.randn((n, ho, wo, cout), device="cuda", generator=g) if use_res else None
out = torch.empty((n, ho, wo, cout), device="cuda")
ctx.profile(True)
for it in range(4):
    ctx.check(ctx.lib.fcp_conv2d(ctx.h, x.data_ptr(), n, h, w, cin, wt.data_ptr(), cout, k, stride, pad, None, None,
                                 res.data_ptr() if use_res else None, 1, 0.0, impl, out.data_ptr()))
    p = ctx.profile_read()
flops = 2.0 * n * ho * wo * cout * cin * k * k
byt = 4.0 * (x.numel() + out.numel() * (2 if use_res else 1))
print(f"conv n{n} {h}x{w} cin{cin} cout{cout} k{k} s{stride} res={use_res} impl={impl}: {p['conv_ms']:.3f} ms  "
      f"{flops / p['conv_ms'] / 1e9:.1f} TFLOP/s  {byt / p['conv_ms'] / 1e6:.0f} GB/s (algorithmic bytes)")
