"""Timing of BASELINE.json configs[3] (RRDBNet x4 enhancement, bs=32, 256x256 -> 1024x1024, 1 GPU).  Not the bench.py
contract (that is the detect+align+parse metric); used to fill the RRDB row of profiles/README.md.

    python profiles/enhance_bench.py [--batch 32] [--size 256] [--iters 2]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from face_crop_plus_b200 import _abi, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
ctx = _abi.Context(0)
ctx.load_state_dict(_abi.MODEL_RRDBNET, synth.make_state_dict("rrdbnet", 0))
x = torch.from_numpy(synth.make_images(min(a.batch, 8), a.size, a.size, seed=5)).permute(0, 3, 1, 2).float()
x = x.repeat(-(-a.batch // len(x)), 1, 1, 1)[:a.batch].contiguous().cuda()
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
ctx.enhance(x.clone())
ctx.profile(True); ctx.profile_read()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(a.iters):
    ctx.enhance(x.clone())
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / a.iters
prof = ctx.profile_read()
gflop = prof["conv_flops"] / a.iters / a.batch / 1e9
print(f"RRDBNet enhance: {a.batch} images {a.size}x{a.size}: {dt * 1e3:.1f} ms/batch = {a.batch / dt:.2f} img/s; "
      f"{gflop:.1f} GFLOP/img in convs, conv kernels {prof['conv_flops'] / (prof['conv_ms'] / 1e3) / 1e12:.1f} TFLOP/s "
      f"(share {prof['conv_ms'] / a.iters / (dt * 1e3):.2f})")
