#!/usr/bin/env python
"""The kernel bar BASELINE.md §3 names: the UNMODIFIED reference (baseline/_ref) run on the same B200 through its own code
path — torch `nn.Conv2d` -> cuDNN — with TF32 off (fp32, the arithmetic our kernel is held to) and on (cuDNN's single-pass
TF32, which the parity bar excludes: SURVEY.md §7.3).  The convolution stacks (`RetinaFace.forward`, `BiSeNet.forward`) are
timed apart from the Python/torch post-processing the reference wraps around them (PriorBox, decode, per-image NMS loop,
crop_align on the CPU via cv2, grouping), so both "beats the reference's kernels" and "beats the reference end to end on the
GPU" can be read off.

    python profiles/ref_on_b200.py [--batch 16] [--steps 3]      ->  one JSON line per (tf32 off / on)
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402


def timed(fn, steps):
    torch.cuda.synchronize()
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--class-bias", type=float, default=4.8)
    args = ap.parse_args()
    if not bench.reference_available():
        print(json.dumps({"unavailable": "baseline/_ref is not installed"}))
        return
    ref = bench.load_reference(args.class_bias)
    ru = ref["utils"]
    dev = torch.device("cuda:0")
    c = ref["Cropper"](output_size=256, resize_size=args.size, strategy="largest", det_threshold=0.6, enh_threshold=None,
                       mask_groups={"skin": [1]}, batch_size=args.batch, device=dev)
    imgs = list(bench.synthetic_batch(args.batch, args.size))
    B = args.batch
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.benchmark = True
        with torch.no_grad():
            batch, _, paddings = ru.as_batch(imgs, c.resize_size)
            x = ru.as_tensor(batch, dev)
            xin = x[:, [2, 1, 0]] - torch.tensor([104, 117, 123], device=dev).view(3, 1, 1)
            t_det_net, _ = timed(lambda: c.det_model(xin), args.steps)
            t_det_predict, (landmarks, indices) = timed(lambda: c.det_model.predict(x.clone()), args.steps)
            landmarks = landmarks - paddings[indices][:, None, [2, 0]]
            t0 = time.perf_counter()
            crops = c.crop_align(ru.as_numpy(x), paddings, indices, landmarks)
            t_align = time.perf_counter() - t0
            ct = ru.as_tensor(crops, dev)
            mean = torch.tensor(c.par_model.mean, device=dev).view(1, 3, 1, 1)
            std = torch.tensor(c.par_model.std, device=dev).view(1, 3, 1, 1)
            pin = (torch.nn.functional.interpolate(ct.div(255), (512, 512), mode="bilinear") - mean) / std
            t_par_net, _ = timed(lambda: c.par_model(pin), args.steps)
            t_par_predict, _ = timed(lambda: c.par_model.predict(ct), args.steps)

            def whole():
                b, _, p = ru.as_batch(imgs, c.resize_size)
                xx = ru.as_tensor(b, dev)
                l, i = c.det_model.predict(xx)
                l -= p[i][:, None, [2, 0]]
                cr = c.crop_align(ru.as_numpy(xx), p, i, l)
                return c.par_model.predict(ru.as_tensor(cr, dev))
            t_whole, _ = timed(whole, args.steps)
        # 221.7 + 4.9 GFLOP / image in the detector convolutions, 26.8 GFLOP / face in the parser's
        print(json.dumps({
            "what": "unmodified reference on cuda:0 (torch nn.Conv2d -> cuDNN)", "tf32": tf32, "batch": B, "size": args.size,
            "faces": len(indices),
            "detector_conv_stack_ms_per_image": 1e3 * t_det_net / B,
            "detector_conv_stack_tflops": 226.6e9 * B / t_det_net / 1e12,
            "detector_predict_ms_per_image": 1e3 * t_det_predict / B,
            "detector_post_python_ms_per_image": 1e3 * (t_det_predict - t_det_net) / B,
            "crop_align_cpu_ms_per_face": 1e3 * t_align / max(1, len(indices)),
            "parser_conv_stack_ms_per_face": 1e3 * t_par_net / max(1, len(crops)),
            "parser_conv_stack_tflops": 26.8e9 * len(crops) / t_par_net / 1e12,
            "parser_predict_ms_per_face": 1e3 * t_par_predict / max(1, len(crops)),
            "conv_stacks_images_per_sec": B / (t_det_net + t_par_net),
            "whole_path_images_per_sec": B / t_whole,
            "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}), flush=True)


if __name__ == "__main__":
    main()
