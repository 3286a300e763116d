"""Synthetic state_dicts and images for parity tests and benchmarks.

No trained checkpoint is reachable offline (the reference downloads
``retinaface_detector.pth`` / ``bise_parser.pth`` / ``bsrgan_x4_enhancer.pth``
in ``_layers.py:27-35``), so every number in this repo is produced with the
seeded, forward-free recipe below.  The tensors depend only on
``(seed, key)`` through a CPU ``torch.Generator``, hence they are identical in
the build container, on the GPU box, for the oracle and for the CUDA path.

The key lists restate the reference module trees:
  RetinaFace  retinaface.py:93-110 (+ torchvision resnet50 body up to layer4)
  BiSeNet     bise.py:191-193, _layers.py:206-368
  RRDBNet     rrdb.py:53-62, _layers.py:168-200
"""
from __future__ import annotations

import zlib

import numpy as np
import torch

# --------------------------------------------------------------------------
# parameter specs: list of (key_prefix, kind, shape-info)
#   ("conv", prefix, cout, cin, k, bias)      -> prefix.weight [, prefix.bias]
#   ("bn",   prefix, c)                       -> prefix.{weight,bias,running_mean,running_var,num_batches_tracked}
# --------------------------------------------------------------------------


def retinaface_spec():
    spec = [("conv", "body.conv1", 64, 3, 7, False), ("bn", "body.bn1", 64)]
    inplanes = 64
    for li, (planes, blocks) in enumerate([(64, 3), (128, 4), (256, 6), (512, 3)], start=1):
        for b in range(blocks):
            p = f"body.layer{li}.{b}"
            spec += [("conv", f"{p}.conv1", planes, inplanes, 1, False), ("bn", f"{p}.bn1", planes),
                     ("conv", f"{p}.conv2", planes, planes, 3, False), ("bn", f"{p}.bn2", planes),
                     ("conv", f"{p}.conv3", planes * 4, planes, 1, False), ("bn", f"{p}.bn3", planes * 4)]
            if b == 0:
                spec += [("conv", f"{p}.downsample.0", planes * 4, inplanes, 1, False),
                         ("bn", f"{p}.downsample.1", planes * 4)]
            inplanes = planes * 4
    for i, cin in enumerate([512, 1024, 2048], start=1):
        spec += [("conv", f"fpn.output{i}.0", 256, cin, 1, False), ("bn", f"fpn.output{i}.1", 256)]
    for m in ("merge1", "merge2"):
        spec += [("conv", f"fpn.{m}.0", 256, 256, 3, False), ("bn", f"fpn.{m}.1", 256)]
    for s in (1, 2, 3):
        for name, cout, cin in (("conv3X3", 128, 256), ("conv5X5_1", 64, 256), ("conv5X5_2", 64, 64),
                                ("conv7X7_2", 64, 64), ("conv7x7_3", 64, 64)):
            spec += [("conv", f"ssh{s}.{name}.0", cout, cin, 3, False), ("bn", f"ssh{s}.{name}.1", cout)]
    for head, nout in (("ClassHead", 2), ("BboxHead", 4), ("LandmarkHead", 10)):
        for i in range(3):
            spec.append(("conv", f"{head}.{i}.conv1x1", 2 * nout, 256, 1, True))
    return spec


def bisenet_spec():
    spec = [("conv", "cp.resnet.conv1", 64, 3, 7, False), ("bn", "cp.resnet.bn1", 64)]
    cin = 64
    for li, cout in enumerate([64, 128, 256, 512], start=1):
        for b in range(2):
            p = f"cp.resnet.layer{li}.{b}"
            spec += [("conv", f"{p}.conv1", cout, cin, 3, False), ("bn", f"{p}.bn1", cout),
                     ("conv", f"{p}.conv2", cout, cout, 3, False), ("bn", f"{p}.bn2", cout)]
            if b == 0 and li > 1:
                spec += [("conv", f"{p}.downsample.0", cout, cin, 1, False), ("bn", f"{p}.downsample.1", cout)]
            cin = cout
    for arm, c in (("arm16", 256), ("arm32", 512)):
        spec += [("conv", f"cp.{arm}.conv.conv", 128, c, 3, False), ("bn", f"cp.{arm}.conv.bn", 128),
                 ("conv", f"cp.{arm}.conv_atten", 128, 128, 1, False), ("bn", f"cp.{arm}.bn_atten", 128)]
    spec += [("conv", "cp.conv_head32.conv", 128, 128, 3, False), ("bn", "cp.conv_head32.bn", 128),
             ("conv", "cp.conv_head16.conv", 128, 128, 3, False), ("bn", "cp.conv_head16.bn", 128),
             ("conv", "cp.conv_avg.conv", 128, 512, 1, False), ("bn", "cp.conv_avg.bn", 128),
             ("conv", "ffm.convblk.conv", 256, 256, 1, False), ("bn", "ffm.convblk.bn", 256),
             ("conv", "ffm.conv1", 64, 256, 1, False), ("conv", "ffm.conv2", 256, 64, 1, False),
             ("conv", "conv_out.conv.conv", 256, 256, 3, False), ("bn", "conv_out.conv.bn", 256),
             ("conv", "conv_out.conv_out", 19, 256, 1, False)]
    return spec


def rrdbnet_spec(nb: int = 23, nf: int = 64, gc: int = 32):
    spec = [("conv", "conv_first", nf, 3, 3, True)]
    for i in range(nb):
        for r in (1, 2, 3):
            p = f"RRDB_trunk.{i}.RDB{r}"
            for k in range(1, 5):
                spec.append(("conv", f"{p}.conv{k}", gc, nf + (k - 1) * gc, 3, True))
            spec.append(("conv", f"{p}.conv5", nf, nf + 4 * gc, 3, True))
    for name, cout in (("trunk_conv", nf), ("upconv1", nf), ("upconv2", nf), ("HRconv", nf), ("conv_last", 3)):
        spec.append(("conv", name, cout, nf, 3, True))
    return spec


SPECS = {"retinaface": retinaface_spec, "bisenet": bisenet_spec, "rrdbnet": rrdbnet_spec}
#: file names the unmodified reference looks for under $TORCH_HOME/hub/checkpoints
REFERENCE_FILENAMES = {"retinaface": "retinaface_detector.pth", "bisenet": "bise_parser.pth",
                       "rrdbnet": "bsrgan_x4_enhancer.pth"}


def _gen(seed: int, key: str) -> torch.Generator:
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFFFFFFFFFF)


def _randn(seed, key, shape):
    return torch.randn(shape, generator=_gen(seed, key), dtype=torch.float32)


def _rand(seed, key, shape):
    return torch.rand(shape, generator=_gen(seed, key), dtype=torch.float32)


def make_state_dict(model: str, seed: int = 0, class_bias: float = 4.8) -> dict[str, torch.Tensor]:
    """Builds the seeded synthetic state_dict of ``model`` (same keys/shapes as the reference's)."""
    sd: dict[str, torch.Tensor] = {}
    for item in SPECS[model]():
        if item[0] == "conv":
            _, p, cout, cin, k, bias = item
            fan_in = cin * k * k
            if model == "rrdbnet":
                # torch's default Conv2d init range (kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +)), kept
                # because the x0.2 residual scaling makes the trunk contractive with it
                bound = 1.0 / fan_in ** 0.5
                sd[f"{p}.weight"] = (_rand(seed, f"{p}.weight", (cout, cin, k, k)) * 2 - 1) * bound
                sd[f"{p}.bias"] = (_rand(seed, f"{p}.bias", (cout,)) * 2 - 1) * bound
            else:
                sd[f"{p}.weight"] = _randn(seed, f"{p}.weight", (cout, cin, k, k)) * (2.0 / fan_in) ** 0.5
                if bias:
                    sd[f"{p}.bias"] = 0.1 * _randn(seed, f"{p}.bias", (cout,))
        else:
            _, p, c = item
            sd[f"{p}.weight"] = 1 + 0.1 * _randn(seed, f"{p}.weight", (c,))
            sd[f"{p}.bias"] = 0.1 * _randn(seed, f"{p}.bias", (c,))
            sd[f"{p}.running_mean"] = 0.1 * _randn(seed, f"{p}.running_mean", (c,))
            sd[f"{p}.running_var"] = 0.8 + 0.4 * _rand(seed, f"{p}.running_var", (c,))
            sd[f"{p}.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
            # keep the residual trunks contractive: damp the last BN of every residual branch
            if (model == "retinaface" and p.endswith(".bn3")) or \
               (model == "bisenet" and ".layer" in p and p.endswith(".bn2")):
                sd[f"{p}.weight"] *= 0.25
    if model == "retinaface":
        # stem BN absorbs the 0..255 mean-subtracted input scale
        sd["body.bn1.running_var"] *= 2 * 74.0 ** 2
        sd["body.bn1.running_mean"] *= 74.0
        for i in range(3):
            # the SSH features are post-ReLU (all positive), so make every head filter zero-sum over its input
            # channels: otherwise each head channel carries a large random constant that swamps the bias
            for head in ("ClassHead", "BboxHead", "LandmarkHead"):
                w = sd[f"{head}.{i}.conv1x1.weight"]
                sd[f"{head}.{i}.conv1x1.weight"] = w - w.mean(dim=1, keepdim=True)
            # (bg, face, bg, face) logits per the two anchors of a cell; sets the candidate density
            sd[f"ClassHead.{i}.conv1x1.bias"] = torch.tensor([class_bias, -class_bias] * 2, dtype=torch.float32)
    elif model == "bisenet":
        sd["cp.resnet.bn1.running_var"] *= 2
    elif model == "rrdbnet":
        # centre the output in [0,1] so the clamp/round tail is exercised on both sides
        sd["conv_last.bias"] = torch.full((3,), 0.5, dtype=torch.float32)
    return sd


def state_dict_digest(sd: dict[str, torch.Tensor]) -> str:
    """Order-independent crc32 digest (hex) of a state_dict, for pinning fixtures."""
    acc = 0
    for k in sorted(sd):
        acc = zlib.crc32(sd[k].contiguous().numpy().tobytes(), zlib.crc32(k.encode(), acc))
    return f"{acc:08x}"


def make_images(n: int, height: int = 1024, width: int = 1024, seed: int = 1234) -> np.ndarray:
    """``n`` synthetic RGB uint8 NHWC images: smooth low-frequency field + texture."""
    import torch.nn.functional as F
    out = np.empty((n, height, width, 3), dtype=np.uint8)
    for i in range(n):
        g = torch.Generator().manual_seed(seed + i)
        lo = torch.rand((1, 3, max(height // 16, 1), max(width // 16, 1)), generator=g)
        img = F.interpolate(lo, size=(height, width), mode="bicubic", align_corners=False) * 255
        img = img + 8 * torch.randn((1, 3, height, width), generator=g)
        out[i] = img.clamp(0, 255).round()[0].permute(1, 2, 0).to(torch.uint8).numpy()
    return out


def make_landmarks(n: int, size: int = 256, seed: int = 0) -> np.ndarray:
    """``n`` plausible 5-point landmark sets inside a ``size``x``size`` image (config C1)."""
    from .landmarks import STANDARD_LANDMARKS_5
    rng = np.random.default_rng(seed)
    out = np.empty((n, 5, 2), dtype=np.float32)
    for i in range(n):
        ang, sc = rng.uniform(-0.5, 0.5), rng.uniform(0.3, 0.8) * size
        rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        pts = (STANDARD_LANDMARKS_5.astype(np.float64) - 0.5) @ rot.T * sc + size / 2
        out[i] = (pts + rng.normal(0, 2, (5, 2))).astype(np.float32)
    return out
