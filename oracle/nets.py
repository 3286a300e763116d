"""Oracle (test infrastructure): torch-CPU functional restatement of the three convnets.

Each function consumes a reference-format ``state_dict`` (same keys the reference's
``LoadMixin.load`` ingests, ``_layers.py:16-35``) and NCHW tensors, and follows the
reference forward passes op for op so that fp32 results agree with the reference modules
to rounding (pinned in ``tests/test_oracle_golden.py``).  ``dtype=torch.float64`` gives the
high-precision run used to measure the fp32 noise floor.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _cast(sd, dtype):
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}


def _bn(x, sd, p):
    # eval-mode BatchNorm2d, eps=1e-5 (nn.BatchNorm2d default; _layers.py:70,115,211 ...)
    return F.batch_norm(x, sd[f"{p}.running_mean"], sd[f"{p}.running_var"], sd[f"{p}.weight"], sd[f"{p}.bias"],
                        False, 0.0, 1e-5)


def _conv(x, sd, p, stride=1, padding=0):
    return F.conv2d(x, sd[f"{p}.weight"], sd.get(f"{p}.bias"), stride, padding)


# ------------------------------------------------------------------ RetinaFace
def _bottleneck(x, sd, p, stride):
    # torchvision resnet.py:143-163 (v1.5: the stride sits on the 3x3 conv)
    out = F.relu(_bn(_conv(x, sd, f"{p}.conv1"), sd, f"{p}.bn1"))
    out = F.relu(_bn(_conv(out, sd, f"{p}.conv2", stride, 1), sd, f"{p}.bn2"))
    out = _bn(_conv(out, sd, f"{p}.conv3"), sd, f"{p}.bn3")
    if f"{p}.downsample.0.weight" in sd:
        x = _bn(_conv(x, sd, f"{p}.downsample.0", stride), sd, f"{p}.downsample.1")
    return F.relu(out + x)


def retinaface_body(x, sd):
    """ResNet-50 body -> (C3, C4, C5).  retinaface.py:93-99,137; torchvision resnet.py:266-276."""
    x = F.relu(_bn(_conv(x, sd, "body.conv1", 2, 3), sd, "body.bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li, blocks in enumerate([3, 4, 6, 3], start=1):
        for b in range(blocks):
            x = _bottleneck(x, sd, f"body.layer{li}.{b}", 2 if (b == 0 and li > 1) else 1)
        if li > 1:
            feats.append(x)
    return feats


def _cbr(x, sd, p, padding, act=True):
    # conv + BN (+ LeakyReLU(0) == ReLU: leaky=0 because out_channels>64, _layers.py:68,103)
    y = _bn(_conv(x, sd, f"{p}.0", 1, padding), sd, f"{p}.1")
    return F.leaky_relu(y, 0.0) if act else y


def retinaface_fpn(feats, sd):
    """_layers.py:127-145."""
    o1, o2, o3 = (_cbr(f, sd, f"fpn.output{i}", 0) for i, f in enumerate(feats, start=1))
    o2 = _cbr(o2 + F.interpolate(o3, size=o2.shape[2:], mode="nearest"), sd, "fpn.merge2", 1)
    o1 = _cbr(o1 + F.interpolate(o2, size=o1.shape[2:], mode="nearest"), sd, "fpn.merge1", 1)
    return [o1, o2, o3]


def retinaface_ssh(x, sd, p):
    """_layers.py:90-97."""
    c3 = _cbr(x, sd, f"{p}.conv3X3", 1, act=False)
    c5_1 = _cbr(x, sd, f"{p}.conv5X5_1", 1)
    c5 = _cbr(c5_1, sd, f"{p}.conv5X5_2", 1, act=False)
    c7_2 = _cbr(c5_1, sd, f"{p}.conv7X7_2", 1)
    c7 = _cbr(c7_2, sd, f"{p}.conv7x7_3", 1, act=False)
    return F.relu(torch.cat([c3, c5, c7], 1))


def retinaface_heads_raw(x, sd, dtype=torch.float32):
    """Raw (pre-softmax) head outputs: (cls[N,A,2], box[N,A,4], ldm[N,A,10]).  retinaface.py:137-142."""
    sd = _cast(sd, dtype)
    fts = [retinaface_ssh(f, sd, f"ssh{i}") for i, f in
           enumerate(retinaface_fpn(retinaface_body(x.to(dtype), sd), sd), start=1)]
    outs = []
    for head, nout in (("ClassHead", 2), ("BboxHead", 4), ("LandmarkHead", 10)):
        per_level = []
        for i, f in enumerate(fts):
            y = _conv(f, sd, f"{head}.{i}.conv1x1").permute(0, 2, 3, 1).contiguous()  # _layers.py:153-157
            per_level.append(y.view(y.size(0), -1, nout))
        outs.append(torch.cat(per_level, 1))
    return tuple(outs)


def retinaface_forward(x, sd, dtype=torch.float32):
    """``RetinaFace.forward`` (retinaface.py:112-144): softmaxed scores, raw boxes, raw landmarks."""
    cls, box, ldm = retinaface_heads_raw(x, sd, dtype)
    return F.softmax(cls, dim=-1), box, ldm


def retinaface_preprocess(images_rgb_f32):
    """RGB->BGR flip and mean subtraction, retinaface.py:450-451 (offset is an int64 tensor, promoted)."""
    x = images_rgb_f32[:, [2, 1, 0]]
    return x - torch.tensor([104, 117, 123]).view(3, 1, 1)


# --------------------------------------------------------------------- BiSeNet
def _convbnrelu(x, sd, p, stride=1, padding=1):
    return F.relu(_bn(_conv(x, sd, f"{p}.conv", stride, padding), sd, f"{p}.bn"))  # _layers.py:279-283


def _basic_block(x, sd, p, stride):
    # _layers.py:226-239
    r = F.relu(_bn(_conv(x, sd, f"{p}.conv1", stride, 1), sd, f"{p}.bn1"))
    r = _bn(_conv(r, sd, f"{p}.conv2", 1, 1), sd, f"{p}.bn2")
    if f"{p}.downsample.0.weight" in sd:
        x = _bn(_conv(x, sd, f"{p}.downsample.0", stride), sd, f"{p}.downsample.1")
    return F.relu(x + r)


def _arm(x, sd, p):
    # _layers.py:305-313
    feat = _convbnrelu(x, sd, f"{p}.conv")
    att = F.avg_pool2d(feat, feat.shape[2:])
    att = torch.sigmoid(_bn(_conv(att, sd, f"{p}.conv_atten"), sd, f"{p}.bn_atten"))
    return feat * att


def bisenet_logits64(x, sd, dtype=torch.float32):
    """BiSeNet up to ``conv_out`` (19-channel logits at 1/8 resolution), bise.py:211 / _layers.py:326-368."""
    sd = _cast(sd, dtype)
    x = x.to(dtype)
    x = F.relu(_bn(_conv(x, sd, "cp.resnet.conv1", 2, 3), sd, "cp.resnet.bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li in range(1, 5):
        for b in range(2):
            x = _basic_block(x, sd, f"cp.resnet.layer{li}.{b}", 2 if (b == 0 and li > 1) else 1)
        feats.append(x)
    _, feat8, feat16, feat32 = feats
    avg = _convbnrelu(F.avg_pool2d(feat32, feat32.shape[2:]), sd, "cp.conv_avg", 1, 0)
    avg_up = F.interpolate(avg, feat32.shape[2:])
    feat32_sum = _arm(feat32, sd, "cp.arm32") + avg_up
    feat32_up = _convbnrelu(F.interpolate(feat32_sum, feat16.shape[2:]), sd, "cp.conv_head32")
    feat16_sum = _arm(feat16, sd, "cp.arm16") + feat32_up
    feat16_up = _convbnrelu(F.interpolate(feat16_sum, feat8.shape[2:]), sd, "cp.conv_head16")
    # FeatureFusionModule, _layers.py:357-368
    feat = _convbnrelu(torch.cat([feat8, feat16_up], 1), sd, "ffm.convblk", 1, 0)
    att = F.avg_pool2d(feat, feat.shape[2:])
    att = torch.sigmoid(_conv(F.relu(_conv(att, sd, "ffm.conv1")), sd, "ffm.conv2"))
    feat = feat * att + feat
    # BiSeNetOutput, _layers.py:291-295
    return _conv(_convbnrelu(feat, sd, "conv_out.conv"), sd, "conv_out.conv_out")


def bisenet_forward(x, sd, dtype=torch.float32):
    """``BiSeNet.forward`` (bise.py:195-212): logits bilinearly upsampled (align_corners=True) to the input size."""
    return F.interpolate(bisenet_logits64(x, sd, dtype), x.shape[2:], None, "bilinear", True)


# --------------------------------------------------------------------- RRDBNet
def _rdb(x, sd, p):
    # _layers.py:179-186
    lr = lambda t: F.leaky_relu(t, 0.2)
    x1 = lr(_conv(x, sd, f"{p}.conv1", 1, 1))
    x2 = lr(_conv(torch.cat((x, x1), 1), sd, f"{p}.conv2", 1, 1))
    x3 = lr(_conv(torch.cat((x, x1, x2), 1), sd, f"{p}.conv3", 1, 1))
    x4 = lr(_conv(torch.cat((x, x1, x2, x3), 1), sd, f"{p}.conv4", 1, 1))
    x5 = _conv(torch.cat((x, x1, x2, x3, x4), 1), sd, f"{p}.conv5", 1, 1)
    return x5 * 0.2 + x


def rrdbnet_forward(x, sd, dtype=torch.float32, nb=23):
    """``RRDBNet.forward`` (rrdb.py:64-81): (N,3,H,W) in [0,1] -> (N,3,4H,4W)."""
    sd = _cast(sd, dtype)
    lr = lambda t: F.leaky_relu(t, 0.2)
    first = _conv(x.to(dtype), sd, "conv_first", 1, 1)
    t = first
    for i in range(nb):
        out = t
        for r in (1, 2, 3):
            out = _rdb(out, sd, f"RRDB_trunk.{i}.RDB{r}")
        t = out * 0.2 + t                                    # _layers.py:200
    fea = first + _conv(t, sd, "trunk_conv", 1, 1)
    fea = lr(_conv(F.interpolate(fea, scale_factor=2), sd, "upconv1", 1, 1))
    fea = lr(_conv(F.interpolate(fea, scale_factor=2), sd, "upconv2", 1, 1))
    return _conv(lr(_conv(fea, sd, "HRconv", 1, 1)), sd, "conv_last", 1, 1)
