import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from face_crop_plus_b200 import _abi, synth
from oracle import pipeline, nets, detpost
torch.set_grad_enabled(False)
sd = synth.make_state_dict("retinaface", 0, class_bias=4.8)
ctx=_abi.Context(0); ctx.load_state_dict(_abi.MODEL_RETINAFACE, sd)
imgs = synth.make_images(2, 1024, 1024, seed=1234)
x = torch.from_numpy(imgs).permute(0,3,1,2).float()
cls, box, ldm = nets.retinaface_heads_raw(nets.retinaface_preprocess(x), sd)
ref = np.concatenate([cls.numpy(), box.numpy(), ldm.numpy()], -1)
cls64, box64, ldm64 = nets.retinaface_heads_raw(nets.retinaface_preprocess(x), sd, torch.float64)
ref64 = np.concatenate([cls64.numpy(), box64.numpy(), ldm64.numpy()], -1)
for impl in (1,0):
    ctx.set_conv_impl(impl)
    heads = ctx.detect_heads(imgs)
    print("impl",impl,"heads max|gpu-ref32|", np.abs(heads-ref).max(), " max|gpu-ref64|", np.abs(heads-ref64).max(), " max|ref32-ref64|", np.abs(ref-ref64).max())
    for strategy in ("all","largest"):
        l,i,a,b = detpost.detect_post(ref[...,:2], ref[...,2:6], ref[...,6:], 1024,1024,0.6,0.4,strategy)
        out = ctx.detect(imgs,0.6,0.4,strategy)
        sa, sb = set(zip(i,a)), set(zip(out["indices"].tolist(), out["anchors"].tolist()))
        print("  ",strategy,"ref faces",len(a),"gpu faces",len(out["anchors"]),"only ref",sorted(sa-sb)[:5],"only gpu",sorted(sb-sa)[:5])
        sc = detpost.softmax_face_score(ref[...,:2])
        for (im,an) in sorted(sa^sb)[:6]:
            print("      img",im,"anchor",an,"ref score",sc[im,an], "gpu logits", heads[im,an,:2], "ref logits", ref[im,an,:2])
        common = [k for k in zip(i,a) if k in sb]
        if common:
            gi = {k:j for j,k in enumerate(zip(out["indices"].tolist(), out["anchors"].tolist()))}
            ri = {k:j for j,k in enumerate(zip(i,a))}
            d = max(np.abs(out["landmarks"][gi[k]] - l[ri[k]]).max() for k in common)
            print("      max landmark diff on common faces", d)
