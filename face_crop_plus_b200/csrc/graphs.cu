// graphs.cu — the three network graphs (RetinaFace, BiSeNet, RRDBNet) expressed over the fused conv kernel and the
// small kernels of misc.cu, plus the C-ABI entry points that run them (detect / align / parse / enhance / pipeline).
//
// Graph structure follows the reference modules op for op (file:line cited at each block); what changes is the
// execution: NHWC float32 activations, BatchNorm folded into the conv epilogue, concatenations written in place
// through channel-offset views, nearest upsampling read through index arithmetic, residual adds in epilogues.
#include <algorithm>
#include <functional>

#include "graphs.h"

namespace fcp {

static inline int odim(int in, int k, int s, int p) { return (in + 2 * p - k) / s + 1; }

// ------------------------------------------------------------------------------------------------------- Exec
Tensor Exec::alloc(int n, int h, int w, int c, int cs) {
    Tensor t;
    t.n = n; t.h = h; t.w = w; t.c = c; t.cs = cs ? cs : c; t.co = 0;
    if (!ok()) return t;
    t.p = static_cast<float*>(ctx->arena.alloc(t.pixels() * t.cs * sizeof(float)));
    if (!t.p) status = fail(ctx, FCP_ERR_CUDA, "activation arena exhausted");
    // FCP_POISON=1 (debugging): every activation tensor starts as 3.4e38s, so a read of memory its producer never wrote
    // (stale arena contents) shows up as a gross error instead of depending on what ran before
    static const bool poison = getenv("FCP_POISON") != nullptr;
    if (poison && t.p && !dry) cudaMemsetAsync(t.p, 0x7f, t.pixels() * t.cs * sizeof(float), ctx->stream);
    return t;
}
float* Exec::alloc_vec(size_t count) {
    if (!ok()) return nullptr;
    float* p = static_cast<float*>(ctx->arena.alloc(count * sizeof(float)));
    if (!p) status = fail(ctx, FCP_ERR_CUDA, "activation arena exhausted");
    return p;
}
void Exec::free(Tensor& t) {
    if (t.p && t.co == 0) ctx->arena.free(t.p);
    t.p = nullptr;
}
void Exec::free_vec(float* p) { ctx->arena.free(p); }

const ConvWeights* Exec::W(const std::string& name) {
    auto it = model->conv.find(name);
    if (it == model->conv.end()) {
        if (ok()) status = fail(ctx, FCP_ERR_STATE, "conv not finalized: " + name);
        return nullptr;
    }
    return &it->second;
}

bool Exec::conv(const std::string& name, Tensor in, Tensor out, int stride, int pad, int act, ConvOp extra) {
    if (!ok()) return false;
    const ConvWeights* wt = W(name);
    if (!wt) return false;
    ConvOp op = extra;
    op.in = in; op.out = out; op.wt = wt; op.stride = stride; op.pad = pad; op.act = act;
    op.impl = ctx->use_tc;
    if (op.up_in && op.impl >= 1 && in.c % 32 == 0) {
        // the tensor-core kernel needs a dense input for its TMA boxes: materialise the nearest x2 upsample (HBM-bound copy)
        Tensor up = alloc(in.n, in.h * 2, in.w * 2, in.c);
        if (!ok()) return false;
        op.in = up;
        op.up_in = 0;
        if (!dry) {
            status = launch_upsample2x(ctx, in, up);
            if (ok()) status = run_conv(ctx, op);
        }
        free(up);
        return ok();
    }
    if (dry) return true;
    status = run_conv(ctx, op);
    return ok();
}

// 7x7/s2 stem + folded BN + ReLU (`body.conv1`, `cp.resnet.conv1`).  Tensor-core route: one HBM-bound kernel gathers, for
// every input row, the 7 horizontal taps x 3 channels of each output column into a 32-channel fp32 tensor; the conv is
// then a 7x1 / stride (2,1) conv over it on the tcgen05 kernel (7 K-blocks; vertical padding = TMA zero fill).
// FCP_CONV_IMPL=0 keeps the CUDA-core stem kernel.
static void stem7(Exec& ex, const std::string& name, const void* src, int mode, int nb, int h, int w, Tensor out) {
    const ConvWeights* stem = ex.W(name);
    if (!stem || !ex.ok()) return;
    if (!ex.ctx->use_tc || getenv("FCP_STEM_FFMA")) {
        if (!ex.dry) ex.status = launch_stem7(ex.ctx, src, mode, nb, h, w, stem->w_kn, stem->scale, stem->shift, out);
        return;
    }
    // f16x3 mode, uint8 input (the detector): the converters read the image itself (one uint8 halo tile per output tile by TMA),
    // no row-patch tensor exists.  The plan still budgets the row-patch route: it is the fallback for image pitches TMA
    // cannot address (w*3 not a multiple of 16 bytes, unaligned batch).
    const bool direct = mode == 0 && ex.ctx->use_tc >= 2 && !getenv("FCP_STEM_ROWS") && ex.model->conv.count(name + ".direct") &&
                        !ex.dry && conv_tc_stem_supported(src, h, w);
    if (direct) {
        ConvOp e;
        e.stem_src = static_cast<const uint8_t*>(src); e.stem_h = h; e.stem_w = w;
        Tensor in;                                     // unused by the stem route; carries the batch size
        in.n = nb; in.h = h; in.w = w; in.c = in.cs = 64;
        ex.conv(name + ".direct", in, out, 1, 0, FCP_ACT_RELU, e);
        return;
    }
    Tensor rows = ex.alloc(nb, h, out.w, 32);
    if (ex.ok() && !ex.dry) ex.status = launch_stem_rows(ex.ctx, src, mode, nb, h, w, rows);
    ConvOp e;
    e.stride_w = 1; e.pad_w = 0;
    e.a_exact = mode == 0;      // u8 minus an integer mean: 9 significant bits, the split's low part is exactly zero
    ex.conv(name + ".rows", rows, out, 2, 3, FCP_ACT_RELU, e);
    ex.free(rows);
}

static ConvOp with_res1(Tensor r) {
    ConvOp e;
    e.res1 = r.p; e.res1_cs = r.cs; e.res1_co = r.co;
    return e;
}

// =========================================================================================== RetinaFace graph
int finalize_retinaface(fcp_ctx* ctx) {
    Model& m = ctx->models[FCP_MODEL_RETINAFACE];
    FCP_TRY(pack_conv(ctx, m, {"body.conv1"}, "body.bn1", "body.conv1"));
    FCP_TRY(pack_stem_rows(ctx, m, "body.conv1", "body.conv1.rows"));
    FCP_TRY(pack_stem_direct(ctx, m, "body.conv1", "body.conv1.direct"));
    const int blocks[4] = {3, 4, 6, 3};
    for (int li = 1; li <= 4; ++li)
        for (int b = 0; b < blocks[li - 1]; ++b) {
            std::string p = "body.layer" + std::to_string(li) + "." + std::to_string(b);
            for (int c = 1; c <= 3; ++c)
                FCP_TRY(pack_conv(ctx, m, {p + ".conv" + std::to_string(c)}, p + ".bn" + std::to_string(c), p + ".conv" + std::to_string(c)));
            if (b == 0) FCP_TRY(pack_conv(ctx, m, {p + ".downsample.0"}, p + ".downsample.1", p + ".downsample"));
        }
    // block 0 of every layer: relu(bn3(conv3(h)) + bn_d(conv_d(x))) is ONE 1x1 conv over the K concatenation [h | x(sampled at
    // the block's stride)], folded weights side by side, folded shifts added (tensor-core routes, `bottleneck`): the shortcut
    // tensor (4 x planes channels at the output resolution; 268 MB per 64 images in layer1) is neither written nor read back.
    for (int li = 1; li <= 4; ++li) {
        const std::string p = "body.layer" + std::to_string(li) + ".0";
        const int cout = 256 << (li - 1), cin3 = 64 << (li - 1), cind = li == 1 ? 64 : 128 << (li - 1);
        auto fold = [&](const std::string& conv, const std::string& bn, int cin, std::vector<float>& w, std::vector<float>& beta) -> bool {
            auto iw = m.host.find(conv + ".weight");
            auto g = m.host.find(bn + ".weight"), bb = m.host.find(bn + ".bias"), mu = m.host.find(bn + ".running_mean"),
                 var = m.host.find(bn + ".running_var");
            if (iw == m.host.end() || g == m.host.end() || bb == m.host.end() || mu == m.host.end() || var == m.host.end()) return false;
            const HostTensor& W = iw->second;
            if (W.shape.size() != 4 || W.shape[0] != cout || W.shape[1] != cin || W.shape[2] != 1 || W.shape[3] != 1) return false;
            w.resize((size_t)cout * cin); beta.resize(cout);
            for (int o = 0; o < cout; ++o) {
                const float invstd = 1.0f / std::sqrt(var->second.data[o] + 1e-5f);          // as pack_conv folds it
                const float alpha = g->second.data[o] * invstd;
                beta[o] = bb->second.data[o] - mu->second.data[o] * alpha;
                for (int c = 0; c < cin; ++c) w[(size_t)o * cin + c] = W.data[(size_t)o * cin + c] * alpha;
            }
            return true;
        };
        std::vector<float> w3, b3, wd, bd;
        if (!fold(p + ".conv3", p + ".bn3", cin3, w3, b3) || !fold(p + ".downsample.0", p + ".downsample.1", cind, wd, bd)) continue;
        HostTensor cat;
        const int K = cin3 + cind;
        cat.shape = {cout, K, 1, 1};
        cat.data.resize((size_t)cout * K);
        std::vector<float> one(cout, 1.f), shift(cout);
        for (int o = 0; o < cout; ++o) {
            for (int c = 0; c < cin3; ++c) cat.data[(size_t)o * K + c] = w3[(size_t)o * cin3 + c];
            for (int c = 0; c < cind; ++c) cat.data[(size_t)o * K + cin3 + c] = wd[(size_t)o * cind + c];
            shift[o] = b3[o] + bd[o];
        }
        m.host[p + ".conv3d.weight"] = std::move(cat);
        int st = pack_conv(ctx, m, {p + ".conv3d"}, "", p + ".conv3d", one.data(), shift.data());
        m.host.erase(p + ".conv3d.weight");
        FCP_TRY(st);
    }
    for (const char* n : {"fpn.output1", "fpn.output2", "fpn.output3", "fpn.merge1", "fpn.merge2"})
        FCP_TRY(pack_conv(ctx, m, {std::string(n) + ".0"}, std::string(n) + ".1", n));
    for (int s = 1; s <= 3; ++s)
        for (const char* n : {"conv3X3", "conv5X5_1", "conv5X5_2", "conv7X7_2", "conv7x7_3"}) {
            std::string p = "ssh" + std::to_string(s) + "." + n;
            FCP_TRY(pack_conv(ctx, m, {p + ".0"}, p + ".1", p));
        }
    for (int i = 0; i < 3; ++i) {
        std::string s = std::to_string(i);
        // the three 1x1 heads of a level share their input: one 256->32 conv = (cls 4 | box 8 | ldm 20) (_layers.py:147-162)
        FCP_TRY(pack_conv(ctx, m, {"ClassHead." + s + ".conv1x1", "BboxHead." + s + ".conv1x1", "LandmarkHead." + s + ".conv1x1"},
                          "", "heads." + s));
    }
    return FCP_OK;
}

// torchvision Bottleneck (resnet.py:143-163): 1x1 -> 3x3(stride) -> 1x1, BN each, residual add, ReLU
static Tensor bottleneck(Exec& ex, const std::string& p, Tensor x, int planes, int stride, bool down) {
    int ho = odim(x.h, 3, stride, 1), wo = odim(x.w, 3, stride, 1);
    Tensor t1 = ex.alloc(x.n, x.h, x.w, planes);
    ex.conv(p + ".conv1", x, t1, 1, 0, FCP_ACT_RELU);
    Tensor t2 = ex.alloc(x.n, ho, wo, planes);
    ex.conv(p + ".conv2", t1, t2, stride, 1, FCP_ACT_RELU);
    ex.free(t1);
    Tensor sc = x;
    // tensor-core routes: the shortcut conv is folded into conv3 (finalize_retinaface) - x is conv3's second K source
    // (maps under 64 pixels run on the CUDA-core kernel, which has no second source: conv_tc_supported)
    const bool folded = down && ex.ctx->use_tc >= 1 && (size_t)ho * wo >= 64 && ex.model->conv.count(p + ".conv3d") &&
                        !getenv("FCP_NO_FUSE_SHORTCUT");
    if (down && !folded) {
        sc = ex.alloc(x.n, ho, wo, planes * 4);
        ex.conv(p + ".downsample", x, sc, stride, 0, FCP_ACT_NONE);
    }
    Tensor out = ex.alloc(x.n, ho, wo, planes * 4);
    if (folded) {
        ConvOp e;
        e.in2 = x; e.in2_stride = stride;
        ex.conv(p + ".conv3d", t2, out, 1, 0, FCP_ACT_RELU, e);
        ex.free(t2);
        return out;
    }
    ex.conv(p + ".conv3", t2, out, 1, 0, FCP_ACT_RELU, with_res1(sc));
    ex.free(t2);
    if (down) ex.free(sc);
    return out;
}

// SSH context module (_layers.py:90-97); the final relu(cat(...)) is distributed into the three producing epilogues
static Tensor ssh(Exec& ex, const std::string& p, Tensor x) {
    Tensor out = ex.alloc(x.n, x.h, x.w, 256);
    ex.conv(p + ".conv3X3", x, out.slice(0, 128), 1, 1, FCP_ACT_RELU);
    Tensor c51 = ex.alloc(x.n, x.h, x.w, 64);
    ex.conv(p + ".conv5X5_1", x, c51, 1, 1, FCP_ACT_RELU);
    ex.conv(p + ".conv5X5_2", c51, out.slice(128, 64), 1, 1, FCP_ACT_RELU);
    Tensor c72 = ex.alloc(x.n, x.h, x.w, 64);
    ex.conv(p + ".conv7X7_2", c51, c72, 1, 1, FCP_ACT_RELU);
    ex.free(c51);
    ex.conv(p + ".conv7x7_3", c72, out.slice(192, 64), 1, 1, FCP_ACT_RELU);
    ex.free(c72);
    return out;
}

// RetinaFace.forward up to the raw head outputs (retinaface.py:112-142); lvl[i] = [nb, fh_i, fw_i, 32]
static int retinaface_forward(Exec& ex, const uint8_t* images, int nb, int h, int w, Tensor lvl[3]) {
    const ConvWeights* stem = ex.W("body.conv1");
    if (!stem) return ex.status;
    Tensor s1 = ex.alloc(nb, odim(h, 7, 2, 3), odim(w, 7, 2, 3), 64);
    stem7(ex, "body.conv1", images, 0, nb, h, w, s1);
    Tensor x = ex.alloc(nb, odim(s1.h, 3, 2, 1), odim(s1.w, 3, 2, 1), 64);
    if (ex.ok() && !ex.dry) ex.status = launch_maxpool3s2(ex.ctx, s1, x);
    ex.free(s1);
    const int blocks[4] = {3, 4, 6, 3}, planes[4] = {64, 128, 256, 512};
    Tensor feats[3];
    for (int li = 1; li <= 4; ++li) {
        for (int b = 0; b < blocks[li - 1]; ++b) {
            Tensor y = bottleneck(ex, "body.layer" + std::to_string(li) + "." + std::to_string(b), x, planes[li - 1],
                                  (b == 0 && li > 1) ? 2 : 1, b == 0);
            // the input of layer3.0 / layer4.0 is C3 / C4, still needed by the FPN; every other block input dies here
            bool x_is_feat = (li >= 3 && b == 0);
            if (!x_is_feat) ex.free(x);
            x = y;
        }
        if (li >= 2) feats[li - 2] = x;
    }
    // FPN (_layers.py:127-145): 1x1 lateral convs, top-down nearest upsample + add, 3x3 merge convs
    Tensor o3 = ex.alloc(nb, feats[2].h, feats[2].w, 256);
    ex.conv("fpn.output3", feats[2], o3, 1, 0, FCP_ACT_RELU);
    ex.free(feats[2]);
    Tensor o2 = ex.alloc(nb, feats[1].h, feats[1].w, 256);
    {
        ConvOp e;
        e.res2 = o3.p; e.res2_cs = o3.cs; e.res2_co = o3.co; e.res2_h = o3.h; e.res2_w = o3.w;
        ex.conv("fpn.output2", feats[1], o2, 1, 0, FCP_ACT_RELU, e);
    }
    ex.free(feats[1]);
    Tensor m2 = ex.alloc(nb, o2.h, o2.w, 256);
    ex.conv("fpn.merge2", o2, m2, 1, 1, FCP_ACT_RELU);
    ex.free(o2);
    Tensor o1 = ex.alloc(nb, feats[0].h, feats[0].w, 256);
    {
        ConvOp e;
        e.res2 = m2.p; e.res2_cs = m2.cs; e.res2_co = m2.co; e.res2_h = m2.h; e.res2_w = m2.w;
        ex.conv("fpn.output1", feats[0], o1, 1, 0, FCP_ACT_RELU, e);
    }
    ex.free(feats[0]);
    Tensor m1 = ex.alloc(nb, o1.h, o1.w, 256);
    ex.conv("fpn.merge1", o1, m1, 1, 1, FCP_ACT_RELU);
    ex.free(o1);
    Tensor fpn[3] = {m1, m2, o3};
    for (int i = 0; i < 3; ++i) {
        Tensor f = ssh(ex, "ssh" + std::to_string(i + 1), fpn[i]);
        ex.free(fpn[i]);
        lvl[i] = ex.alloc(nb, f.h, f.w, 32);
        ex.conv("heads." + std::to_string(i), f, lvl[i], 1, 0, FCP_ACT_NONE);
        ex.free(f);
    }
    return ex.status;
}

// ============================================================================================== BiSeNet graph
int finalize_bisenet(fcp_ctx* ctx) {
    Model& m = ctx->models[FCP_MODEL_BISENET];
    FCP_TRY(pack_conv(ctx, m, {"cp.resnet.conv1"}, "cp.resnet.bn1", "cp.resnet.conv1"));
    FCP_TRY(pack_stem_rows(ctx, m, "cp.resnet.conv1", "cp.resnet.conv1.rows"));
    for (int li = 1; li <= 4; ++li)
        for (int b = 0; b < 2; ++b) {
            std::string p = "cp.resnet.layer" + std::to_string(li) + "." + std::to_string(b);
            FCP_TRY(pack_conv(ctx, m, {p + ".conv1"}, p + ".bn1", p + ".conv1"));
            FCP_TRY(pack_conv(ctx, m, {p + ".conv2"}, p + ".bn2", p + ".conv2"));
            if (b == 0 && li > 1) FCP_TRY(pack_conv(ctx, m, {p + ".downsample.0"}, p + ".downsample.1", p + ".downsample"));
        }
    for (const char* a : {"cp.arm16", "cp.arm32"}) {
        FCP_TRY(pack_conv(ctx, m, {std::string(a) + ".conv.conv"}, std::string(a) + ".conv.bn", std::string(a) + ".conv"));
        FCP_TRY(pack_conv(ctx, m, {std::string(a) + ".conv_atten"}, std::string(a) + ".bn_atten", std::string(a) + ".atten"));
    }
    for (const char* c : {"cp.conv_head32", "cp.conv_head16", "cp.conv_avg", "ffm.convblk", "conv_out.conv"})
        FCP_TRY(pack_conv(ctx, m, {std::string(c) + ".conv"}, std::string(c) + ".bn", c));
    FCP_TRY(pack_conv(ctx, m, {"ffm.conv1"}, "", "ffm.conv1"));
    FCP_TRY(pack_conv(ctx, m, {"ffm.conv2"}, "", "ffm.conv2"));
    FCP_TRY(pack_conv(ctx, m, {"conv_out.conv_out"}, "", "conv_out.conv_out"));
    return FCP_OK;
}

// BasicBlock (_layers.py:226-239)
static Tensor basic_block(Exec& ex, const std::string& p, Tensor x, int cout, int stride, bool down, Tensor* out_view) {
    int ho = odim(x.h, 3, stride, 1), wo = odim(x.w, 3, stride, 1);
    Tensor t = ex.alloc(x.n, ho, wo, cout);
    ex.conv(p + ".conv1", x, t, stride, 1, FCP_ACT_RELU);
    Tensor sc = x;
    if (down) {
        sc = ex.alloc(x.n, ho, wo, cout);
        ex.conv(p + ".downsample", x, sc, stride, 0, FCP_ACT_NONE);
    }
    Tensor out = out_view ? *out_view : ex.alloc(x.n, ho, wo, cout);
    ex.conv(p + ".conv2", t, out, 1, 1, FCP_ACT_RELU, with_res1(sc));
    ex.free(t);
    if (down) ex.free(sc);
    return out;
}

// AttentionRefinementModule (_layers.py:305-313) followed by "+ addvec" (ContextPath, _layers.py:336)
static Tensor arm(Exec& ex, const std::string& p, Tensor x, const float* addvec) {
    Tensor feat = ex.alloc(x.n, x.h, x.w, 128);
    ex.conv(p + ".conv", x, feat, 1, 1, FCP_ACT_RELU);
    float* pooled = ex.alloc_vec((size_t)x.n * 128);
    float* att = ex.alloc_vec((size_t)x.n * 128);
    const ConvWeights* wa = ex.W(p + ".atten");
    if (ex.ok() && !ex.dry) ex.status = launch_global_avgpool(ex.ctx, feat, pooled);
    if (ex.ok() && !ex.dry) ex.status = launch_fc(ex.ctx, pooled, x.n, 128, wa, FCP_ACT_SIGMOID, att);
    Tensor out = ex.alloc(x.n, x.h, x.w, 128);
    if (ex.ok() && !ex.dry) ex.status = launch_channel_affine(ex.ctx, feat, att, addvec, 0, out);
    ex.free(feat);
    ex.free_vec(pooled);
    ex.free_vec(att);
    return out;
}

// BiSeNet.forward up to conv_out (bise.py:211; _layers.py:326-368); logits = [nb, 64, 64, 19] with channel stride 32
static int bisenet_forward(Exec& ex, const float* in3, int nb, Tensor& logits) {
    const ConvWeights* stem = ex.W("cp.resnet.conv1");
    if (!stem) return ex.status;
    Tensor s1 = ex.alloc(nb, 256, 256, 64);
    stem7(ex, "cp.resnet.conv1", in3, 1, nb, 512, 512, s1);
    Tensor x = ex.alloc(nb, 128, 128, 64);
    if (ex.ok() && !ex.dry) ex.status = launch_maxpool3s2(ex.ctx, s1, x);
    ex.free(s1);
    // the FFM concat buffer: [feat8 | feat16_up] (_layers.py:358)
    Tensor cat = ex.alloc(nb, 64, 64, 256);
    Tensor feat8v = cat.slice(0, 128), feat16upv = cat.slice(128, 128);
    const int couts[4] = {64, 128, 256, 512};
    Tensor feat16, feat32;
    for (int li = 1; li <= 4; ++li) {
        std::string p = "cp.resnet.layer" + std::to_string(li);
        Tensor y = basic_block(ex, p + ".0", x, couts[li - 1], li > 1 ? 2 : 1, li > 1, nullptr);
        if (li != 3 && li != 4) ex.free(x);          // feat8 lives in `cat`; feat16 is freed later
        else if (li == 3) { /* x == feat8 view inside cat: keep */ }
        Tensor z = basic_block(ex, p + ".1", y, couts[li - 1], 1, false, li == 2 ? &feat8v : nullptr);
        ex.free(y);
        x = z;
        if (li == 3) feat16 = z;
        if (li == 4) feat32 = z;
    }
    // ContextPath (_layers.py:326-346)
    float* pooled = ex.alloc_vec((size_t)nb * 512);
    float* avgv = ex.alloc_vec((size_t)nb * 128);
    if (ex.ok() && !ex.dry) ex.status = launch_global_avgpool(ex.ctx, feat32, pooled);
    if (ex.ok() && !ex.dry) ex.status = launch_fc(ex.ctx, pooled, nb, 512, ex.W("cp.conv_avg"), FCP_ACT_RELU, avgv);
    Tensor feat32_sum = arm(ex, "cp.arm32", feat32, avgv);          // arm32(feat32) + avg_up (broadcast of a 1x1 map)
    ex.free(feat32);
    ex.free_vec(pooled);
    Tensor feat16_arm = arm(ex, "cp.arm16", feat16, nullptr);
    ex.free(feat16);
    Tensor feat16_sum = ex.alloc(nb, 32, 32, 128);
    {
        ConvOp e;                                                   // conv_head32(up(feat32_sum)) + feat16_arm
        e.up_in = 1;
        e.res2 = feat16_arm.p; e.res2_cs = feat16_arm.cs; e.res2_co = feat16_arm.co;
        ex.conv("cp.conv_head32", feat32_sum, feat16_sum, 1, 1, FCP_ACT_RELU, e);
    }
    ex.free(feat32_sum);
    ex.free(feat16_arm);
    ex.free_vec(avgv);
    {
        ConvOp e;
        e.up_in = 1;
        ex.conv("cp.conv_head16", feat16_sum, feat16upv, 1, 1, FCP_ACT_RELU, e);
    }
    ex.free(feat16_sum);
    // FeatureFusionModule (_layers.py:357-368)
    Tensor feat = ex.alloc(nb, 64, 64, 256);
    ex.conv("ffm.convblk", cat, feat, 1, 0, FCP_ACT_RELU);
    ex.free(cat);
    float* p1 = ex.alloc_vec((size_t)nb * 256);
    float* p2 = ex.alloc_vec((size_t)nb * 64);
    float* p3 = ex.alloc_vec((size_t)nb * 256);
    if (ex.ok() && !ex.dry) ex.status = launch_global_avgpool(ex.ctx, feat, p1);
    if (ex.ok() && !ex.dry) ex.status = launch_fc(ex.ctx, p1, nb, 256, ex.W("ffm.conv1"), FCP_ACT_RELU, p2);
    if (ex.ok() && !ex.dry) ex.status = launch_fc(ex.ctx, p2, nb, 64, ex.W("ffm.conv2"), FCP_ACT_SIGMOID, p3);
    Tensor fused = ex.alloc(nb, 64, 64, 256);
    if (ex.ok() && !ex.dry) ex.status = launch_channel_affine(ex.ctx, feat, p3, nullptr, 1, fused);   // feat*att + feat
    ex.free(feat);
    ex.free_vec(p1); ex.free_vec(p2); ex.free_vec(p3);
    // BiSeNetOutput (_layers.py:291-295)
    Tensor mid = ex.alloc(nb, 64, 64, 256);
    ex.conv("conv_out.conv", fused, mid, 1, 1, FCP_ACT_RELU);
    ex.free(fused);
    logits = ex.alloc(nb, 64, 64, 19, 32);
    ex.conv("conv_out.conv_out", mid, logits, 1, 0, FCP_ACT_NONE);
    ex.free(mid);
    return ex.status;
}

// ============================================================================================== RRDBNet graph
int finalize_rrdbnet(fcp_ctx* ctx) {
    Model& m = ctx->models[FCP_MODEL_RRDBNET];
    FCP_TRY(pack_conv(ctx, m, {"conv_first"}, "", "conv_first"));
    for (int i = 0; i < m.rrdb_blocks; ++i)
        for (int r = 1; r <= 3; ++r)
            for (int c = 1; c <= 5; ++c) {
                std::string p = "RRDB_trunk." + std::to_string(i) + ".RDB" + std::to_string(r) + ".conv" + std::to_string(c);
                FCP_TRY(pack_conv(ctx, m, {p}, "", p));
            }
    for (const char* c : {"trunk_conv", "upconv1", "upconv2", "HRconv", "conv_last"}) FCP_TRY(pack_conv(ctx, m, {c}, "", c));
    // ---- upconv1 / upconv2 on a nearest x2 upsampled input (rrdb.py:78-79): output pixel (2y+py, 2x+px) only ever sees the
    // low-resolution pixels (y-1+py .. y+py) x (x-1+px .. x+px), because pairs of the 3x3 taps land on the same source pixel.
    // Per output parity class the conv is therefore a 2x2 conv on the LOW-resolution tensor whose weights are sums of the
    // original taps (rows: py = 0 -> {w0, w1 + w2}, py = 1 -> {w0 + w1, w2}; columns alike): 4/9 of the MMAs and no
    // materialised upsample.  rrdbnet_forward runs the four classes as four launches into strided views of the output.
    for (const char* name : {"upconv1", "upconv2"}) {
        auto iw = m.host.find(std::string(name) + ".weight");
        auto ib = m.host.find(std::string(name) + ".bias");
        if (iw == m.host.end() || iw->second.shape.size() != 4 || iw->second.shape[0] != 64 || iw->second.shape[1] != 64 ||
            iw->second.shape[2] != 3 || iw->second.shape[3] != 3)
            continue;
        const std::vector<float>& W = iw->second.data;
        std::vector<float> one(64, 1.f), bias(64, 0.f);
        if (ib != m.host.end())
            for (int o = 0; o < 64; ++o) bias[o] = ib->second.data[o];
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                HostTensor w2;
                w2.shape = {64, 64, 2, 2};
                w2.data.assign((size_t)64 * 64 * 4, 0.f);
                for (int o = 0; o < 64; ++o)
                    for (int c = 0; c < 64; ++c)
                        for (int r = 0; r < 3; ++r)
                            for (int t = 0; t < 3; ++t) {
                                const int r2 = (r + 1 - py) >> 1, t2 = (t + 1 - px) >> 1;      // the 2x2 tap the 3x3 tap falls on
                                w2.data[(((size_t)o * 64 + c) * 2 + r2) * 2 + t2] += W[(((size_t)o * 64 + c) * 3 + r) * 3 + t];
                            }
                const std::string full = std::string(name) + ".p" + std::to_string(py) + std::to_string(px);
                m.host[full + ".weight"] = std::move(w2);
                int st = pack_conv(ctx, m, {full}, "", full, one.data(), bias.data());
                m.host.erase(full + ".weight");
                FCP_TRY(st);
                m.conv[full].alg_k = 9 * 64;            // FLOP accounting: a quarter of the reference's 3x3 conv per class
            }
    }
    // ---- conv_last (64 -> 3) for the direct fp32 kernel (misc.cu conv3_last_kernel): [tap][cin][cout] + bias
    {
        auto iw = m.host.find("conv_last.weight");
        auto ib = m.host.find("conv_last.bias");
        if (iw != m.host.end() && iw->second.shape.size() == 4 && iw->second.shape[0] == 3 && iw->second.shape[1] == 64 &&
            iw->second.shape[2] == 3 && iw->second.shape[3] == 3) {
            std::vector<float> pk(9 * 64 * 3 + 4, 0.f);
            for (int o = 0; o < 3; ++o)
                for (int c = 0; c < 64; ++c)
                    for (int t = 0; t < 9; ++t) pk[(t * 64 + c) * 3 + o] = iw->second.data[((size_t)o * 64 + c) * 9 + t];
            if (ib != m.host.end())
                for (int o = 0; o < 3; ++o) pk[9 * 64 * 3 + o] = ib->second.data[o];
            m.vec["conv_last.direct"] = pk;
        }
    }
    // ---- source-major packing of the dense blocks (rrdbnet_forward, tensor-core route).  ResidualDenseBlock_5C
    // (_layers.py:179-186) is conv_k([x, x1 .. x_{k-1}]) for k = 1..5: conv_k's weight splits by input channel range into one
    // block per SOURCE x_j, and one pass per source computes that source's contribution to every later conv at once.
    //   s0a: x  -> [c1 c2 c3 c4]          (128)   bias 1-4, c1 final (leaky)
    //   s0b: x  -> [c5]                   ( 64)   0.2 * (W5 x + b5), + x in the epilogue
    //   s1a: x1 -> [c2 c3 c4 c5[0:32]]    (128)   accumulate, c2 final        s1b: x1 -> c5[32:64] (32)
    //   s2 : x2 -> [c3 c4 c5]             (128)   accumulate, c3 final
    //   s3 : x3 -> [c4 c5]                ( 96)   accumulate, c4 final
    //   s4 : x4 -> next x                 ( 64)   + c5 = x + 0.2 * conv5(..)   (_layers.py:186)
    // conv5's 0.2 is folded into its weight blocks (pack_conv's per-channel scale), so its partial sums accumulate pre-scaled.
    struct Part { int conv, o0, o1; float scale; };
    auto pack_source = [&](const std::string& prefix, const std::string& name, int c0, int c1, const std::vector<Part>& parts,
                           bool with_bias) -> int {
        HostTensor wt;
        std::vector<float> scale, shift;
        const int cn = c1 - c0;
        for (const Part& pt : parts) {
            const HostTensor* w = nullptr;
            const HostTensor* b = nullptr;
            {
                const std::string cname = prefix + ".conv" + std::to_string(pt.conv);
                auto iw = m.host.find(cname + ".weight");
                auto ib = m.host.find(cname + ".bias");
                if (iw == m.host.end() || iw->second.shape.size() != 4 || iw->second.shape[1] < c1 || iw->second.shape[0] < pt.o1 ||
                    iw->second.shape[2] != 3 || iw->second.shape[3] != 3)
                    return fail(ctx, FCP_ERR_STATE, "missing or mis-shaped dense-block weight: " + cname);
                w = &iw->second;
                if (ib != m.host.end()) b = &ib->second;
            }
            const int cin = (int)w->shape[1];
            for (int o = pt.o0; o < pt.o1; ++o) {
                for (int c = c0; c < c1; ++c)
                    for (int t = 0; t < 9; ++t) wt.data.push_back(w->data[((size_t)o * cin + c) * 9 + t]);
                scale.push_back(pt.scale);
                shift.push_back(with_bias && b ? b->data[o] * pt.scale : 0.f);
            }
        }
        wt.shape = {(int64_t)scale.size(), cn, 3, 3};
        const std::string full = prefix + "." + name;
        m.host[full + ".weight"] = std::move(wt);
        int st = pack_conv(ctx, m, {full}, "", full, scale.data(), shift.data());
        m.host.erase(full + ".weight");
        return st;
    };
    for (int i = 0; i < m.rrdb_blocks; ++i)
        for (int r = 1; r <= 3; ++r) {
            const std::string p = "RRDB_trunk." + std::to_string(i) + ".RDB" + std::to_string(r);
            FCP_TRY(pack_source(p, "s0a", 0, 64, {{1, 0, 32, 1.f}, {2, 0, 32, 1.f}, {3, 0, 32, 1.f}, {4, 0, 32, 1.f}}, true));
            FCP_TRY(pack_source(p, "s0b", 0, 64, {{5, 0, 64, 0.2f}}, true));
            FCP_TRY(pack_source(p, "s1a", 64, 96, {{2, 0, 32, 1.f}, {3, 0, 32, 1.f}, {4, 0, 32, 1.f}, {5, 0, 32, 0.2f}}, false));
            FCP_TRY(pack_source(p, "s1b", 64, 96, {{5, 32, 64, 0.2f}}, false));
            FCP_TRY(pack_source(p, "s2", 96, 128, {{3, 0, 32, 1.f}, {4, 0, 32, 1.f}, {5, 0, 64, 0.2f}}, false));
            FCP_TRY(pack_source(p, "s3", 128, 160, {{4, 0, 32, 1.f}, {5, 0, 64, 0.2f}}, false));
            FCP_TRY(pack_source(p, "s4", 160, 192, {{5, 0, 64, 0.2f}}, false));
        }
    return FCP_OK;
}

// RRDBNet.forward (rrdb.py:64-81) on nb images; `fill_first(first)` runs conv_first of the nb inputs into
// first = [nb,h,w,64] (f32 NCHW / in_div or u8 NHWC / 255 sources); result x4 = [nb,4h,4w,3] (cs 4)
static int rrdbnet_forward(Exec& ex, const std::function<int(Tensor)>& fill_first, int nb, int h, int w, Tensor& x4) {
    const ConvWeights* cf = ex.W("conv_first");
    if (!cf) return ex.status;
    Tensor first = ex.alloc(nb, h, w, 64);
    if (ex.ok() && !ex.dry) ex.status = fill_first(first);
    // three rotating slabs [x | x1 | x2 | x3 | x4 (| c5)]: the dense concatenations of ResidualDenseBlock_5C
    // (_layers.py:179-186) are channel-prefix views of one slab.  Tensor-core f16 route: source-major passes (see
    // finalize_rrdbnet) - every x_j tile is read by ONE or two launches instead of by each of the 5 - j later convs, the
    // partial sums of the not-yet-final convs accumulate in place in the slab (fp32, residual path of the conv epilogue).
    // FCP_RRDB_LAYERWISE=1 keeps the conv-by-conv schedule (also the schedule of the 3xTF32 and CUDA-core routes).
    const bool source_major = ex.ctx->use_tc >= 2 && !getenv("FCP_RRDB_LAYERWISE") &&
                              (size_t)h * w >= 64;        // smaller maps run on the CUDA-core kernel (conv_tc_supported)
    const int SC = source_major ? 256 : 192;
    Tensor slab[3];
    for (auto& s : slab) s = ex.alloc(nb, h, w, SC);
    if (ex.ok() && !ex.dry)
        ex.status = cudaMemcpy2DAsync(slab[0].p, SC * sizeof(float), first.p, 64 * sizeof(float), 64 * sizeof(float),
                                      first.pixels(), cudaMemcpyDeviceToDevice, ex.ctx->stream) == cudaSuccess
                        ? FCP_OK : fail(ex.ctx, FCP_ERR_CUDA, "slab init copy failed");
    int cur = 0;
    for (int i = 0; i < ex.model->rrdb_blocks && ex.ok(); ++i) {
        // source-major: the third dense block ADDS 0.2 * its output onto the RRDB input where it lies (slab[cur].x, TMA
        // reduce-add store) - no second residual in the epilogue, and the next RRDB starts from the same slab
        const int order_in[3] = {cur, (cur + 1) % 3, (cur + 2) % 3};
        const int order_out[3] = {(cur + 1) % 3, (cur + 2) % 3, source_major ? cur : (cur + 1) % 3};
        for (int r = 0; r < 3; ++r) {
            std::string p = "RRDB_trunk." + std::to_string(i) + ".RDB" + std::to_string(r + 1);
            Tensor S = slab[order_in[r]], T = slab[order_out[r]];
            if (source_major) {
                auto acc_into = [&](Tensor dst, bool in_place, int act_cols) {       // (+ dst) -> dst, leaky on the first act_cols channels
                    ConvOp o;
                    o.slope = 0.2f;
                    o.act_cols = act_cols;
                    if (in_place) { o.res1 = dst.p; o.res1_cs = dst.cs; o.res1_co = dst.co; }
                    return o;
                };
                ex.conv(p + ".s0a", S.slice(0, 64), S.slice(64, 128), 1, 1, FCP_ACT_LRELU, acc_into(S, false, 32));
                ConvOp b0;                                                            // c5 = x + 0.2 * (W5[:, x] x + b5)
                b0.res1 = S.p; b0.res1_cs = S.cs; b0.res1_co = S.co;
                ex.conv(p + ".s0b", S.slice(0, 64), S.slice(192, 64), 1, 1, FCP_ACT_NONE, b0);
                ex.conv(p + ".s1a", S.slice(64, 32), S.slice(96, 128), 1, 1, FCP_ACT_LRELU, acc_into(S.slice(96, 128), true, 32));
                ex.conv(p + ".s1b", S.slice(64, 32), S.slice(224, 32), 1, 1, FCP_ACT_NONE, acc_into(S.slice(224, 32), true, 1 << 30));
                ex.conv(p + ".s2", S.slice(96, 32), S.slice(128, 128), 1, 1, FCP_ACT_LRELU, acc_into(S.slice(128, 128), true, 32));
                ex.conv(p + ".s3", S.slice(128, 32), S.slice(160, 96), 1, 1, FCP_ACT_LRELU, acc_into(S.slice(160, 96), true, 32));
                ConvOp e = acc_into(S.slice(192, 64), true, 1 << 30);                      // next x = c5 + 0.2 * W5[:, x4] x4
                if (r == 2) {                                                         // RRDB: out * 0.2 + x   (_layers.py:200)
                    e.post_scale = 0.2f;                                              // T.x (= the RRDB input) += 0.2 * (c5 + W5[:, x4] x4)
                    e.out_add = 1;
                }
                ex.conv(p + ".s4", S.slice(160, 32), T.slice(0, 64), 1, 1, FCP_ACT_NONE, e);
                continue;
            }
            ConvOp lre;
            lre.slope = 0.2f;
            for (int c = 1; c <= 4; ++c)
                ex.conv(p + ".conv" + std::to_string(c), S.slice(0, 64 + 32 * (c - 1)), S.slice(32 + 32 * c, 32), 1, 1,
                        FCP_ACT_LRELU, lre);
            ConvOp e;                                   // x5 * 0.2 + x   (_layers.py:186)
            e.post_scale = 0.2f;
            e.res2 = S.p; e.res2_cs = S.cs; e.res2_co = 0;
            if (r == 2) {                               // RRDB: out * 0.2 + x   (_layers.py:200)
                e.post_scale2 = 0.2f;
                e.res3 = slab[cur].p; e.res3_cs = SC; e.res3_co = 0;
            }
            ex.conv(p + ".conv5", S, T.slice(0, 64), 1, 1, FCP_ACT_NONE, e);
        }
        if (!source_major) cur = (cur + 1) % 3;
    }
    Tensor fea = ex.alloc(nb, h, w, 64);
    ex.conv("trunk_conv", slab[cur].slice(0, 64), fea, 1, 1, FCP_ACT_NONE, with_res1(first));   // first + trunk_conv(..)
    for (auto& s : slab) ex.free(s);
    ex.free(first);
    ConvOp up;
    up.up_in = 1; up.slope = 0.2f;
    // tensor-core routes: four 2x2 convs on the low-resolution input, one per output-pixel parity class (finalize_rrdbnet)
    auto upconv = [&](const std::string& name, Tensor lo, Tensor hi) {
        const bool classes = ex.ctx->use_tc >= 1 && (size_t)lo.h * lo.w >= 64 && ex.model->conv.count(name + ".p00") &&
                             !getenv("FCP_UPCONV_MATERIALIZE");
        if (!classes) { ex.conv(name, lo, hi, 1, 1, FCP_ACT_LRELU, up); return; }
        for (int py = 0; py < 2; ++py)
            for (int px = 0; px < 2; ++px) {
                ConvOp o;
                o.slope = 0.2f;
                o.pad_w = px == 0 ? 1 : 0;                                      // taps (y - 1 + py .. y + py) x (x - 1 + px .. x + px)
                Tensor view = hi;                                               // output pixel (y, x) of the class -> (2y + py, 2x + px)
                view.h = lo.h; view.w = lo.w;
                view.cs = 2 * hi.cs;
                view.co = hi.co + (py * hi.w + px) * hi.cs;
                o.out_rs = 2 * hi.w * hi.cs;
                o.out_is = (size_t)hi.h * hi.w * hi.cs;
                ex.conv(name + ".p" + std::to_string(py) + std::to_string(px), lo, view, 1, py == 0 ? 1 : 0, FCP_ACT_LRELU, o);
            }
    };
    Tensor u1 = ex.alloc(nb, 2 * h, 2 * w, 64);
    upconv("upconv1", fea, u1);
    ex.free(fea);
    Tensor u2 = ex.alloc(nb, 4 * h, 4 * w, 64);
    upconv("upconv2", u1, u2);
    ex.free(u1);
    ConvOp lre;
    lre.slope = 0.2f;
    Tensor hr = ex.alloc(nb, 4 * h, 4 * w, 64);
    ex.conv("HRconv", u2, hr, 1, 1, FCP_ACT_LRELU, lre);
    ex.free(u2);
    x4 = ex.alloc(nb, 4 * h, 4 * w, 3, 4);
    auto direct = ex.model->vec.find("conv_last.direct");
    if (ex.ctx->use_tc >= 1 && direct != ex.model->vec.end() && !getenv("FCP_CONV_LAST_GEMM")) {
        if (ex.ok() && !ex.dry) ex.status = launch_conv3_last(ex.ctx, hr, direct->second.data(), x4);   // 3 output channels: direct fp32 conv
    } else {
        ex.conv("conv_last", hr, x4, 1, 1, FCP_ACT_NONE);
    }
    ex.free(hr);
    return ex.status;
}

// ------------------------------------------------------------------------------------------ plan + reserve
static int plan_reserve(fcp_ctx* ctx, Model* model, const std::vector<std::function<int(Exec&)>>& graphs) {
    size_t need = 0;
    for (auto& g : graphs) {
        ctx->arena.set_plan_mode(true);
        Exec ex{ctx, model, true};
        int s = g(ex);
        size_t hw = ctx->arena.high_water();
        ctx->arena.set_plan_mode(false);
        if (s != FCP_OK) return s;
        need = std::max(need, hw);
    }
    if (!ctx->arena.reserve(std::max(need, ctx->arena.capacity())))
        return fail(ctx, FCP_ERR_CUDA, "cannot reserve activation arena of " + std::to_string(need >> 20) + " MiB");
    ctx->arena.reset();
    return FCP_OK;
}

static int need_model(fcp_ctx* ctx, int model) {
    if (!ctx) return FCP_ERR_INVALID;
    if (!ctx->models[model].finalized) return fail(ctx, FCP_ERR_STATE, "model weights not finalized (fcp_load_tensor + fcp_finalize)");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, FCP_ERR_CUDA, "cudaSetDevice failed");
    return FCP_OK;
}

// ============================================================================================== detect core
// Runs the detector over micro-batches; faces (device) receives up to max_faces 16-float records in image order.
static int detect_core(fcp_ctx* ctx, const uint8_t* images, int n, int h, int w, float vis, float nms, int strategy,
                       int max_faces, float* faces, int32_t* face_img, int32_t* face_count, float* heads_out) {
    Model* model = &ctx->models[FCP_MODEL_RETINAFACE];
    const int mb = std::min(ctx->det_mb, n);
    // micro-batch schedule: uniform, or the one fcp_pipeline installed for host images (a short first micro-batch, so that
    // the only exposed H2D copy is small)
    std::vector<int> sched;
    {
        int total = 0;
        for (int v : ctx->mb_sched) total += v;
        if (!ctx->mb_sched.empty() && total == n) sched = ctx->mb_sched;
        else for (int b0 = 0; b0 < n; b0 += mb) sched.push_back(std::min(mb, n - b0));
    }
    std::vector<std::function<int(Exec&)>> plans;
    {
        std::vector<int> sizes(sched);
        std::sort(sizes.begin(), sizes.end());
        sizes.erase(std::unique(sizes.begin(), sizes.end()), sizes.end());
        for (int nb : sizes) plans.push_back([=](Exec& ex) { Tensor l[3]; return retinaface_forward(ex, nullptr, nb, h, w, l); });
    }
    FCP_TRY(plan_reserve(ctx, model, plans));
    const int A = det_num_priors(h, w), key_cap = det_key_capacity(h, w);
    float* rec = nullptr; unsigned long long* keys = nullptr; unsigned char* supp = nullptr; int32_t* counts = nullptr;
    if (!heads_out) {
        FCP_CUDA(ctx, cudaMallocAsync(&rec, (size_t)mb * A * 16 * sizeof(float), ctx->stream));
        FCP_CUDA(ctx, cudaMallocAsync(&keys, (size_t)mb * key_cap * sizeof(unsigned long long), ctx->stream));
        FCP_CUDA(ctx, cudaMallocAsync(&supp, (size_t)mb * key_cap, ctx->stream));
        FCP_CUDA(ctx, cudaMallocAsync(&counts, sizeof(int32_t) * 2 * mb, ctx->stream));
        FCP_CUDA(ctx, cudaMemsetAsync(face_count, 0, sizeof(int32_t), ctx->stream));
    }
    int status = FCP_OK;
    int b0 = 0;
    for (size_t mi = 0; mi < sched.size() && status == FCP_OK; b0 += sched[mi], ++mi) {
        const int nb = sched[mi];
        if (ctx->on_microbatch && (status = ctx->on_microbatch((int)mi)) != FCP_OK) break;
        ctx->arena.reset();
        Exec ex{ctx, model, false};
        Tensor lvl[3];
        {
            StageScope st(ctx, ST_DETECT_NET);
            status = retinaface_forward(ex, images + (size_t)b0 * h * w * 3, nb, h, w, lvl);
        }
        if (status != FCP_OK) break;
        const float* ptrs[3] = {lvl[0].p, lvl[1].p, lvl[2].p};
        StageScope st(ctx, ST_DETECT_POST);
        if (heads_out) status = launch_heads_to_flat(ctx, ptrs, nb, h, w, heads_out + (size_t)b0 * A * 16);
        else status = launch_det_post(ctx, ptrs, nullptr, nb, b0, h, w, vis, nms, strategy, max_faces, rec, keys, supp,
                                      counts, counts + mb, faces, face_img, face_count);
    }
    if (rec) cudaFreeAsync(rec, ctx->stream);
    if (keys) cudaFreeAsync(keys, ctx->stream);
    if (supp) cudaFreeAsync(supp, ctx->stream);
    if (counts) cudaFreeAsync(counts, ctx->stream);
    return status;
}

// =============================================================================================== parse core
static int parse_core(fcp_ctx* ctx, const uint8_t* crops, int f, int h, int w, uint8_t* labels, int32_t* hist,
                      float* logits_nchw) {
    if (f == 0) return FCP_OK;
    Model* model = &ctx->models[FCP_MODEL_BISENET];
    const int mb = std::min(ctx->par_mb, f);
    auto graph = [=](Exec& ex, const uint8_t* src, int nb, Tensor& logits) -> int {
        Tensor in3 = ex.alloc(nb, 512, 512, 3, 3);
        if (ex.ok() && !ex.dry) ex.status = launch_parse_prep(ctx, src, nb, h, w, in3.p);
        bisenet_forward(ex, in3.p, nb, logits);
        ex.free(in3);
        return ex.status;
    };
    std::vector<std::function<int(Exec&)>> plans;
    for (int nb : {mb, f % mb})
        if (nb > 0) plans.push_back([=](Exec& ex) { Tensor l; return graph(ex, nullptr, nb, l); });
    FCP_TRY(plan_reserve(ctx, model, plans));
    for (int b0 = 0; b0 < f; b0 += mb) {
        int nb = std::min(mb, f - b0);
        ctx->arena.reset();
        Exec ex{ctx, model, false};
        Tensor logits;
        {
            StageScope st(ctx, ST_PARSE_NET);
            FCP_TRY(graph(ex, crops + (size_t)b0 * h * w * 3, nb, logits));
        }
        StageScope st(ctx, ST_PARSE_TAIL);
        if (logits_nchw)
            FCP_TRY(launch_nhwc_to_nchw(ctx, logits.p, nb, 64, 64, 19, logits.cs, logits_nchw + (size_t)b0 * 19 * 64 * 64));
        if (labels || hist)
            FCP_TRY(launch_parse_tail(ctx, logits.p, 1, logits.cs, nb, 64, 64, h, w, labels ? labels + (size_t)b0 * h * w : nullptr,
                                      hist ? hist + (size_t)b0 * 19 : nullptr));
    }
    return FCP_OK;
}

// =============================================================================================== align core
static int align_core(fcp_ctx* ctx, const uint8_t* images, int n, int h, int w, const uint8_t* const* ptrs,
                      const int32_t* hs, const int32_t* ws, const int32_t* paddings, const int32_t* indices,
                      const int32_t* face_count_dev, const float* landmarks, int f, const float* target, int out_w,
                      int out_h, int border, int skew, uint8_t* crops, double* matrices, uint8_t* valid) {
    if (f == 0) return FCP_OK;
    double* inv = nullptr;
    FCP_CUDA(ctx, cudaMallocAsync(&inv, sizeof(double) * 6 * f, ctx->stream));
    StageScope st(ctx, ST_ALIGN);
    int s = launch_solve(ctx, landmarks, face_count_dev, f, target, skew, matrices, inv, valid);
    if (s == FCP_OK)
        s = launch_warp(ctx, images, n, h, w, ptrs, hs, ws, paddings, indices, face_count_dev, f, inv, valid, out_w, out_h, border, crops);
    cudaFreeAsync(inv, ctx->stream);
    return s;
}

}  // namespace fcp

// ======================================================================================================= C ABI
using namespace fcp;

extern "C" {

static int detect_outputs(fcp_ctx* ctx, const float* faces, const int32_t* face_img, const int32_t* face_count, int max_faces,
                          float* out_landmarks, int32_t* out_indices, float* out_boxes, float* out_scores,
                          int32_t* out_anchors, int32_t* out_count) {
    DevOut lms, idx, box, sc, an;
    FCP_TRY(lms.init(ctx, out_landmarks, sizeof(float) * 10 * max_faces));
    FCP_TRY(box.init(ctx, out_boxes, sizeof(float) * 4 * max_faces));
    FCP_TRY(sc.init(ctx, out_scores, sizeof(float) * max_faces));
    FCP_TRY(an.init(ctx, out_anchors, sizeof(int32_t) * max_faces));
    FCP_TRY(launch_unpack_faces(ctx, faces, face_img, face_count, max_faces, nullptr, lms.as<float>(), box.as<float>(),
                                sc.as<float>(), an.as<int32_t>()));
    int32_t count = 0;
    FCP_CUDA(ctx, cudaMemcpyAsync(&count, face_count, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int got = std::min(count, max_faces);
    FCP_TRY(lms.flush(sizeof(float) * 10 * got));
    FCP_TRY(box.flush(sizeof(float) * 4 * got));
    FCP_TRY(sc.flush(sizeof(float) * got));
    FCP_TRY(an.flush(sizeof(int32_t) * got));
    if (out_indices && got) {
        FCP_CUDA(ctx, cudaMemcpyAsync(out_indices, face_img, sizeof(int32_t) * got,
                                      is_device_ptr(out_indices) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (is_device_ptr(out_count)) FCP_CUDA(ctx, cudaMemcpyAsync(out_count, &count, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    else *out_count = count;
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (count > max_faces) return fail(ctx, FCP_ERR_CAPACITY, "max_faces too small: " + std::to_string(count) + " faces found");
    return FCP_OK;
}

int fcp_detect(fcp_ctx* ctx, const uint8_t* images, int n, int h, int w, float vis_threshold, float nms_threshold,
               int strategy, int max_faces, float* out_landmarks, int32_t* out_indices, float* out_boxes,
               float* out_scores, int32_t* out_anchors, int32_t* out_count) {
    FCP_TRY(need_model(ctx, FCP_MODEL_RETINAFACE));
    if (!images || n < 1 || h < 32 || w < 32 || max_faces < 1 || !out_count || strategy < 0 || strategy > 2)
        return fail(ctx, FCP_ERR_INVALID, "fcp_detect: bad argument");
    DevIn img;
    FCP_TRY(img.init(ctx, images, (size_t)n * h * w * 3));
    float* faces; int32_t* face_img; int32_t* face_count;
    FCP_CUDA(ctx, cudaMallocAsync(&faces, sizeof(float) * 16 * max_faces, ctx->stream));
    FCP_CUDA(ctx, cudaMallocAsync(&face_img, sizeof(int32_t) * max_faces, ctx->stream));
    FCP_CUDA(ctx, cudaMallocAsync(&face_count, sizeof(int32_t), ctx->stream));
    int s = detect_core(ctx, img.as<uint8_t>(), n, h, w, vis_threshold, nms_threshold, strategy, max_faces, faces, face_img,
                        face_count, nullptr);
    if (s == FCP_OK)
        s = detect_outputs(ctx, faces, face_img, face_count, max_faces, out_landmarks, out_indices, out_boxes, out_scores,
                           out_anchors, out_count);
    cudaFreeAsync(faces, ctx->stream); cudaFreeAsync(face_img, ctx->stream); cudaFreeAsync(face_count, ctx->stream);
    return s;
}

int fcp_detect_heads(fcp_ctx* ctx, const uint8_t* images, int n, int h, int w, float* out_heads) {
    FCP_TRY(need_model(ctx, FCP_MODEL_RETINAFACE));
    if (!images || !out_heads || n < 1 || h < 32 || w < 32) return fail(ctx, FCP_ERR_INVALID, "fcp_detect_heads: bad argument");
    DevIn img;
    FCP_TRY(img.init(ctx, images, (size_t)n * h * w * 3));
    DevOut heads;
    FCP_TRY(heads.init(ctx, out_heads, sizeof(float) * 16 * (size_t)det_num_priors(h, w) * n));
    FCP_TRY(detect_core(ctx, img.as<uint8_t>(), n, h, w, 0, 0, 0, 0, nullptr, nullptr, nullptr, heads.as<float>()));
    FCP_TRY(heads.flush());
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

int fcp_detect_post(fcp_ctx* ctx, const float* heads, int n, int h, int w, float vis_threshold, float nms_threshold,
                    int strategy, int max_faces, float* out_landmarks, int32_t* out_indices, float* out_boxes,
                    float* out_scores, int32_t* out_anchors, int32_t* out_count) {
    if (!ctx || !heads || n < 1 || max_faces < 1 || !out_count || strategy < 0 || strategy > 2)
        return fail(ctx, FCP_ERR_INVALID, "fcp_detect_post: bad argument");
    FCP_CUDA(ctx, cudaSetDevice(ctx->device));
    const int A = det_num_priors(h, w), key_cap = det_key_capacity(h, w);
    DevIn hd;
    FCP_TRY(hd.init(ctx, heads, sizeof(float) * 16 * (size_t)A * n));
    float *rec, *faces; unsigned long long* keys; unsigned char* supp; int32_t *counts, *face_img, *face_count;
    FCP_CUDA(ctx, cudaMallocAsync(&rec, (size_t)n * A * 16 * sizeof(float), ctx->stream));
    FCP_CUDA(ctx, cudaMallocAsync(&keys, (size_t)n * key_cap * sizeof(unsigned long long), ctx->stream));
    FCP_CUDA(ctx, cudaMallocAsync(&supp, (size_t)n * key_cap, ctx->stream));
    FCP_CUDA(ctx, cudaMallocAsync(&counts, sizeof(int32_t) * 2 * n, ctx->stream));
    FCP_CUDA(ctx, cudaMallocAsync(&faces, sizeof(float) * 16 * max_faces, ctx->stream));
    FCP_CUDA(ctx, cudaMallocAsync(&face_img, sizeof(int32_t) * max_faces, ctx->stream));
    FCP_CUDA(ctx, cudaMallocAsync(&face_count, sizeof(int32_t), ctx->stream));
    FCP_CUDA(ctx, cudaMemsetAsync(face_count, 0, sizeof(int32_t), ctx->stream));
    int s = launch_det_post(ctx, nullptr, hd.as<float>(), n, 0, h, w, vis_threshold, nms_threshold, strategy, max_faces, rec,
                            keys, supp, counts, counts + n, faces, face_img, face_count);
    if (s == FCP_OK)
        s = detect_outputs(ctx, faces, face_img, face_count, max_faces, out_landmarks, out_indices, out_boxes, out_scores,
                           out_anchors, out_count);
    for (void* p : {(void*)rec, (void*)keys, (void*)supp, (void*)counts, (void*)faces, (void*)face_img, (void*)face_count})
        cudaFreeAsync(p, ctx->stream);
    return s;
}

static int align_api(fcp_ctx* ctx, const uint8_t* images, int n, int h, int w, const uint8_t* const* image_ptrs,
                     const int32_t* hs, const int32_t* ws, const int32_t* paddings, const int32_t* indices,
                     const float* landmarks, int f, const float* target, int out_w, int out_h, int border_mode,
                     int allow_skew, uint8_t* out_crops, double* out_matrices, uint8_t* out_valid) {
    if (!ctx || n < 0 || f < 0 || !target || out_w < 1 || out_h < 1 || border_mode < 0 || border_mode > 4 || (f && (!indices || !landmarks || !out_crops)))
        return fail(ctx, FCP_ERR_INVALID, "fcp_align: bad argument");
    FCP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f == 0) return FCP_OK;
    DevIn img, pad, idx, lms, tgt, dptrs, dhs, dws;
    std::vector<DevIn> list_imgs(image_ptrs ? n : 0);
    std::vector<const uint8_t*> dev_ptrs;
    if (image_ptrs) {
        for (int i = 0; i < n; ++i) {
            FCP_TRY(list_imgs[i].init(ctx, image_ptrs[i], (size_t)hs[i] * ws[i] * 3));
            dev_ptrs.push_back(list_imgs[i].as<uint8_t>());
        }
        FCP_TRY(dptrs.init(ctx, dev_ptrs.data(), sizeof(void*) * n));
        FCP_TRY(dhs.init(ctx, hs, sizeof(int32_t) * n));
        FCP_TRY(dws.init(ctx, ws, sizeof(int32_t) * n));
    } else {
        FCP_TRY(img.init(ctx, images, (size_t)n * h * w * 3));
    }
    FCP_TRY(pad.init(ctx, paddings, sizeof(int32_t) * 4 * n));
    FCP_TRY(idx.init(ctx, indices, sizeof(int32_t) * f));
    FCP_TRY(lms.init(ctx, landmarks, sizeof(float) * 10 * f));
    FCP_TRY(tgt.init(ctx, target, sizeof(float) * 10));
    DevOut crops, mats, valid;
    FCP_TRY(crops.init(ctx, out_crops, (size_t)f * out_h * out_w * 3));
    FCP_TRY(mats.init(ctx, out_matrices, sizeof(double) * 6 * f, true));
    FCP_TRY(valid.init(ctx, out_valid, f, true));
    FCP_TRY(align_core(ctx, img.as<uint8_t>(), n, h, w, dptrs.as<const uint8_t*>(), dhs.as<int32_t>(), dws.as<int32_t>(),
                       pad.as<int32_t>(), idx.as<int32_t>(), nullptr, lms.as<float>(), f, tgt.as<float>(), out_w, out_h,
                       border_mode, allow_skew, crops.as<uint8_t>(), mats.as<double>(), valid.as<uint8_t>()));
    FCP_TRY(crops.flush()); FCP_TRY(mats.flush()); FCP_TRY(valid.flush());
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

int fcp_align(fcp_ctx* ctx, const uint8_t* images, int n, int h, int w, const int32_t* paddings, const int32_t* indices,
              const float* landmarks, int f, const float* target, int out_w, int out_h, int border_mode, int allow_skew,
              uint8_t* out_crops, double* out_matrices, uint8_t* out_valid) {
    if (f > 0 && !images) return fail(ctx, FCP_ERR_INVALID, "fcp_align: images is NULL");
    return align_api(ctx, images, n, h, w, nullptr, nullptr, nullptr, paddings, indices, landmarks, f, target, out_w, out_h,
                     border_mode, allow_skew, out_crops, out_matrices, out_valid);
}

int fcp_align_list(fcp_ctx* ctx, const uint8_t* const* image_ptrs, const int32_t* hs, const int32_t* ws, int n,
                   const int32_t* paddings, const int32_t* indices, const float* landmarks, int f, const float* target,
                   int out_w, int out_h, int border_mode, int allow_skew, uint8_t* out_crops, double* out_matrices,
                   uint8_t* out_valid) {
    if (f > 0 && (!image_ptrs || !hs || !ws)) return fail(ctx, FCP_ERR_INVALID, "fcp_align_list: NULL image list");
    return align_api(ctx, nullptr, n, 0, 0, image_ptrs, hs, ws, paddings, indices, landmarks, f, target, out_w, out_h,
                     border_mode, allow_skew, out_crops, out_matrices, out_valid);
}

int fcp_reduce_landmarks(fcp_ctx* ctx, const float* landmarks, int f, int k, float* out) {
    // utils.py:90-132: which source points are averaged into each of the 5 standard landmarks
    static const struct { int k; int b[10]; } table[] = {
        {5, {0, 1, 1, 2, 2, 3, 3, 4, 4, 5}},          {12, {10, 11, 11, 12, 2, 3, 3, 4, 4, 5}},
        {17, {2, 5, 7, 10, 10, 11, 13, 14, 16, 17}},  {21, {6, 9, 9, 12, 14, 15, 17, 18, 19, 20}},
        {29, {4, 9, 13, 18, 19, 20, 22, 23, 27, 28}}, {49, {19, 25, 25, 31, 13, 14, 31, 32, 37, 38}},
        {68, {36, 42, 42, 48, 30, 31, 48, 49, 54, 55}}, {98, {60, 68, 68, 76, 54, 55, 76, 77, 82, 83}},
        {106, {66, 75, 75, 84, 54, 55, 85, 86, 91, 92}}};
    if (!ctx || f < 0 || (f && (!landmarks || !out))) return fail(ctx, FCP_ERR_INVALID, "fcp_reduce_landmarks: bad argument");
    const int* bounds = nullptr;
    for (auto& t : table)
        if (t.k == k) bounds = t.b;
    if (!bounds) return fail(ctx, FCP_ERR_INVALID, "Invalid number of landmarks: " + std::to_string(k));
    FCP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f == 0) return FCP_OK;
    DevIn in;
    FCP_TRY(in.init(ctx, landmarks, sizeof(float) * 2 * k * f));
    DevOut o;
    FCP_TRY(o.init(ctx, out, sizeof(float) * 10 * f));
    FCP_TRY(launch_reduce_landmarks(ctx, in.as<float>(), f, k, bounds, o.as<float>()));
    FCP_TRY(o.flush());
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

int fcp_as_batch(fcp_ctx* ctx, const uint8_t* const* image_ptrs, const int32_t* hs, const int32_t* ws, int n, int size_w,
                 int size_h, int border_mode, uint8_t* out_batch, double* out_unscales, int32_t* out_paddings) {
    if (!ctx || n < 0 || size_w < 1 || size_h < 1 || border_mode < 0 || border_mode > 4 || (n && (!image_ptrs || !hs || !ws || !out_batch)))
        return fail(ctx, FCP_ERR_INVALID, "fcp_as_batch: bad argument");
    FCP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return FCP_OK;
    std::vector<DevIn> imgs(n);
    std::vector<const uint8_t*> dev_ptrs(n);
    for (int i = 0; i < n; ++i) {
        if (!image_ptrs[i] || hs[i] < 1 || ws[i] < 1) return fail(ctx, FCP_ERR_INVALID, "fcp_as_batch: empty image");
        FCP_TRY(imgs[i].init(ctx, image_ptrs[i], (size_t)hs[i] * ws[i] * 3));
        dev_ptrs[i] = imgs[i].as<uint8_t>();
    }
    DevOut out;
    FCP_TRY(out.init(ctx, out_batch, (size_t)n * size_h * size_w * 3));
    FCP_TRY(launch_ingest(ctx, dev_ptrs.data(), hs, ws, n, size_w, size_h, border_mode, out.as<uint8_t>(), out_unscales, out_paddings));
    FCP_TRY(out.flush());
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

int fcp_parse(fcp_ctx* ctx, const uint8_t* crops, int f, int h, int w, uint8_t* out_labels, int32_t* out_hist) {
    FCP_TRY(need_model(ctx, FCP_MODEL_BISENET));
    if (f < 0 || h < 1 || w < 1 || (f && !crops)) return fail(ctx, FCP_ERR_INVALID, "fcp_parse: bad argument");
    if (f == 0) return FCP_OK;
    DevIn c;
    FCP_TRY(c.init(ctx, crops, (size_t)f * h * w * 3));
    DevOut lab, hist;
    FCP_TRY(lab.init(ctx, out_labels, (size_t)f * h * w));
    FCP_TRY(hist.init(ctx, out_hist, sizeof(int32_t) * 19 * f));
    FCP_TRY(parse_core(ctx, c.as<uint8_t>(), f, h, w, lab.as<uint8_t>(), hist.as<int32_t>(), nullptr));
    FCP_TRY(lab.flush()); FCP_TRY(hist.flush());
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

int fcp_parse_logits(fcp_ctx* ctx, const uint8_t* crops, int f, int h, int w, float* out_logits) {
    FCP_TRY(need_model(ctx, FCP_MODEL_BISENET));
    if (f < 1 || !crops || !out_logits) return fail(ctx, FCP_ERR_INVALID, "fcp_parse_logits: bad argument");
    DevIn c;
    FCP_TRY(c.init(ctx, crops, (size_t)f * h * w * 3));
    DevOut lg;
    FCP_TRY(lg.init(ctx, out_logits, sizeof(float) * 19 * 64 * 64 * f));
    FCP_TRY(parse_core(ctx, c.as<uint8_t>(), f, h, w, nullptr, nullptr, lg.as<float>()));
    FCP_TRY(lg.flush());
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

int fcp_parse_tail(fcp_ctx* ctx, const float* logits, int f, int h, int w, uint8_t* out_labels, int32_t* out_hist) {
    if (!ctx || f < 0 || (f && !logits)) return fail(ctx, FCP_ERR_INVALID, "fcp_parse_tail: bad argument");
    FCP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f == 0) return FCP_OK;
    DevIn lg;
    FCP_TRY(lg.init(ctx, logits, sizeof(float) * 19 * 64 * 64 * f));
    DevOut lab, hist;
    FCP_TRY(lab.init(ctx, out_labels, (size_t)f * h * w));
    FCP_TRY(hist.init(ctx, out_hist, sizeof(int32_t) * 19 * f));
    FCP_TRY(launch_parse_tail(ctx, lg.as<float>(), 0, 0, f, 64, 64, h, w, lab.as<uint8_t>(), hist.as<int32_t>()));
    FCP_TRY(lab.flush()); FCP_TRY(hist.flush());
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

int fcp_masks(fcp_ctx* ctx, const uint8_t* labels, int f, int h, int w, const uint8_t* class_lut19, uint8_t* out_masks) {
    if (!ctx || f < 0 || !class_lut19 || (f && (!labels || !out_masks))) return fail(ctx, FCP_ERR_INVALID, "fcp_masks: bad argument");
    FCP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f == 0) return FCP_OK;
    size_t count = (size_t)f * h * w;
    DevIn lab, lut;
    FCP_TRY(lab.init(ctx, labels, count));
    FCP_TRY(lut.init(ctx, class_lut19, 19));
    DevOut out;
    FCP_TRY(out.init(ctx, out_masks, count));
    FCP_TRY(launch_masks(ctx, lab.as<uint8_t>(), count, lut.as<uint8_t>(), out.as<uint8_t>()));
    FCP_TRY(out.flush());
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

// RRDBNet over the gated images of a batch.  Sources: f32 NCHW `x` (value / in_div) or u8 NHWC `x_u8` (value / 255);
// sinks: the raw x4 output (f32 NCHW), the predict() result bicubic(x4, 1/4) -> clamp -> *255 -> round as f32 NCHW, or as
// u8 NHWC `out_u8[job]` (densely packed: one h*w*3 image per gated job, in image order).
// rrdb.py:124 enhances image by image to bound the reference's memory; the images are independent, so several small ones
// share a launch here (fills the 148 SMs: a 256x256 image is only 512 tiles per conv).
static int enhance_core(fcp_ctx* ctx, const float* x, float in_div, const uint8_t* x_u8, int n, int h, int w,
                        const uint8_t* gate_host, float* out_x4_nchw, float* out_x1_nchw, uint8_t* out_u8) {
    Model* model = &ctx->models[FCP_MODEL_RRDBNET];
    std::vector<int> jobs;
    for (int i = 0; i < n; ++i)
        if (!gate_host || gate_host[i]) jobs.push_back(i);
    if (jobs.empty()) return FCP_OK;
    const long long px = (long long)h * w;
    int nb = (int)std::max(1LL, std::min<long long>(8, (1LL << 20) / std::max(1LL, px)));
    if (const char* e = getenv("FCP_ENH_BATCH")) nb = std::max(1, atoi(e));
    nb = std::min<int>(nb, (int)jobs.size());
    const ConvWeights* cf = nullptr;
    {
        auto it = model->conv.find("conv_first");
        if (it == model->conv.end()) return fail(ctx, FCP_ERR_STATE, "conv not finalized: conv_first");
        cf = &it->second;
    }
    std::vector<std::function<int(Exec&)>> plans;
    auto noop = [](Tensor) { return FCP_OK; };
    for (int b : {nb, (int)(jobs.size() % nb)})
        if (b > 0) plans.push_back([=](Exec& ex) { Tensor t; return rrdbnet_forward(ex, noop, b, h, w, t); });
    FCP_TRY(plan_reserve(ctx, model, plans));
    for (size_t j0 = 0; j0 < jobs.size(); j0 += nb) {
        const int b = (int)std::min<size_t>(nb, jobs.size() - j0);
        ctx->arena.reset();
        Exec ex{ctx, model, false};
        Tensor x4;
        auto fill = [&](Tensor first) -> int {
            for (int k = 0; k < b; ++k) {
                const int i = jobs[j0 + k];
                Tensor f1 = first;
                f1.n = 1;
                f1.p = first.p + (size_t)k * px * first.cs;
                if (x_u8) FCP_TRY(launch_conv3_first_u8(ctx, x_u8 + (size_t)i * px * 3, 1, h, w, cf->w_kn, cf->shift, f1));
                else FCP_TRY(launch_conv3_first(ctx, x + (size_t)i * 3 * px, in_div, 1, h, w, cf->w_kn, cf->shift, f1));
            }
            return FCP_OK;
        };
        FCP_TRY(rrdbnet_forward(ex, fill, b, h, w, x4));
        for (int k = 0; k < b; ++k) {
            const int i = jobs[j0 + k];
            Tensor x1 = x4;
            x1.n = 1;
            x1.p = x4.p + (size_t)k * 16 * px * x4.cs;
            if (out_x4_nchw) FCP_TRY(launch_nhwc_to_nchw(ctx, x1.p, 1, 4 * h, 4 * w, 3, x1.cs, out_x4_nchw + (size_t)i * 3 * 16 * px));
            if (out_x1_nchw) FCP_TRY(launch_rrdb_tail(ctx, x1, out_x1_nchw + (size_t)i * 3 * px, h, w));
            if (out_u8) FCP_TRY(launch_rrdb_tail_u8(ctx, x1, out_u8 + (j0 + k) * px * 3, h, w));
        }
    }
    return FCP_OK;
}

int fcp_group(fcp_ctx* ctx, const uint8_t* labels, const int32_t* hist, int f, int h, int w, const int32_t* attr_codes,
              const int32_t* attr_offsets, int n_attr, int attr_threshold, int join_and, const uint8_t* mask_lut, int n_mask,
              int mask_threshold, uint8_t* out_attr, uint8_t* out_mask, uint8_t* out_masks) {
    if (!ctx || f < 0 || n_attr < 0 || n_mask < 0 || n_mask > 32 || (f && !hist) || (n_attr && (!attr_codes || !attr_offsets || !out_attr)) ||
        (n_mask && (!mask_lut || !out_mask)) || (out_masks && (!labels || h < 1 || w < 1)))
        return fail(ctx, FCP_ERR_INVALID, "fcp_group: bad argument");
    FCP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (f == 0 || n_attr + n_mask == 0) return FCP_OK;
    if (n_attr && is_device_ptr(attr_offsets)) return fail(ctx, FCP_ERR_INVALID, "fcp_group: attr_offsets must be host memory");
    const int n_codes = n_attr ? attr_offsets[n_attr] : 0;
    DevIn hs, codes, offs, lut, lab;
    FCP_TRY(hs.init(ctx, hist, sizeof(int32_t) * 19 * f));
    FCP_TRY(codes.init(ctx, attr_codes, sizeof(int32_t) * (n_codes > 0 ? n_codes : 1)));
    FCP_TRY(offs.init(ctx, attr_offsets, sizeof(int32_t) * (n_attr + 1)));
    FCP_TRY(lut.init(ctx, mask_lut, (size_t)19 * (n_mask > 0 ? n_mask : 1)));
    DevOut oa, om, omasks;
    FCP_TRY(oa.init(ctx, out_attr, (size_t)n_attr * f));
    FCP_TRY(om.init(ctx, out_mask, (size_t)n_mask * f));
    FCP_TRY(launch_group(ctx, hs.as<int32_t>(), f, codes.as<int32_t>(), offs.as<int32_t>(), n_attr, attr_threshold, join_and,
                         lut.as<uint8_t>(), n_mask, mask_threshold, oa.as<uint8_t>(), om.as<uint8_t>()));
    if (out_masks && n_mask) {
        const size_t count = (size_t)f * h * w;
        FCP_TRY(lab.init(ctx, labels, count));
        FCP_TRY(omasks.init(ctx, out_masks, count * n_mask));
        FCP_TRY(launch_multi_masks(ctx, lab.as<uint8_t>(), count, lut.as<uint8_t>(), n_mask, omasks.as<uint8_t>()));
        FCP_TRY(omasks.flush());
    }
    FCP_TRY(oa.flush()); FCP_TRY(om.flush());
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

int fcp_enhance(fcp_ctx* ctx, float* images, int n, int h, int w, const uint8_t* do_enhance) {
    FCP_TRY(need_model(ctx, FCP_MODEL_RRDBNET));
    if (!images || n < 1 || h < 1 || w < 1) return fail(ctx, FCP_ERR_INVALID, "fcp_enhance: bad argument");
    size_t bytes = sizeof(float) * 3 * h * w * (size_t)n;
    std::vector<uint8_t> gate(n, 1);
    if (do_enhance) {
        if (is_device_ptr(do_enhance)) FCP_CUDA(ctx, cudaMemcpy(gate.data(), do_enhance, n, cudaMemcpyDeviceToHost));
        else gate.assign(do_enhance, do_enhance + n);
    }
    if (is_device_ptr(images)) {
        // in-place on device: the tail writes image i only after the whole graph of image i has consumed it
        FCP_TRY(enhance_core(ctx, images, 255.f, nullptr, n, h, w, gate.data(), nullptr, images, nullptr));
    } else {
        float* dev;
        FCP_CUDA(ctx, cudaMallocAsync(&dev, bytes, ctx->stream));
        FCP_CUDA(ctx, cudaMemcpyAsync(dev, images, bytes, cudaMemcpyHostToDevice, ctx->stream));
        int s = enhance_core(ctx, dev, 255.f, nullptr, n, h, w, gate.data(), nullptr, dev, nullptr);
        if (s == FCP_OK && cudaMemcpyAsync(images, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
            s = fail(ctx, FCP_ERR_CUDA, "D2H copy failed");
        cudaFreeAsync(dev, ctx->stream);
        FCP_TRY(s);
    }
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

int fcp_enhance_u8(fcp_ctx* ctx, uint8_t* images, int n, int h, int w, const uint8_t* do_enhance) {
    FCP_TRY(need_model(ctx, FCP_MODEL_RRDBNET));
    if (!images || n < 1 || h < 1 || w < 1) return fail(ctx, FCP_ERR_INVALID, "fcp_enhance_u8: bad argument");
    const size_t img_bytes = (size_t)h * w * 3;
    std::vector<uint8_t> gate(n, 1);
    if (do_enhance) {
        if (is_device_ptr(do_enhance)) FCP_CUDA(ctx, cudaMemcpy(gate.data(), do_enhance, n, cudaMemcpyDeviceToHost));
        else gate.assign(do_enhance, do_enhance + n);
    }
    int jobs = 0;
    for (uint8_t g : gate) jobs += g != 0;
    if (jobs == 0) return FCP_OK;
    DevIn in;
    FCP_TRY(in.init(ctx, images, img_bytes * n));
    uint8_t* dense = nullptr;
    FCP_CUDA(ctx, cudaMallocAsync(&dense, img_bytes * jobs, ctx->stream));
    int s = enhance_core(ctx, nullptr, 255.f, in.as<uint8_t>(), n, h, w, gate.data(), nullptr, nullptr, dense);
    const bool dev = is_device_ptr(images);
    for (int i = 0, j = 0; i < n && s == FCP_OK; ++i)
        if (gate[i] && cudaMemcpyAsync(images + (size_t)i * img_bytes, dense + (size_t)(j++) * img_bytes, img_bytes,
                                       dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
            s = fail(ctx, FCP_ERR_CUDA, "fcp_enhance_u8: result copy failed");
    cudaFreeAsync(dense, ctx->stream);
    FCP_TRY(s);
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

int fcp_enhance_gate(fcp_ctx* ctx, const float* landmarks, const int32_t* indices, int f, int n, int h, int w,
                     float min_face_factor, uint8_t* out_gate) {
    if (!ctx || f < 0 || n < 0 || !out_gate || (f && (!landmarks || !indices))) return fail(ctx, FCP_ERR_INVALID, "fcp_enhance_gate: bad argument");
    FCP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return FCP_OK;
    DevIn lms, idx, cnt;
    FCP_TRY(lms.init(ctx, landmarks, sizeof(float) * 10 * f));
    FCP_TRY(idx.init(ctx, indices, sizeof(int32_t) * f));
    const int32_t c = f;
    FCP_TRY(cnt.init(ctx, &c, sizeof c));
    DevOut g;
    FCP_TRY(g.init(ctx, out_gate, n));
    FCP_TRY(launch_enhance_gate(ctx, lms.as<float>(), idx.as<int32_t>(), cnt.as<int32_t>(), f, n, h, w, min_face_factor, g.as<uint8_t>()));
    FCP_TRY(g.flush());
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

int fcp_enhance_forward(fcp_ctx* ctx, const float* x, int n, int h, int w, float* out) {
    FCP_TRY(need_model(ctx, FCP_MODEL_RRDBNET));
    if (!x || !out || n < 1) return fail(ctx, FCP_ERR_INVALID, "fcp_enhance_forward: bad argument");
    DevIn in;
    FCP_TRY(in.init(ctx, x, sizeof(float) * 3 * h * w * (size_t)n));
    DevOut o;
    FCP_TRY(o.init(ctx, out, sizeof(float) * 3 * 16 * h * w * (size_t)n));
    FCP_TRY(enhance_core(ctx, in.as<float>(), 1.f, nullptr, n, h, w, nullptr, o.as<float>(), nullptr, nullptr));
    FCP_TRY(o.flush());
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

int fcp_pipeline(fcp_ctx* ctx, const uint8_t* images, int n, int h, int w, const int32_t* paddings, float vis_threshold,
                 float nms_threshold, int strategy, const float* target, int out_w, int out_h, int border_mode,
                 int allow_skew, int max_faces, float* out_landmarks, int32_t* out_indices, int32_t* out_count,
                 uint8_t* out_crops, double* out_matrices, uint8_t* out_valid, uint8_t* out_labels, int32_t* out_hist) {
    FCP_TRY(need_model(ctx, FCP_MODEL_RETINAFACE));
    const bool do_parse = out_labels || out_hist;
    const bool do_enhance = ctx->enh_enabled;
    if (do_parse) FCP_TRY(need_model(ctx, FCP_MODEL_BISENET));
    if (do_enhance) FCP_TRY(need_model(ctx, FCP_MODEL_RRDBNET));
    if (!images || n < 1 || max_faces < 1 || !out_count || !target || !out_crops || strategy < 0 || strategy > 2 ||
        border_mode < 0 || border_mode > 4)
        return fail(ctx, FCP_ERR_INVALID, "fcp_pipeline: bad argument");
    DevIn pad, tgt;
    FCP_TRY(pad.init(ctx, paddings, sizeof(int32_t) * 4 * n));
    FCP_TRY(tgt.init(ctx, target, sizeof(float) * 10));
    const size_t img_bytes = (size_t)h * w * 3;
    // every stream-ordered temporary of the call; released (and the micro-batch hook cleared) on EVERY exit path
    struct Temps {
        fcp_ctx* ctx;
        std::vector<void*> ptrs;
        bool staged = false;
        ~Temps() {
            ctx->on_microbatch = nullptr;
            ctx->mb_sched.clear();
            if (staged && ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);   // no copy may outlive its buffer
            for (void* p : ptrs) cudaFreeAsync(p, ctx->stream);
        }
        int alloc(void** p, size_t bytes) {
            if (cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream) != cudaSuccess) {
                cudaGetLastError();
                return fail(ctx, FCP_ERR_CUDA, "fcp_pipeline: out of device memory (" + std::to_string(bytes >> 20) + " MiB)");
            }
            ptrs.push_back(*p);
            return FCP_OK;
        }
    } tmp{ctx};
    float* faces; int32_t* face_img; int32_t* face_count; uint8_t* gate = nullptr;
    FCP_TRY(tmp.alloc((void**)&faces, sizeof(float) * 16 * max_faces));
    FCP_TRY(tmp.alloc((void**)&face_img, sizeof(int32_t) * max_faces));
    FCP_TRY(tmp.alloc((void**)&face_count, sizeof(int32_t)));
    if (do_enhance) FCP_TRY(tmp.alloc((void**)&gate, n));
    // Images in host memory are copied per detector micro-batch on a second stream, one micro-batch ahead of the compute
    // stream, so that (with pinned buffers) only the first micro-batch's copy is exposed.
    const uint8_t* dimg = images;
    if (!is_device_ptr(images)) {
        uint8_t* staged = nullptr;
        // schedule: a first micro-batch of at most 16 images (its copy is the only one the compute stream has to wait for
        // from cold), then det_mb images each; micro-batch c+1 is copied while micro-batch c computes
        const int mb = std::min(ctx->det_mb, n);
        std::vector<int> sched, start;
        for (int b0 = 0; b0 < n;) {
            const int nb = std::min(b0 == 0 && n > 16 ? std::min(mb, 16) : mb, n - b0);
            start.push_back(b0); sched.push_back(nb); b0 += nb;
        }
        const int chunks = (int)sched.size();
        ctx->mb_sched = sched;
        FCP_TRY(tmp.alloc((void**)&staged, img_bytes * n));
        if (!ctx->copy_stream) FCP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        while ((int)ctx->copy_events.size() < chunks + 1) {
            cudaEvent_t e;
            FCP_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->copy_events.push_back(e);
        }
        FCP_CUDA(ctx, cudaEventRecord(ctx->copy_events[chunks], ctx->stream));              // the allocation is stream-ordered
        FCP_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->copy_events[chunks], 0));
        tmp.staged = true;
        auto issue = [ctx, images, staged, img_bytes, sched, start, chunks](int c) -> int {
            if (c >= chunks) return FCP_OK;
            const size_t off = (size_t)start[c] * img_bytes, bytes = (size_t)sched[c] * img_bytes;
            FCP_CUDA(ctx, cudaMemcpyAsync(staged + off, images + off, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
            FCP_CUDA(ctx, cudaEventRecord(ctx->copy_events[c], ctx->copy_stream));
            return FCP_OK;
        };
        FCP_TRY(issue(0));
        ctx->on_microbatch = [ctx, issue](int c) -> int {
            FCP_TRY(issue(c + 1));
            FCP_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_events[c], 0));
            return FCP_OK;
        };
        dimg = staged;
    }
    FCP_TRY(detect_core(ctx, dimg, n, h, w, vis_threshold, nms_threshold, strategy, max_faces, faces, face_img, face_count, nullptr));
    ctx->on_microbatch = nullptr;
    ctx->mb_sched.clear();
    // landmark un-pad (cropper.py:822) happens while unpacking the face records; the enhancement gate (rrdb.py:124-141)
    // is evaluated on the device from the un-padded landmarks and comes back with the face count in the one host sync
    DevOut lms, crops, mats, valid, lab, hist;
    int32_t count = 0;
    std::vector<uint8_t> gate_host(do_enhance ? n : 0, 0);
    FCP_TRY(lms.init(ctx, out_landmarks, sizeof(float) * 10 * max_faces, true));
    FCP_TRY(launch_unpack_faces(ctx, faces, face_img, face_count, max_faces, pad.as<int32_t>(), lms.as<float>(), nullptr, nullptr, nullptr));
    if (do_enhance) {
        FCP_TRY(launch_enhance_gate(ctx, lms.as<float>(), face_img, face_count, max_faces, n, h, w, ctx->enh_threshold, gate));
        FCP_CUDA(ctx, cudaMemcpyAsync(gate_host.data(), gate, n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    FCP_CUDA(ctx, cudaMemcpyAsync(&count, face_count, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int f = std::min(count, max_faces);
    if (f > 0) {
        // ---- enhance (rrdb.py:83-146, cropper.py:833-836): the gated images are rebuilt as u8 on the device; the warp
        //      reads them through a per-image pointer table, the caller's batch is never modified
        const uint8_t* const* ptr_table = nullptr; const int32_t* hs_dev = nullptr; const int32_t* ws_dev = nullptr;
        int n_gated = 0;
        for (uint8_t g : gate_host) n_gated += g;
        if (n_gated > 0) {
            StageScope st(ctx, ST_ENHANCE);
            uint8_t* enhanced = nullptr; void* table = nullptr; int32_t* dims = nullptr;
            FCP_TRY(tmp.alloc((void**)&enhanced, img_bytes * n_gated));
            FCP_TRY(tmp.alloc(&table, sizeof(void*) * n));
            FCP_TRY(tmp.alloc((void**)&dims, sizeof(int32_t) * 2 * n));
            FCP_TRY(enhance_core(ctx, nullptr, 255.f, dimg, n, h, w, gate_host.data(), nullptr, nullptr, enhanced));
            std::vector<const uint8_t*> ptrs(n);
            std::vector<int32_t> hw(2 * n);
            for (int i = 0, j = 0; i < n; ++i) {
                ptrs[i] = gate_host[i] ? enhanced + (size_t)(j++) * img_bytes : dimg + (size_t)i * img_bytes;
                hw[i] = h; hw[n + i] = w;
            }
            FCP_CUDA(ctx, cudaMemcpyAsync(table, ptrs.data(), sizeof(void*) * n, cudaMemcpyHostToDevice, ctx->stream));
            FCP_CUDA(ctx, cudaMemcpyAsync(dims, hw.data(), sizeof(int32_t) * 2 * n, cudaMemcpyHostToDevice, ctx->stream));
            FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));                            // the host vectors die with this scope
            ptr_table = static_cast<const uint8_t* const*>(table); hs_dev = dims; ws_dev = dims + n;
        }
        FCP_TRY(crops.init(ctx, out_crops, (size_t)f * out_h * out_w * 3));
        FCP_TRY(mats.init(ctx, out_matrices, sizeof(double) * 6 * f, true));
        FCP_TRY(valid.init(ctx, out_valid, f, true));
        FCP_TRY(align_core(ctx, ptr_table ? nullptr : dimg, n, h, w, ptr_table, hs_dev, ws_dev, pad.as<int32_t>(), face_img, nullptr,
                           lms.as<float>(), f, tgt.as<float>(), out_w, out_h, border_mode, allow_skew, crops.as<uint8_t>(),
                           mats.as<double>(), valid.as<uint8_t>()));
        // ---- the one collective: this rank's face records, all-gathered on the side stream while the parser runs
        if (ctx->gather_out)
            FCP_TRY(gather_meta_async(ctx, lms.as<float>(), face_img, mats.as<double>(), valid.as<uint8_t>(), face_count,
                                      ctx->gather_cap, ctx->gather_base, ctx->gather_out));
        if (do_parse) {
            FCP_TRY(lab.init(ctx, out_labels, (size_t)f * out_h * out_w));
            FCP_TRY(hist.init(ctx, out_hist, sizeof(int32_t) * 19 * f));
            FCP_TRY(parse_core(ctx, crops.as<uint8_t>(), f, out_h, out_w, lab.as<uint8_t>(), hist.as<int32_t>(), nullptr));
        }
        if (ctx->gather_out) FCP_TRY(gather_meta_join(ctx));
        FCP_TRY(lms.flush(sizeof(float) * 10 * f));
        FCP_TRY(crops.flush()); FCP_TRY(mats.flush()); FCP_TRY(valid.flush()); FCP_TRY(lab.flush()); FCP_TRY(hist.flush());
        if (out_indices)
            FCP_CUDA(ctx, cudaMemcpyAsync(out_indices, face_img, sizeof(int32_t) * f,
                                          is_device_ptr(out_indices) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    } else if (ctx->gather_out) {
        FCP_TRY(gather_meta_async(ctx, lms.as<float>(), face_img, nullptr, nullptr, face_count, ctx->gather_cap, ctx->gather_base,
                                  ctx->gather_out));
        FCP_TRY(gather_meta_join(ctx));
    }
    if (is_device_ptr(out_count)) FCP_CUDA(ctx, cudaMemcpyAsync(out_count, &count, sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    else *out_count = count;
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (count > max_faces) return fail(ctx, FCP_ERR_CAPACITY, "max_faces too small: " + std::to_string(count) + " faces found");
    return FCP_OK;
}

int fcp_conv2d(fcp_ctx* ctx, const float* x, int n, int h, int w, int cin, const float* weight, int cout, int k, int stride,
               int pad, const float* scale, const float* shift, const float* residual, int act, float slope, int impl,
               float* out) {
    if (!ctx || !x || !weight || !out || n < 1 || cin < 1 || cout < 1) return fail(ctx, FCP_ERR_INVALID, "fcp_conv2d: bad argument");
    FCP_CUDA(ctx, cudaSetDevice(ctx->device));
    Model tmp;
    HostTensor wt;
    wt.shape = {cout, cin, k, k};
    wt.data.assign(weight, weight + (size_t)cout * cin * k * k);
    tmp.host["c.weight"] = wt;
    size_t mark = ctx->device_allocs.size();
    FCP_TRY(pack_conv(ctx, tmp, {"c"}, "", "c", scale, shift));
    ConvWeights& cw = tmp.conv["c"];
    const int ho = odim(h, k, stride, pad), wo = odim(w, k, stride, pad);
    const int cs_out = (cout + 3) / 4 * 4;
    DevIn xin, res;
    FCP_TRY(xin.init(ctx, x, sizeof(float) * (size_t)n * h * w * cin));
    FCP_TRY(res.init(ctx, residual, sizeof(float) * (size_t)n * ho * wo * cout));
    float* dout;
    FCP_CUDA(ctx, cudaMallocAsync(&dout, sizeof(float) * (size_t)n * ho * wo * cs_out, ctx->stream));
    ConvOp op;
    op.in = Tensor{const_cast<float*>(xin.as<float>()), n, h, w, cin, cin, 0};
    op.out = Tensor{dout, n, ho, wo, cout, cs_out, 0};
    op.wt = &cw; op.stride = stride; op.pad = pad; op.act = act; op.slope = slope; op.impl = impl;
    if (residual) { op.res1 = res.as<float>(); op.res1_cs = cout; op.res1_co = 0; }
    int s = run_conv(ctx, op);
    if (s == FCP_OK) {
        cudaError_t e = cudaMemcpy2DAsync(out, sizeof(float) * cout, dout, sizeof(float) * cs_out, sizeof(float) * cout,
                                          (size_t)n * ho * wo, is_device_ptr(out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                                          ctx->stream);
        if (e != cudaSuccess) s = fail(ctx, FCP_ERR_CUDA, std::string("conv2d output copy: ") + cudaGetErrorString(e));
    }
    cudaFreeAsync(dout, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    for (size_t i = mark; i < ctx->device_allocs.size(); ++i) cudaFree(ctx->device_allocs[i]);
    ctx->device_allocs.resize(mark);
    return s;
}

}  // extern "C"
