// runtime.cu — context, stream-ordered arena, weight ingestion (BN folding + packing) and small host helpers.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include <cuda_fp16.h>

#include "common.h"
#include "graphs.h"

namespace fcp {

int fail(fcp_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->error = msg;
    return code;
}

// ------------------------------------------------------------------------------------------------ stage timers
StageScope::StageScope(fcp_ctx* c, int stage, cudaStream_t st, bool use_stream) : ctx(c), stream(use_stream ? st : c->stream) {
    if (!ctx->profile) return;
    if (ctx->stage_used == ctx->stage_recs.size()) {
        fcp_ctx::StageRec r{};
        if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
        ctx->stage_recs.push_back(r);
    }
    fcp_ctx::StageRec& r = ctx->stage_recs[ctx->stage_used++];
    r.stage = stage;
    cudaEventRecord(r.a, stream);
    end = r.b;
}
StageScope::~StageScope() {
    if (end) cudaEventRecord(end, stream);
}

// ------------------------------------------------------------------------------------------------------ Arena
Arena::~Arena() {
    if (base_) cudaFree(base_);
}

bool Arena::reserve(size_t bytes) {
    if (bytes <= cap_ && base_) return true;
    if (base_) {
        cudaDeviceSynchronize();
        cudaFree(base_);
        base_ = nullptr;
    }
    if (cudaMalloc(&base_, bytes) != cudaSuccess) {
        cap_ = 0;
        base_ = nullptr;
        return false;
    }
    cap_ = bytes;
    reset();
    return true;
}

void Arena::reset() {
    blocks_.clear();
    blocks_.push_back({0, plan_ ? (size_t)1 << 60 : cap_, false});
}

void Arena::set_plan_mode(bool on) {
    plan_ = on;
    high_ = 0;
    reset();
}

void* Arena::alloc(size_t bytes) {
    bytes = (bytes + 1023) & ~(size_t)1023;
    for (size_t i = 0; i < blocks_.size(); ++i) {
        Block& b = blocks_[i];
        if (b.used || b.size < bytes) continue;
        size_t off = b.off;
        if (b.size > bytes) {
            Block rest{b.off + bytes, b.size - bytes, false};
            b.size = bytes;
            b.used = true;
            blocks_.insert(blocks_.begin() + i + 1, rest);
        } else {
            b.used = true;
        }
        high_ = std::max(high_, off + bytes);
        // plan mode hands out fake (never dereferenced) addresses above a non-null base
        return (plan_ ? reinterpret_cast<char*>((uintptr_t)1 << 40) : base_) + off;
    }
    return nullptr;
}

void Arena::free(void* p) {
    if (!p) return;
    size_t off = static_cast<char*>(p) - (plan_ ? reinterpret_cast<char*>((uintptr_t)1 << 40) : base_);
    for (size_t i = 0; i < blocks_.size(); ++i) {
        if (blocks_[i].off != off) continue;
        blocks_[i].used = false;
        if (i + 1 < blocks_.size() && !blocks_[i + 1].used) {
            blocks_[i].size += blocks_[i + 1].size;
            blocks_.erase(blocks_.begin() + i + 1);
        }
        if (i > 0 && !blocks_[i - 1].used) {
            blocks_[i - 1].size += blocks_[i].size;
            blocks_.erase(blocks_.begin() + i);
        }
        return;
    }
}

// --------------------------------------------------------------------------------------------- weight helpers
static const HostTensor* find(const Model& m, const std::string& key) {
    auto it = m.host.find(key);
    return it == m.host.end() ? nullptr : &it->second;
}

static int upload(fcp_ctx* ctx, const std::vector<float>& v, float** out) {
    FCP_CUDA(ctx, cudaMalloc(out, v.size() * sizeof(float)));
    ctx->device_allocs.push_back(*out);
    FCP_CUDA(ctx, cudaMemcpy(*out, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return FCP_OK;
}

static inline float tf32_trunc(float x) {
    uint32_t u;
    std::memcpy(&u, &x, 4);
    u &= 0xFFFFE000u;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}

// f16x3 packing of a K-major fp32 matrix wk[cout_pad][taps][cin] (BN scale already folded in): one power-of-two scale per
// layer puts max |w| into [2^13, 2^14); hi = rn_f16(w * 2^e), lo = rn_f16(w * 2^e - hi); each tap is padded to a multiple
// of 64 channels (one K-block of the kernel) with zeros.
static int pack_f16(fcp_ctx* ctx, ConvWeights& cw, const std::vector<float>& wk, int taps, int cin) {
    float wmax = 0.f;
    for (float v : wk) wmax = std::max(wmax, std::fabs(v));
    int ew = 0;
    if (wmax > 0.f && std::isfinite(wmax)) std::frexp(wmax, &ew);        // wmax = m * 2^ew, 0.5 <= m < 1
    cw.w_exp = std::min(40, std::max(-40, 14 - ew));                     // wmax * 2^w_exp in [2^13, 2^14)
    cw.cin_p = (cin + 31) / 32 * 32;                                     // K order: (tap, channel), 32-channel units (conv_tc.cu)
    const size_t Kp = ((size_t)taps * cw.cin_p + 63) / 64 * 64;          // whole K-blocks: a trailing half block reads zeros
    std::vector<__half> hi((size_t)cw.cout_pad * Kp, __float2half_rn(0.f)), lo(hi);
    const float sc = std::ldexp(1.0f, cw.w_exp);
    for (int o = 0; o < cw.cout_pad; ++o)
        for (int t = 0; t < taps; ++t)
            for (int c = 0; c < cin; ++c) {
                const float v = wk[((size_t)o * taps + t) * cin + c] * sc;
                const __half h = __float2half_rn(v);
                hi[(size_t)o * Kp + (size_t)t * cw.cin_p + c] = h;
                lo[(size_t)o * Kp + (size_t)t * cw.cin_p + c] = __float2half_rn(v - __half2float(h));
            }
    for (auto* pp : {&cw.h_hi, &cw.h_lo}) {
        FCP_CUDA(ctx, cudaMalloc(pp, hi.size() * sizeof(__half)));
        ctx->device_allocs.push_back(*pp);
    }
    FCP_CUDA(ctx, cudaMemcpy(cw.h_hi, hi.data(), hi.size() * sizeof(__half), cudaMemcpyHostToDevice));
    FCP_CUDA(ctx, cudaMemcpy(cw.h_lo, lo.data(), lo.size() * sizeof(__half), cudaMemcpyHostToDevice));
    return FCP_OK;
}

// Packs conv `convs[i].weight` (OIHW, concatenated along Cout) with optional bias and optional BatchNorm `bn`
// (running stats folded like ATen's eval batch_norm: alpha = gamma/sqrt(var+eps), beta = bias - mean*alpha).
int pack_conv(fcp_ctx* ctx, Model& m, const std::vector<std::string>& convs, const std::string& bn,
              const std::string& name, const float* explicit_scale, const float* explicit_shift) {
    ConvWeights cw;
    std::vector<const HostTensor*> ws, bs;
    for (auto& c : convs) {
        const HostTensor* w = find(m, c + ".weight");
        if (!w || w->shape.size() != 4) return fail(ctx, FCP_ERR_STATE, "missing conv weight: " + c + ".weight");
        ws.push_back(w);
        bs.push_back(find(m, c + ".bias"));
        if (cw.cin == 0) { cw.cin = (int)w->shape[1]; cw.k = (int)w->shape[2]; }
        if (w->shape[1] != cw.cin || w->shape[2] != cw.k || w->shape[3] != cw.k)
            return fail(ctx, FCP_ERR_INVALID, "conv shapes of a fused group differ: " + c);
        cw.cout += (int)w->shape[0];
    }
    cw.cout_pad = (cw.cout + 31) / 32 * 32;
    const int K = cw.k * cw.k * cw.cin;
    // ---- per-channel affine after the conv: conv bias, then BatchNorm (or the explicit scale/shift of the test hook)
    const int vec_pad = (cw.cout_pad + 127) / 128 * 128;                 // the tcgen05 epilogue reads whole tile halves
    std::vector<float> scale(vec_pad, 1.f), shift(vec_pad, 0.f);
    {
        int o0 = 0;
        for (size_t g = 0; g < ws.size(); ++g) {
            const int co_n = (int)ws[g]->shape[0];
            if (bs[g])
                for (int o = 0; o < co_n; ++o) shift[o0 + o] = bs[g]->data[o];
            o0 += co_n;
        }
    }
    if (!bn.empty()) {
        const HostTensor *g = find(m, bn + ".weight"), *b = find(m, bn + ".bias"), *mu = find(m, bn + ".running_mean"),
                         *var = find(m, bn + ".running_var");
        if (!g || !b || !mu || !var || (int)g->data.size() != cw.cout)
            return fail(ctx, FCP_ERR_STATE, "missing/mismatched BatchNorm tensors: " + bn);
        for (int o = 0; o < cw.cout; ++o) {
            float invstd = 1.0f / std::sqrt(var->data[o] + 1e-5f);
            float alpha = g->data[o] * invstd;
            float beta = b->data[o] - mu->data[o] * alpha;
            shift[o] = shift[o] * alpha + beta;   // conv bias (0 here) goes through the BN as well
            scale[o] = alpha;
        }
    }
    for (int o = 0; o < cw.cout; ++o) {
        if (explicit_scale) scale[o] = explicit_scale[o];
        if (explicit_shift) shift[o] = explicit_shift[o];
    }
    // ---- weights.  CUDA-core kernel: [K][cout_pad], plain (it applies scale/shift in its epilogue).  Tensor-core kernel:
    // K-major [cout_pad][K] with the per-channel SCALE FOLDED IN (w*scale, then split into tf32 hi + lo), so that its
    // epilogue only adds the shift - done while the K loop still runs, off the tile-boundary critical path.
    std::vector<float> wkn((size_t)K * cw.cout_pad, 0.f);
    std::vector<float> whi((size_t)cw.cout_pad * K, 0.f), wlo((size_t)cw.cout_pad * K, 0.f), wfull((size_t)cw.cout_pad * K, 0.f);
    int o0 = 0;
    for (size_t g = 0; g < ws.size(); ++g) {
        const HostTensor& w = *ws[g];
        int co_n = (int)w.shape[0];
        for (int o = 0; o < co_n; ++o)
            for (int c = 0; c < cw.cin; ++c)
                for (int r = 0; r < cw.k; ++r)
                    for (int s = 0; s < cw.k; ++s) {
                        float v = w.data[(((size_t)o * cw.cin + c) * cw.k + r) * cw.k + s];
                        size_t kk = (size_t)(r * cw.k + s) * cw.cin + c;
                        wkn[kk * cw.cout_pad + o0 + o] = v;
                        const float vs = v * scale[o0 + o];
                        float hi = tf32_trunc(vs);
                        wfull[(size_t)(o0 + o) * K + kk] = vs;
                        whi[(size_t)(o0 + o) * K + kk] = hi;
                        wlo[(size_t)(o0 + o) * K + kk] = tf32_trunc(vs - hi);
                    }
        o0 += co_n;
    }
    FCP_TRY(upload(ctx, wkn, &cw.w_kn));
    FCP_TRY(upload(ctx, whi, &cw.w_hi));
    FCP_TRY(upload(ctx, wlo, &cw.w_lo));
    if (cw.cin % 32 == 0) FCP_TRY(pack_f16(ctx, cw, wfull, cw.k * cw.k, cw.cin));
    FCP_TRY(upload(ctx, scale, &cw.scale));
    FCP_TRY(upload(ctx, shift, &cw.shift));
    m.conv[name] = cw;
    m.vec[name + ".scale"] = scale;
    return FCP_OK;
}

int pack_stem_rows(fcp_ctx* ctx, Model& m, const std::string& conv, const std::string& name) {
    const HostTensor* w = find(m, conv + ".weight");
    auto it = m.conv.find(conv);
    if (!w || w->shape.size() != 4 || w->shape[1] != 3 || w->shape[2] != 7 || w->shape[3] != 7 || it == m.conv.end())
        return fail(ctx, FCP_ERR_STATE, "stem weights missing or not 7x7x3: " + conv);
    ConvWeights cw = it->second;                       // shares scale / shift (folded BN) with the square packing
    const std::vector<float>& scale = m.vec[conv + ".scale"];
    if ((int)scale.size() < cw.cout) return fail(ctx, FCP_ERR_STATE, "stem scale missing: " + conv);
    cw.cin = 32; cw.k = 7; cw.kh = 7; cw.kw = 1; cw.alg_k = 147;
    cw.w_kn = nullptr;                                 // no CUDA-core packing: this route exists on the tensor cores only
    const int K = 7 * 32;
    std::vector<float> whi((size_t)cw.cout_pad * K, 0.f), wlo((size_t)cw.cout_pad * K, 0.f), wfull((size_t)cw.cout_pad * K, 0.f);
    for (int o = 0; o < cw.cout; ++o)
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 7; ++r)
                for (int sx = 0; sx < 7; ++sx) {
                    const float v = w->data[(((size_t)o * 3 + c) * 7 + r) * 7 + sx] * scale[o];   // folded BN scale, like pack_conv
                    const size_t kk = (size_t)r * 32 + sx * 3 + c;
                    const float hi = tf32_trunc(v);
                    wfull[(size_t)o * K + kk] = v;
                    whi[(size_t)o * K + kk] = hi;
                    wlo[(size_t)o * K + kk] = tf32_trunc(v - hi);
                }
    FCP_TRY(upload(ctx, whi, &cw.w_hi));
    FCP_TRY(upload(ctx, wlo, &cw.w_lo));
    FCP_TRY(pack_f16(ctx, cw, wfull, 7, 32));
    m.conv[name] = cw;
    return FCP_OK;
}

// Direct uint8 stem (conv_tc.cu stem mode): K-block b of 64 = tap rows 2b, 2b+1 of the 7x7 kernel, 32 slots per tap row =
// 7 horizontal taps x 3 channels in MEMORY order (R,G,B: the BGR flip of retinaface.py:450 is applied here), 21 used.
int pack_stem_direct(fcp_ctx* ctx, Model& m, const std::string& conv, const std::string& name) {
    const HostTensor* w = find(m, conv + ".weight");
    auto it = m.conv.find(conv);
    if (!w || w->shape.size() != 4 || w->shape[1] != 3 || w->shape[2] != 7 || w->shape[3] != 7 || it == m.conv.end())
        return fail(ctx, FCP_ERR_STATE, "stem weights missing or not 7x7x3: " + conv);
    ConvWeights cw = it->second;                       // shares scale / shift (folded BN) with the square packing
    const std::vector<float>& scale = m.vec[conv + ".scale"];
    if ((int)scale.size() < cw.cout || cw.cout_pad != 64) return fail(ctx, FCP_ERR_STATE, "stem scale missing: " + conv);
    cw.cin = 64; cw.k = 7; cw.kh = 4; cw.kw = 1; cw.alg_k = 147;
    cw.w_kn = cw.w_hi = cw.w_lo = nullptr;             // exists for the f16x3 tensor-core mode only
    std::vector<float> wfull((size_t)cw.cout_pad * 4 * 64, 0.f);
    for (int o = 0; o < cw.cout; ++o)
        for (int r = 0; r < 7; ++r)
            for (int sx = 0; sx < 7; ++sx)
                for (int mch = 0; mch < 3; ++mch) {
                    const int c = 2 - mch;             // conv input channel of memory channel mch (x[:, [2, 1, 0]])
                    const float v = w->data[(((size_t)o * 3 + c) * 7 + r) * 7 + sx] * scale[o];
                    wfull[((size_t)o * 4 + r / 2) * 64 + (r % 2) * 32 + sx * 3 + mch] = v;
                }
    FCP_TRY(pack_f16(ctx, cw, wfull, 4, 64));
    m.conv[name] = cw;
    return FCP_OK;
}

// ------------------------------------------------------------------------------------- host/device staging
bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int DevIn::init(fcp_ctx* ctx, const void* src, size_t bytes) {
    ctx_ = ctx;
    if (!src || bytes == 0) { p_ = nullptr; return FCP_OK; }
    if (is_device_ptr(src)) { p_ = const_cast<void*>(src); return FCP_OK; }
    FCP_CUDA(ctx, cudaMallocAsync(&p_, bytes, ctx->stream));
    owned_ = true;
    FCP_CUDA(ctx, cudaMemcpyAsync(p_, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return FCP_OK;
}
DevIn::~DevIn() {
    if (owned_ && p_) cudaFreeAsync(p_, ctx_->stream);
}

int DevOut::init(fcp_ctx* ctx, void* dst, size_t bytes, bool need_scratch) {
    ctx_ = ctx; dst_ = dst; bytes_ = bytes;
    if (bytes == 0) return FCP_OK;
    if (dst && is_device_ptr(dst)) { p_ = dst; return FCP_OK; }
    if (!dst && !need_scratch) return FCP_OK;
    FCP_CUDA(ctx, cudaMallocAsync(&p_, bytes, ctx->stream));
    owned_ = true;
    return FCP_OK;
}
int DevOut::flush(size_t bytes) {
    if (owned_ && dst_ && p_) {
        size_t nbytes = bytes == (size_t)-1 ? bytes_ : std::min(bytes, bytes_);
        if (nbytes) FCP_CUDA(ctx_, cudaMemcpyAsync(dst_, p_, nbytes, cudaMemcpyDeviceToHost, ctx_->stream));
    }
    return FCP_OK;
}
DevOut::~DevOut() {
    if (owned_ && p_) cudaFreeAsync(p_, ctx_->stream);
}

int run_conv(fcp_ctx* ctx, const ConvOp& op) {
    const bool tc = op.impl >= 1 && conv_tc_supported(op);   // shapes the tensor-core kernel does not cover use the CUDA-core kernel
    if (!tc && !op.wt->w_kn) return fail(ctx, FCP_ERR_INVALID, "conv: this packing exists for the tensor-core kernel only");
    if (!tc && (op.act_cols < op.wt->cout || op.out_add || op.in2.p || op.out_rs || op.out_is))
        return fail(ctx, FCP_ERR_INVALID, "conv: a partial activation / accumulating output needs the tensor-core kernel");
    static const bool log_conv = getenv("FCP_LOG_CONV") != nullptr;      // one line per tensor-core launch, in launch order:
    if (log_conv && tc) {                                                  // lets an ncu capture (-k conv_tc -s N) be matched to layer shapes
        static long long seq = 0;
        const ConvWeights& w = *op.wt;
        fprintf(stderr, "[conv_tc %lld] impl=%d k=%dx%d s=%d cin=%d cout=%d M=%zu res=%d\n", seq++, op.impl, w.kh ? w.kh : w.k, w.kw ? w.kw : w.k,
                op.stride, w.cin, w.cout, op.out.pixels(), (int)(op.res1 || op.res2));
    }
    if (!ctx->profile) return tc ? launch_conv_tc(ctx, op) : launch_conv_ffma(ctx, op);
    if (ctx->prof_used + 2 > ctx->prof_events.size()) {
        cudaEvent_t a, b;
        FCP_CUDA(ctx, cudaEventCreate(&a));
        FCP_CUDA(ctx, cudaEventCreate(&b));
        ctx->prof_events.push_back(a);
        ctx->prof_events.push_back(b);
    }
    cudaEvent_t e0 = ctx->prof_events[ctx->prof_used], e1 = ctx->prof_events[ctx->prof_used + 1];
    ctx->prof_used += 2;
    FCP_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    int s = tc ? launch_conv_tc(ctx, op) : launch_conv_ffma(ctx, op);
    FCP_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    const ConvWeights& w = *op.wt;
    const double M = (double)op.out.n * op.out.h * op.out.w, K = w.alg_k ? (double)w.alg_k : (double)w.k * w.k * w.cin;
    ctx->prof_flops += 2.0 * M * w.cout * K;
    ctx->prof_bytes += (op.stem_src ? 3.0 * op.out.n * op.stem_h * op.stem_w : 4.0 * (double)op.in.pixels() * w.cin) + 4.0 * (M * w.cout + K * w.cout);
    ctx->prof_recs.push_back({(int)M, w.cout, w.alg_k == 147 ? 3 : w.cin, w.k, op.stride, tc ? 1 : 0});   // stems: the reference's 7x7x3
    return s;
}

}  // namespace fcp

// ================================================================================================ C ABI: core
using namespace fcp;

extern "C" {

const char* fcp_version(void) { return "fcp_b200 0.1.0 sm_100a"; }

int fcp_create(int device, fcp_ctx** out) {
    if (!out) return FCP_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return FCP_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return FCP_ERR_CUDA;
    fcp_ctx* ctx = new fcp_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return FCP_ERR_CUDA;
    }
    ctx->own_stream = true;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    const char* tc = getenv("FCP_CONV_IMPL");
    if (tc) ctx->use_tc = atoi(tc);
    if (const char* cm = getenv("FCP_CUBIC")) ctx->cubic_float = std::string(cm) != "opencv" && std::string(cm) != "fixed";
    *out = ctx;
    return FCP_OK;
}

void fcp_destroy(fcp_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (void* p : ctx->device_allocs) cudaFree(p);
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->copy_events) cudaEventDestroy(e);
    for (auto& r : ctx->stage_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    fcp_comm_destroy(ctx);
    if (ctx->gather_send) cudaFree(ctx->gather_send);
    if (ctx->comm_ready) cudaEventDestroy(ctx->comm_ready);
    if (ctx->comm_done) cudaEventDestroy(ctx->comm_done);
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* fcp_last_error(const fcp_ctx* ctx) { return ctx ? ctx->error.c_str() : "null context"; }

int fcp_set_stream(fcp_ctx* ctx, void* cuda_stream) {
    if (!ctx) return FCP_ERR_INVALID;
    if (ctx->own_stream && ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
    return FCP_OK;
}

int fcp_sync(fcp_ctx* ctx) {
    if (!ctx) return FCP_ERR_INVALID;
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FCP_OK;
}

int64_t fcp_launch_count(const fcp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int fcp_set_micro_batch(fcp_ctx* ctx, int detect_images, int parse_faces) {
    if (!ctx || detect_images < 1 || parse_faces < 1) return fail(ctx, FCP_ERR_INVALID, "micro batch sizes must be >= 1");
    ctx->det_mb = detect_images;
    ctx->par_mb = parse_faces;
    return FCP_OK;
}

int fcp_set_conv_impl(fcp_ctx* ctx, int impl) {
    if (!ctx || impl < 0 || impl > 3) return fail(ctx, FCP_ERR_INVALID, "conv impl must be 0, 1, 2 or 3");
    ctx->use_tc = impl;
    return FCP_OK;
}

int fcp_profile(fcp_ctx* ctx, int enable) {
    if (!ctx) return FCP_ERR_INVALID;
    ctx->profile = enable != 0;
    return FCP_OK;
}

int fcp_profile_read(fcp_ctx* ctx, double* out4) {
    if (!ctx || !out4) return fail(ctx, FCP_ERR_INVALID, "fcp_profile_read: bad argument");
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double ms = 0;
    std::map<std::string, std::pair<double, double>> table;   // shape -> (ms, flops)
    std::map<std::string, int> counts;
    for (size_t i = 0; i + 1 < ctx->prof_used; i += 2) {
        float t = 0;
        FCP_CUDA(ctx, cudaEventElapsedTime(&t, ctx->prof_events[i], ctx->prof_events[i + 1]));
        ms += t;
        if (i / 2 < ctx->prof_recs.size()) {
            const auto& r = ctx->prof_recs[i / 2];
            char key[128];
            snprintf(key, sizeof key, "%s k%d s%d cin%-4d cout%-4d M%-8d", r.tc ? "tc  " : "ffma", r.k, r.stride, r.cin, r.cout, r.m);
            table[key].first += t;
            table[key].second += 2.0 * r.m * r.cout * (double)r.k * r.k * r.cin;
            counts[key]++;
        }
    }
    if (getenv("FCP_TRACE")) {
        for (auto& kv : table)
            fprintf(stderr, "[fcp trace] %s n=%-4d ms=%9.3f  %7.2f TFLOP/s  share=%.3f\n", kv.first.c_str(), counts[kv.first],
                    kv.second.first, kv.second.second / kv.second.first / 1e9, kv.second.first / ms);
    }
    ctx->prof_recs.clear();
    out4[0] = ms; out4[1] = (double)(ctx->prof_used / 2); out4[2] = ctx->prof_flops; out4[3] = ctx->prof_bytes;
    ctx->prof_used = 0; ctx->prof_flops = 0; ctx->prof_bytes = 0;
    return FCP_OK;
}

int fcp_profile_stages(fcp_ctx* ctx, double* out_ms8) {
    if (!ctx || !out_ms8) return fail(ctx, FCP_ERR_INVALID, "fcp_profile_stages: bad argument");
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->comm_stream) FCP_CUDA(ctx, cudaStreamSynchronize(ctx->comm_stream));
    for (int i = 0; i < 8; ++i) out_ms8[i] = 0.0;
    for (size_t i = 0; i < ctx->stage_used; ++i) {
        float t = 0;
        const auto& r = ctx->stage_recs[i];
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.stage >= 0 && r.stage < 8) out_ms8[r.stage] += t;
        else cudaGetLastError();
    }
    ctx->stage_used = 0;
    return FCP_OK;
}

int fcp_set_cubic_mode(fcp_ctx* ctx, int floating_point) {
    if (!ctx) return FCP_ERR_INVALID;
    ctx->cubic_float = floating_point != 0;
    return FCP_OK;
}

int fcp_set_enhance(fcp_ctx* ctx, int enable, float min_face_factor) {
    if (!ctx) return FCP_ERR_INVALID;
    ctx->enh_enabled = enable != 0;
    ctx->enh_threshold = min_face_factor;
    return FCP_OK;
}

int fcp_load_tensor(fcp_ctx* ctx, int model, const char* key, const float* host_data, const int64_t* shape, int ndim) {
    if (!ctx || model < 0 || model > 2 || !key || !host_data || ndim < 0 || ndim > 8)
        return fail(ctx, FCP_ERR_INVALID, "fcp_load_tensor: bad argument");
    HostTensor t;
    size_t count = 1;
    for (int i = 0; i < ndim; ++i) {
        t.shape.push_back(shape[i]);
        count *= (size_t)shape[i];
    }
    t.data.assign(host_data, host_data + count);
    ctx->models[model].host[key] = std::move(t);
    ctx->models[model].finalized = false;
    return FCP_OK;
}

int fcp_finalize(fcp_ctx* ctx, int model, int rrdb_blocks) {
    if (!ctx || model < 0 || model > 2) return fail(ctx, FCP_ERR_INVALID, "fcp_finalize: bad model id");
    FCP_CUDA(ctx, cudaSetDevice(ctx->device));
    Model& m = ctx->models[model];
    m.conv.clear();
    m.rrdb_blocks = rrdb_blocks > 0 ? rrdb_blocks : 23;
    int s = model == FCP_MODEL_RETINAFACE ? finalize_retinaface(ctx)
          : model == FCP_MODEL_BISENET  ? finalize_bisenet(ctx) : finalize_rrdbnet(ctx);
    if (s != FCP_OK) return s;
    m.finalized = true;
    m.host.clear();   // host copies are no longer needed
    return FCP_OK;
}

}  // extern "C"
