// parse.cu — BiSeNet pre/post-processing and the RRDBNet output tail.  All HBM-bound, no tensor cores.
//
//  parse_prep : crops u8 -> /255 -> bilinear (align_corners=False) to 512x512 -> (x-mean)/std   (bise.py:387-392)
//  parse_tail : logits at 1/8 res -> bilinear(align_corners=True) to 512x512 (bise.py:212) -> nearest to (h,w) ->
//               argmax (bise.py:394), evaluated ONLY at the pixels the nearest resize picks (the reference
//               materialises 19x512x512 floats per face instead) + per-class pixel histogram (bise.py:253,314)
//  masks      : 0/255 mask of a class set (bise.py:310-316)
//  rrdb_tail  : conv_last output -> bicubic x0.25 (fixed 4x4 stencil) -> clamp(0,1)*255 -> round  (rrdb.py:143-144)
#include "common.h"

namespace fcp {

namespace {

// ATen area_pixel_compute_source_index(align_corners=False) + guard_index_and_lambda, float32
__device__ __forceinline__ void linear_taps(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
    float real = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
    real = real < 0.f ? 0.f : real;
    i0 = min((int)real, in_size - 1);
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = fminf(fmaxf(__fsub_rn(real, (float)i0), 0.f), 1.f);
    l0 = __fsub_rn(1.f, l1);
}

__global__ void parse_prep_kernel(const uint8_t* __restrict__ crops, int F, int h, int w, int OH, int OW,
                                  float* __restrict__ out) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)F * OH * OW;
    if (idx >= total) return;
    int ox = idx % OW;
    size_t t = idx / OW;
    int oy = t % OH;
    int f = t / OH;
    const float sy = (float)h / (float)OH, sx = (float)w / (float)OW;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    linear_taps(oy, sy, h, y0, y1, ly0, ly1);
    linear_taps(ox, sx, w, x0, x1, lx0, lx1);
    const uint8_t* base = crops + (size_t)f * h * w * 3;
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v00 = __fdiv_rn((float)base[((size_t)y0 * w + x0) * 3 + c], 255.f);
        float v01 = __fdiv_rn((float)base[((size_t)y0 * w + x1) * 3 + c], 255.f);
        float v10 = __fdiv_rn((float)base[((size_t)y1 * w + x0) * 3 + c], 255.f);
        float v11 = __fdiv_rn((float)base[((size_t)y1 * w + x1) * 3 + c], 255.f);
        float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
        float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
        float v = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
        o[c] = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
    }
    float* d = out + idx * 3;
    d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
}

struct TailArgs {
    const float* logits; int nhwc, cs;   // nhwc: [f,fh,fw,cs] with 19 valid channels; else NCHW [f,19,fh,fw]
    int F, fh, fw, IH, IW, h, w;
    uint8_t* labels; int* hist;
};

__global__ void __launch_bounds__(256) parse_tail_kernel(const TailArgs a) {
    __shared__ int shist[19];
    const int f = blockIdx.y;
    if (threadIdx.x < 19) shist[threadIdx.x] = 0;
    __syncthreads();
    int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix < a.h * a.w) {
        int ox = pix % a.w, oy = pix / a.w;
        // nearest: src = min(floor(dst * (float)in/out), in-1)
        int iy = min((int)floorf(__fmul_rn((float)oy, (float)a.IH / (float)a.h)), a.IH - 1);
        int ix = min((int)floorf(__fmul_rn((float)ox, (float)a.IW / (float)a.w)), a.IW - 1);
        // bilinear align_corners=True from (fh,fw) to (IH,IW)
        float scy = a.IH > 1 ? (float)(a.fh - 1) / (float)(a.IH - 1) : 0.f;
        float scx = a.IW > 1 ? (float)(a.fw - 1) / (float)(a.IW - 1) : 0.f;
        float ry = __fmul_rn(scy, (float)iy), rx = __fmul_rn(scx, (float)ix);
        int y0 = min((int)ry, a.fh - 1), x0 = min((int)rx, a.fw - 1);
        int y1 = y0 + (y0 < a.fh - 1 ? 1 : 0), x1 = x0 + (x0 < a.fw - 1 ? 1 : 0);
        float ly1 = fminf(fmaxf(__fsub_rn(ry, (float)y0), 0.f), 1.f), lx1 = fminf(fmaxf(__fsub_rn(rx, (float)x0), 0.f), 1.f);
        float ly0 = __fsub_rn(1.f, ly1), lx0 = __fsub_rn(1.f, lx1);
        int best = 0;
        float bestv = -INFINITY;
        size_t p00, p01, p10, p11, cstep;
        if (a.nhwc) {
            size_t fb = (size_t)f * a.fh * a.fw;
            p00 = (fb + (size_t)y0 * a.fw + x0) * a.cs; p01 = (fb + (size_t)y0 * a.fw + x1) * a.cs;
            p10 = (fb + (size_t)y1 * a.fw + x0) * a.cs; p11 = (fb + (size_t)y1 * a.fw + x1) * a.cs;
            cstep = 1;
        } else {
            size_t fb = (size_t)f * 19 * a.fh * a.fw;
            p00 = fb + (size_t)y0 * a.fw + x0; p01 = fb + (size_t)y0 * a.fw + x1;
            p10 = fb + (size_t)y1 * a.fw + x0; p11 = fb + (size_t)y1 * a.fw + x1;
            cstep = (size_t)a.fh * a.fw;
        }
#pragma unroll
        for (int c = 0; c < 19; ++c) {
            float v00 = a.logits[p00 + c * cstep], v01 = a.logits[p01 + c * cstep];
            float v10 = a.logits[p10 + c * cstep], v11 = a.logits[p11 + c * cstep];
            float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
            float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
            float v = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
            if (v > bestv) { bestv = v; best = c; }            // first maximum wins, like argmax
        }
        if (a.labels) a.labels[(size_t)f * a.h * a.w + pix] = (uint8_t)best;
        atomicAdd(&shist[best], 1);
    }
    __syncthreads();
    if (a.hist && threadIdx.x < 19 && shist[threadIdx.x]) atomicAdd(&a.hist[f * 19 + threadIdx.x], shist[threadIdx.x]);
}

__global__ void masks_kernel(const uint8_t* __restrict__ labels, size_t count, const uint8_t* __restrict__ lut,
                             uint8_t* __restrict__ out) {
    __shared__ uint8_t sl[32];
    if (threadIdx.x < 19) sl[threadIdx.x] = lut[threadIdx.x] ? 255 : 0;
    __syncthreads();
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < count) {
        uchar4 v = *reinterpret_cast<const uchar4*>(labels + i);
        uchar4 o = make_uchar4(sl[v.x], sl[v.y], sl[v.z], sl[v.w]);
        *reinterpret_cast<uchar4*>(out + i) = o;
    } else {
        for (; i < count; ++i) out[i] = sl[labels[i]];
    }
}

// group_by_attributes / group_by_masks membership (bise.py:214-325) from the per-class pixel histogram: integer compares.
//   attribute group g = codes[offs[g] .. offs[g+1]) of signed class ids: a > 0 -> count(|a|) > thr, else count(|a|) <= thr,
//   joined by AND (attr_join_by_and) or OR;  mask group m: sum of hist over the classes with lut[m][c] != 0 > mask_thr.
__global__ void group_kernel(const int32_t* __restrict__ hist, int f, const int32_t* __restrict__ codes,
                             const int32_t* __restrict__ offs, int n_attr, int attr_thr, int join_and,
                             const uint8_t* __restrict__ lut, int n_mask, int mask_thr, uint8_t* __restrict__ out_attr,
                             uint8_t* __restrict__ out_mask) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= f * (n_attr + n_mask)) return;
    const int g = idx / f, face = idx - g * f;
    const int32_t* h = hist + face * 19;
    if (g < n_attr) {
        bool all = true, any = false;
        for (int e = offs[g]; e < offs[g + 1]; ++e) {
            const int a = codes[e], c = a < 0 ? -a : a;
            const int cnt = c < 19 ? h[c] : 0;
            const bool t = a > 0 ? cnt > attr_thr : cnt <= attr_thr;
            all &= t; any |= t;
        }
        out_attr[g * f + face] = (join_and ? all : any) ? 1 : 0;
    } else {
        const int m = g - n_attr;
        int sum = 0;
        for (int c = 0; c < 19; ++c) sum += lut[m * 19 + c] ? h[c] : 0;
        out_mask[m * f + face] = sum > mask_thr ? 1 : 0;
    }
}

// all mask groups in one pass over the labels: out[m][i] = 255 where lut[m][label[i]] != 0   (n_mask <= 32)
__global__ void multi_masks_kernel(const uint8_t* __restrict__ labels, size_t count, const uint8_t* __restrict__ lut, int n_mask,
                                   uint8_t* __restrict__ out) {
    __shared__ uint32_t bits[32];                         // bits[class] = set of groups containing the class
    if (threadIdx.x < 32) {
        uint32_t b = 0;
        if (threadIdx.x < 19)
            for (int m = 0; m < n_mask; ++m) b |= lut[m * 19 + threadIdx.x] ? (1u << m) : 0u;
        bits[threadIdx.x] = b;
    }
    __syncthreads();
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= count) return;
    if (i + 3 < count) {
        const uchar4 v = *reinterpret_cast<const uchar4*>(labels + i);
        const uint32_t b0 = bits[v.x & 31], b1 = bits[v.y & 31], b2 = bits[v.z & 31], b3 = bits[v.w & 31];
        for (int m = 0; m < n_mask; ++m)
            *reinterpret_cast<uchar4*>(out + (size_t)m * count + i) =
                make_uchar4((b0 >> m) & 1 ? 255 : 0, (b1 >> m) & 1 ? 255 : 0, (b2 >> m) & 1 ? 255 : 0, (b3 >> m) & 1 ? 255 : 0);
    } else {
        for (size_t j = i; j < count; ++j)
            for (int m = 0; m < n_mask; ++m) out[(size_t)m * count + j] = (bits[labels[j] & 31] >> m) & 1 ? 255 : 0;
    }
}

// x4: NHWC [n,4h,4w,cs] (3 valid channels) -> out NCHW [n,3,h,w] (image i at out + i*3*h*w)
template <bool U8>
__global__ void rrdb_tail_kernel(const float* __restrict__ x4, int cs, int co, int n, int h, int w,
                                 void* __restrict__ out_v) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)n * h * w;
    if (idx >= total) return;
    int ox = idx % w;
    size_t t = idx / w;
    int oy = t % h;
    int im = t / h;
    const float k[4] = {-0.09375f, 0.59375f, 0.59375f, -0.09375f};
    float acc[3] = {0.f, 0.f, 0.f};
    const int W4 = 4 * w, H4 = 4 * h;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float row[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const float* p = x4 + (((size_t)im * H4 + oy * 4 + r) * W4 + ox * 4 + s) * cs + co;
            row[0] = fmaf(k[s], p[0], row[0]); row[1] = fmaf(k[s], p[1], row[1]); row[2] = fmaf(k[s], p[2], row[2]);
        }
        acc[0] = fmaf(k[r], row[0], acc[0]); acc[1] = fmaf(k[r], row[1], acc[1]); acc[2] = fmaf(k[r], row[2], acc[2]);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float v = fminf(fmaxf(acc[c], 0.f), 1.f);
        if (U8) static_cast<uint8_t*>(out_v)[(((size_t)im * h + oy) * w + ox) * 3 + c] = (uint8_t)rintf(v * 255.f);
        else static_cast<float*>(out_v)[(((size_t)im * 3 + c) * h + oy) * w + ox] = rintf(v * 255.f);
    }
}

}  // namespace

int launch_parse_prep(fcp_ctx* ctx, const uint8_t* crops, int f, int h, int w, float* out_nhwc3) {
    size_t total = (size_t)f * 512 * 512;
    parse_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(crops, f, h, w, 512, 512, out_nhwc3);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_parse_tail(fcp_ctx* ctx, const float* logits, int layout_nhwc, int cs, int f, int fh, int fw, int h, int w,
                      uint8_t* labels, int32_t* hist) {
    if (f == 0) return FCP_OK;
    if (hist) FCP_CUDA(ctx, cudaMemsetAsync(hist, 0, sizeof(int32_t) * 19 * f, ctx->stream));
    TailArgs a{logits, layout_nhwc, cs, f, fh, fw, 512, 512, h, w, labels, hist};
    dim3 grid((h * w + 255) / 256, f);
    parse_tail_kernel<<<grid, 256, 0, ctx->stream>>>(a);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_masks(fcp_ctx* ctx, const uint8_t* labels, size_t count, const uint8_t* lut_dev, uint8_t* out) {
    if (count == 0) return FCP_OK;
    size_t threads = (count + 3) / 4;
    masks_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(labels, count, lut_dev, out);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_group(fcp_ctx* ctx, const int32_t* hist, int f, const int32_t* codes, const int32_t* offs, int n_attr, int attr_thr,
                 int join_and, const uint8_t* lut, int n_mask, int mask_thr, uint8_t* out_attr, uint8_t* out_mask) {
    const int total = f * (n_attr + n_mask);
    if (total == 0) return FCP_OK;
    group_kernel<<<(total + 127) / 128, 128, 0, ctx->stream>>>(hist, f, codes, offs, n_attr, attr_thr, join_and, lut, n_mask, mask_thr,
                                                              out_attr, out_mask);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_multi_masks(fcp_ctx* ctx, const uint8_t* labels, size_t count, const uint8_t* lut_dev, int n_mask, uint8_t* out) {
    if (count == 0 || n_mask == 0) return FCP_OK;
    const size_t threads = (count + 3) / 4;
    multi_masks_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ctx->stream>>>(labels, count, lut_dev, n_mask, out);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_rrdb_tail(fcp_ctx* ctx, Tensor x4, float* out_nchw, int h, int w) {
    size_t total = (size_t)x4.n * h * w;
    rrdb_tail_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(x4.p, x4.cs, x4.co, x4.n, h, w, out_nchw);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_rrdb_tail_u8(fcp_ctx* ctx, Tensor x4, uint8_t* out_nhwc, int h, int w) {
    size_t total = (size_t)x4.n * h * w;
    rrdb_tail_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(x4.p, x4.cs, x4.co, x4.n, h, w, out_nhwc);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

}  // namespace fcp
