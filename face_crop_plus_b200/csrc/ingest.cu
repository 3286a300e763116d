// ingest.cu — batch ingest on the device: replaces utils.as_batch (utils.py:273-342), i.e. per image
//   cv2.resize(img, (ww, hh), INTER_AREA if max(h, w) > max(size) else INTER_CUBIC)  ->  cv2.copyMakeBorder(centre)
// for a ragged list of uint8 RGB images, writing the uint8 NHWC batch the detector consumes.
//
// The arithmetic restated is OpenCV's own (imgproc/resize.cpp, not vendored by the reference):
//   INTER_AREA, integer scale : int32 block sum; 2x2 -> (s + 2) >> 2, otherwise cvRound(float(s) * (1.f / area))
//   INTER_AREA, fractional    : DecimateAlpha tables built in float64 (host, a few hundred entries per image);
//                               buf = buf + S*alpha over x, sum = (sum +) beta*buf over y, float32, mul and add rounded
//                               separately, cvRound
//   INTER_CUBIC               : float32 coefficients (A = -0.75) quantised to 11 bits, exact int32 horizontal pass,
//                               vertical pass in float32 ((S0*b0 + (S1*b1 + (S2*b2 + S3*b3))), cvRound) for the first
//                               8*floor(3*width/8) values of a row - what OpenCV's SSE code does - and in integers
//                               ((v + 2^21) >> 22) for the tail
//   INTER_CUBIC, float (default): what the opencv-python x86 wheels really compute for 8-bit images - they pass the call to
//                               Intel IPP: the separable Keys cubic (a = -0.75, clamped taps) in floating point, rounded to
//                               nearest.  Evaluated here in float64 (horizontal, then vertical; mul and add rounded
//                               separately): one grey level off cv2-with-IPP in < 1e-5 of the bytes (the fixed-point
//                               path above: 4.5 %).  fcp_set_cubic_mode(ctx, 0) selects OpenCV's own fixed-point code.
//   copyMakeBorder            : constant 0 / replicate / reflect / wrap / reflect_101 (borderInterpolate)
// One thread per output pixel (3 channels); the kernel is HBM-bound: it reads every source pixel about once (a few
// times through L1/L2 for the overlapping taps) and writes 3 bytes per output pixel.
#include <cmath>
#include <cstdint>
#include <vector>

#include "common.h"

namespace fcp {

namespace {

enum { ING_COPY = 0, ING_AREA_2X2 = 1, ING_AREA_INT = 2, ING_AREA_GEN = 3, ING_CUBIC = 4, ING_CUBIC_FLOAT = 5 };

struct IngestImage {
    const uint8_t* src;      // u8 [h, w, 3]
    int h, w;                // source size
    int nw, nh;              // resized size
    int top, left;           // padding in front
    int mode;
    int ix, iy;              // ING_AREA_INT: block size
    float inv_area;          // ING_AREA_INT: 1.f / (ix * iy)
    int xt, yt;              // offsets of this image's x / y tables (meaning depends on the mode, see build_tables)
    int simd_end;            // ING_CUBIC: first flat column (x*3 + c) of the integer tail
};

__device__ __forceinline__ int border_index(int p, int n, int mode) {   // OpenCV borderInterpolate
    if ((unsigned)p < (unsigned)n) return p;
    if (mode == FCP_BORDER_REPLICATE) return p < 0 ? 0 : n - 1;
    if (mode == FCP_BORDER_WRAP) { p %= n; return p < 0 ? p + n : p; }
    if (n == 1) return 0;
    const int delta = mode == FCP_BORDER_REFLECT_101 ? 1 : 0;
    do {
        if (p < 0) p = -p - 1 + delta;
        else p = n - 1 - (p - n) - delta;
    } while ((unsigned)p >= (unsigned)n);
    return p;
}

__device__ __forceinline__ uint8_t sat_u8(int v) { return (uint8_t)min(max(v, 0), 255); }

// tables: ti = int table, tf = float table
//   ING_AREA_GEN: ti[xt + d], ti[xt + d + 1] = range of entries of destination column d in the entry arrays that start at
//                 ti[xt + nw + 1 + e] (source index) and tf[... same index] (weight); same for rows with yt / nh.
//   ING_CUBIC   : ti[xt + 5*d] = source offset of the second tap, ti[xt + 5*d + 1..4] = 11-bit weights; rows likewise.
//   ING_CUBIC_FLOAT: ti[xt + d] = source offset of the second tap, td[xt*4 .. ] = four float64 weights per destination index
__global__ void ingest_kernel(const IngestImage* __restrict__ imgs, const int* __restrict__ ti, const float* __restrict__ tf,
                              const double* __restrict__ td, int size_w, int size_h, int border_mode, uint8_t* __restrict__ out) {
    const int X = blockIdx.x * blockDim.x + threadIdx.x, Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (X >= size_w || Y >= size_h) return;
    const IngestImage im = imgs[blockIdx.z];
    uint8_t* dst = out + (((size_t)blockIdx.z * size_h + Y) * size_w + X) * 3;
    int rx = X - im.left, ry = Y - im.top;
    if ((unsigned)rx >= (unsigned)im.nw || (unsigned)ry >= (unsigned)im.nh) {
        if (border_mode == FCP_BORDER_CONSTANT) { dst[0] = dst[1] = dst[2] = 0; return; }
        rx = border_index(rx, im.nw, border_mode);
        ry = border_index(ry, im.nh, border_mode);
    }
    const uint8_t* src = im.src;
    const int W3 = im.w * 3;
    if (im.mode == ING_COPY) {
        const uint8_t* s = src + (size_t)ry * W3 + rx * 3;
        dst[0] = s[0]; dst[1] = s[1]; dst[2] = s[2];
    } else if (im.mode == ING_AREA_2X2 || im.mode == ING_AREA_INT) {
        int s0 = 0, s1 = 0, s2 = 0;
        for (int y = 0; y < im.iy; ++y) {
            const uint8_t* s = src + (size_t)(ry * im.iy + y) * W3 + rx * im.ix * 3;
            for (int x = 0; x < im.ix; ++x) { s0 += s[3 * x]; s1 += s[3 * x + 1]; s2 += s[3 * x + 2]; }
        }
        if (im.mode == ING_AREA_2X2) {
            dst[0] = (uint8_t)((s0 + 2) >> 2); dst[1] = (uint8_t)((s1 + 2) >> 2); dst[2] = (uint8_t)((s2 + 2) >> 2);
        } else {
            dst[0] = sat_u8(__float2int_rn(__fmul_rn((float)s0, im.inv_area)));
            dst[1] = sat_u8(__float2int_rn(__fmul_rn((float)s1, im.inv_area)));
            dst[2] = sat_u8(__float2int_rn(__fmul_rn((float)s2, im.inv_area)));
        }
    } else if (im.mode == ING_AREA_GEN) {
        const int xe0 = ti[im.xt + rx], xe1 = ti[im.xt + rx + 1], xbase = im.xt + im.nw + 1;
        const int ye0 = ti[im.yt + ry], ye1 = ti[im.yt + ry + 1], ybase = im.yt + im.nh + 1;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int ye = ye0; ye < ye1; ++ye) {
            const uint8_t* row = src + (size_t)ti[ybase + ye] * W3;
            const float beta = tf[ybase + ye];
            float b0 = 0.f, b1 = 0.f, b2 = 0.f;
            for (int xe = xe0; xe < xe1; ++xe) {
                const uint8_t* s = row + ti[xbase + xe] * 3;
                const float alpha = tf[xbase + xe];
                b0 = __fadd_rn(b0, __fmul_rn((float)s[0], alpha));
                b1 = __fadd_rn(b1, __fmul_rn((float)s[1], alpha));
                b2 = __fadd_rn(b2, __fmul_rn((float)s[2], alpha));
            }
            if (ye == ye0) { a0 = __fmul_rn(beta, b0); a1 = __fmul_rn(beta, b1); a2 = __fmul_rn(beta, b2); }
            else { a0 = __fadd_rn(a0, __fmul_rn(beta, b0)); a1 = __fadd_rn(a1, __fmul_rn(beta, b1)); a2 = __fadd_rn(a2, __fmul_rn(beta, b2)); }
        }
        dst[0] = sat_u8(__float2int_rn(a0)); dst[1] = sat_u8(__float2int_rn(a1)); dst[2] = sat_u8(__float2int_rn(a2));
    } else if (im.mode == ING_CUBIC_FLOAT) {
        const int sx0 = ti[im.xt + rx], sy0 = ti[im.yt + ry];
        const double* cx = td + (size_t)(im.xt + rx) * 4;
        const double* cy = td + (size_t)(im.yt + ry) * 4;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint8_t* row = src + (size_t)min(max(sy0 - 1 + k, 0), im.h - 1) * W3;
            double h0 = 0.0, h1 = 0.0, h2 = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint8_t* s = row + min(max(sx0 - 1 + j, 0), im.w - 1) * 3;
                h0 = __dadd_rn(h0, __dmul_rn((double)s[0], cx[j]));
                h1 = __dadd_rn(h1, __dmul_rn((double)s[1], cx[j]));
                h2 = __dadd_rn(h2, __dmul_rn((double)s[2], cx[j]));
            }
            v0 = __dadd_rn(v0, __dmul_rn(h0, cy[k]));
            v1 = __dadd_rn(v1, __dmul_rn(h1, cy[k]));
            v2 = __dadd_rn(v2, __dmul_rn(h2, cy[k]));
        }
        dst[0] = sat_u8(__double2int_rn(v0)); dst[1] = sat_u8(__double2int_rn(v1)); dst[2] = sat_u8(__double2int_rn(v2));
    } else {   // ING_CUBIC
        const int* xe = ti + im.xt + 5 * rx;
        const int* ye = ti + im.yt + 5 * ry;
        int hor[4][3];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int sy = min(max(ye[0] - 1 + k, 0), im.h - 1);
            const uint8_t* row = src + (size_t)sy * W3;
            int h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int sx = min(max(xe[0] - 1 + j, 0), im.w - 1);
                const int a = xe[1 + j];
                h0 += row[sx * 3] * a; h1 += row[sx * 3 + 1] * a; h2 += row[sx * 3 + 2] * a;
            }
            hor[k][0] = h0; hor[k][1] = h1; hor[k][2] = h2;
        }
        const float scale = 1.f / (2048.f * 2048.f);
        const float b0 = (float)ye[1] * scale, b1 = (float)ye[2] * scale, b2 = (float)ye[3] * scale, b3 = (float)ye[4] * scale;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int v;
            if (rx * 3 + c < im.simd_end) {
                float t = __fmul_rn((float)hor[3][c], b3);
                t = __fadd_rn(__fmul_rn((float)hor[2][c], b2), t);
                t = __fadd_rn(__fmul_rn((float)hor[1][c], b1), t);
                t = __fadd_rn(__fmul_rn((float)hor[0][c], b0), t);
                v = __float2int_rn(t);
            } else {
                v = (hor[0][c] * ye[1] + hor[1][c] * ye[2] + hor[2][c] * ye[3] + hor[3][c] * ye[4] + (1 << 21)) >> 22;
            }
            dst[c] = sat_u8(v);
        }
    }
}

// ------------------------------------------------------------------------------------------- host: tables
void area_table(int ssize, int dsize, double scale, std::vector<int>& ti, std::vector<float>& tf) {
    // computeResizeAreaTab; layout: [dsize + 1 entry offsets][entries: source index | weight]
    const size_t head = ti.size();
    ti.resize(head + dsize + 1);
    tf.resize(head + dsize + 1);
    std::vector<int> si; std::vector<float> al;
    for (int d = 0; d < dsize; ++d) {
        ti[head + d] = (int)si.size();
        const double f1 = d * scale, f2 = f1 + scale, cell = std::min(scale, ssize - f1);
        int s1 = (int)std::ceil(f1), s2 = (int)std::floor(f2);
        s2 = std::min(s2, ssize - 1);
        s1 = std::min(s1, s2);
        if (s1 - f1 > 1e-3) { si.push_back(s1 - 1); al.push_back((float)((s1 - f1) / cell)); }
        for (int s = s1; s < s2; ++s) { si.push_back(s); al.push_back((float)(1.0 / cell)); }
        if (f2 - s2 > 1e-3) { si.push_back(s2); al.push_back((float)(std::min(std::min(f2 - s2, 1.), cell) / cell)); }
    }
    ti[head + dsize] = (int)si.size();
    for (size_t e = 0; e < si.size(); ++e) { ti.push_back(si[e]); tf.push_back(al[e]); }
}

void cubic_table(int ssize, int dsize, std::vector<int>& ti, std::vector<float>& tf) {
    const double scale = 1.0 / ((double)dsize / ssize);
    for (int d = 0; d < dsize; ++d) {
        float fx = (float)((d + 0.5) * scale - 0.5);
        const int sx = (int)std::floor(fx);
        fx -= sx;
        const float A = -0.75f, x = fx;
        float c[4];
        c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
        c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        c[3] = 1.f - c[0] - c[1] - c[2];
        ti.push_back(sx);
        for (int k = 0; k < 4; ++k) {
            long v = lrintf(c[k] * 2048.f);                       // saturate_cast<short>(float): round half to even
            ti.push_back((int)std::min(std::max(v, -32768L), 32767L));
        }
        tf.resize(ti.size());
    }
}

// float64 Keys weights (interpolateCubic's formulas, A = -0.75) + source offset per destination index; ti and td advance together
void cubic_table_f64(int ssize, int dsize, std::vector<int>& ti, std::vector<double>& td) {
    const double scale = 1.0 / ((double)dsize / ssize);
    td.resize(ti.size() * 4, 0.0);
    for (int d = 0; d < dsize; ++d) {
        const double f = (d + 0.5) * scale - 0.5;
        const int sx = (int)std::floor(f);
        const double x = f - sx, A = -0.75;
        const double c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        const double c1 = ((A + 2) * x - (A + 3)) * x * x + 1;
        const double c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        ti.push_back(sx);
        td.push_back(c0); td.push_back(c1); td.push_back(c2); td.push_back(1.0 - c0 - c1 - c2);
    }
}

}  // namespace

// utils.py:317-331: (new_w, new_h, unscale, paddings) of one image; python float arithmetic == C double arithmetic
void ingest_plan(int h, int w, int size_w, int size_h, int* nw, int* nh, double* unscale, int32_t pad[4]) {
    const double rw = (double)size_w / w, rh = (double)size_h / h;
    if (rw < rh) {
        *unscale = rw; *nw = size_w; *nh = (int)(h * rw);
        pad[0] = (size_h - *nh) / 2; pad[1] = (size_h - *nh + 1) / 2; pad[2] = pad[3] = 0;
    } else {
        *unscale = rh; *nw = (int)(w * rh); *nh = size_h;
        pad[0] = pad[1] = 0; pad[2] = (size_w - *nw) / 2; pad[3] = (size_w - *nw + 1) / 2;
    }
}

int launch_ingest(fcp_ctx* ctx, const uint8_t* const* dev_ptrs, const int32_t* hs, const int32_t* ws, int n, int size_w,
                  int size_h, int border_mode, uint8_t* out, double* unscales, int32_t* paddings) {
    std::vector<IngestImage> imgs(n);
    std::vector<int> ti; std::vector<float> tf; std::vector<double> td;
    for (int i = 0; i < n; ++i) {
        IngestImage& im = imgs[i];
        im = IngestImage{};
        im.src = dev_ptrs[i]; im.h = hs[i]; im.w = ws[i];
        double unscale; int32_t pad[4];
        ingest_plan(im.h, im.w, size_w, size_h, &im.nw, &im.nh, &unscale, pad);
        if (im.nw < 1 || im.nh < 1) return fail(ctx, FCP_ERR_INVALID, "fcp_as_batch: an image collapses to zero pixels at this size");
        im.top = pad[0]; im.left = pad[2];
        if (unscales) unscales[i] = unscale;
        if (paddings) for (int k = 0; k < 4; ++k) paddings[4 * i + k] = pad[k];
        if (im.nw == im.w && im.nh == im.h) { im.mode = ING_COPY; continue; }
        if (std::max(im.h, im.w) > std::max(size_w, size_h)) {                      // utils.py:320 -> INTER_AREA
            const double sx = 1.0 / ((double)im.nw / im.w), sy = 1.0 / ((double)im.nh / im.h);
            if (sx < 1.0 || sy < 1.0) return fail(ctx, FCP_ERR_INVALID, "fcp_as_batch: INTER_AREA enlargement is not a path as_batch can take");
            const int ix = (int)lrint(sx), iy = (int)lrint(sy);
            if (std::abs(sx - ix) < 2.220446049250313e-16 && std::abs(sy - iy) < 2.220446049250313e-16) {
                im.mode = (ix == 2 && iy == 2) ? ING_AREA_2X2 : ING_AREA_INT;
                im.ix = ix; im.iy = iy; im.inv_area = 1.f / (float)(ix * iy);
            } else {
                im.mode = ING_AREA_GEN;
                im.xt = (int)ti.size(); area_table(im.w, im.nw, sx, ti, tf);
                im.yt = (int)ti.size(); area_table(im.h, im.nh, sy, ti, tf);
            }
        } else if (ctx->cubic_float) {
            im.mode = ING_CUBIC_FLOAT;
            im.xt = (int)ti.size(); cubic_table_f64(im.w, im.nw, ti, td);
            im.yt = (int)ti.size(); cubic_table_f64(im.h, im.nh, ti, td);
            tf.resize(ti.size());
        } else {
            im.mode = ING_CUBIC;
            im.xt = (int)ti.size(); cubic_table(im.w, im.nw, ti, tf);
            im.yt = (int)ti.size(); cubic_table(im.h, im.nh, ti, tf);
            im.simd_end = (im.nw * 3) / 8 * 8;
        }
    }
    if (ti.empty()) { ti.push_back(0); tf.push_back(0.f); }
    td.resize(ti.size() * 4, 0.0);
    // the area tables index the entry arrays relative to their own start: rebase is done in the kernel (xbase / ybase)
    void *d_imgs = nullptr, *d_ti = nullptr, *d_tf = nullptr, *d_td = nullptr;
    FCP_CUDA(ctx, cudaMallocAsync(&d_imgs, sizeof(IngestImage) * n, ctx->stream));
    FCP_CUDA(ctx, cudaMallocAsync(&d_ti, sizeof(int) * ti.size(), ctx->stream));
    FCP_CUDA(ctx, cudaMallocAsync(&d_tf, sizeof(float) * tf.size(), ctx->stream));
    FCP_CUDA(ctx, cudaMallocAsync(&d_td, sizeof(double) * td.size(), ctx->stream));
    FCP_CUDA(ctx, cudaMemcpyAsync(d_td, td.data(), sizeof(double) * td.size(), cudaMemcpyHostToDevice, ctx->stream));
    FCP_CUDA(ctx, cudaMemcpyAsync(d_imgs, imgs.data(), sizeof(IngestImage) * n, cudaMemcpyHostToDevice, ctx->stream));
    FCP_CUDA(ctx, cudaMemcpyAsync(d_ti, ti.data(), sizeof(int) * ti.size(), cudaMemcpyHostToDevice, ctx->stream));
    FCP_CUDA(ctx, cudaMemcpyAsync(d_tf, tf.data(), sizeof(float) * tf.size(), cudaMemcpyHostToDevice, ctx->stream));
    const dim3 block(32, 8), grid((size_w + 31) / 32, (size_h + 7) / 8, n);
    ingest_kernel<<<grid, block, 0, ctx->stream>>>(static_cast<const IngestImage*>(d_imgs), static_cast<const int*>(d_ti),
                                                    static_cast<const float*>(d_tf), static_cast<const double*>(d_td), size_w, size_h, border_mode, out);
    FCP_KERNEL_CHECK(ctx);
    FCP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));              // the pageable host tables above must outlive the copies
    FCP_CUDA(ctx, cudaFreeAsync(d_imgs, ctx->stream));
    FCP_CUDA(ctx, cudaFreeAsync(d_ti, ctx->stream));
    FCP_CUDA(ctx, cudaFreeAsync(d_tf, ctx->stream));
    FCP_CUDA(ctx, cudaFreeAsync(d_td, ctx->stream));
    return FCP_OK;
}

}  // namespace fcp
