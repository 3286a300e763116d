// align.cu — per-face least-squares transform (float64) + OpenCV-exact fixed-point bilinear warpAffine.
//
// Replaces Cropper.crop_align (cropper.py:441-552): cv2.estimateAffinePartial2D / estimateAffine2D with
// ransacReprojThreshold=inf (cropper.py:515-527) and cv2.warpAffine INTER_LINEAR with the 5 CLI-reachable border
// modes (cropper.py:542-547).  The arithmetic restates OpenCV's published algorithm:
// coordinates in 1/1024 px (AB_BITS=10) rounded half-to-even, 1/32 px interpolation grid (INTER_BITS=5), 15-bit
// integer weights.  float64 products/sums use explicit _rn intrinsics: an FMA contraction would change rint() ties.
// HBM-bound gather (reads ~the source footprint, writes out_w*out_h*3 bytes per face); no tensor cores.
#include "common.h"

namespace fcp {

namespace {

// one thread per face.  out: M[f][6] (row-major 2x3), inv[f][6] (i00,i01,i02,i10,i11,i12), valid[f]
__global__ void solve_kernel(const float* __restrict__ lms, const int* __restrict__ face_count, int f,
                             const float* __restrict__ target, int allow_skew, double* __restrict__ M,
                             double* __restrict__ inv, unsigned char* __restrict__ valid) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f) return;
    if (face_count && i >= *face_count) { valid[i] = 0; return; }
    double sx[5], sy[5], dx[5], dy[5];
    double msx = 0, msy = 0, mdx = 0, mdy = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        sx[k] = (double)lms[i * 10 + 2 * k]; sy[k] = (double)lms[i * 10 + 2 * k + 1];
        dx[k] = (double)target[2 * k]; dy[k] = (double)target[2 * k + 1];
        msx = __dadd_rn(msx, sx[k]); msy = __dadd_rn(msy, sy[k]);
        mdx = __dadd_rn(mdx, dx[k]); mdy = __dadd_rn(mdy, dy[k]);
    }
    msx = __ddiv_rn(msx, 5.0); msy = __ddiv_rn(msy, 5.0); mdx = __ddiv_rn(mdx, 5.0); mdy = __ddiv_rn(mdy, 5.0);
    double m[6];
    bool ok = true;
    if (!allow_skew) {
        // 4-dof similarity: [[a,-b,tx],[b,a,ty]]
        double den = 0, na = 0, nb = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            double x = __dsub_rn(sx[k], msx), y = __dsub_rn(sy[k], msy);
            double u = __dsub_rn(dx[k], mdx), v = __dsub_rn(dy[k], mdy);
            den = __dadd_rn(den, __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)));
            na = __dadd_rn(na, __dadd_rn(__dmul_rn(x, u), __dmul_rn(y, v)));
            nb = __dadd_rn(nb, __dsub_rn(__dmul_rn(x, v), __dmul_rn(y, u)));
        }
        ok = den != 0.0 && isfinite(den);
        double a = ok ? __ddiv_rn(na, den) : 0.0, b = ok ? __ddiv_rn(nb, den) : 0.0;
        m[0] = a; m[1] = -b; m[2] = __dsub_rn(mdx, __dsub_rn(__dmul_rn(a, msx), __dmul_rn(b, msy)));
        m[3] = b; m[4] = a;  m[5] = __dsub_rn(mdy, __dadd_rn(__dmul_rn(b, msx), __dmul_rn(a, msy)));
    } else {
        // 6-dof affine: centred normal equations
        double sxx = 0, sxy = 0, syy = 0, bxu = 0, byu = 0, bxv = 0, byv = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            double x = __dsub_rn(sx[k], msx), y = __dsub_rn(sy[k], msy);
            double u = __dsub_rn(dx[k], mdx), v = __dsub_rn(dy[k], mdy);
            sxx = __dadd_rn(sxx, __dmul_rn(x, x)); sxy = __dadd_rn(sxy, __dmul_rn(x, y)); syy = __dadd_rn(syy, __dmul_rn(y, y));
            bxu = __dadd_rn(bxu, __dmul_rn(x, u)); byu = __dadd_rn(byu, __dmul_rn(y, u));
            bxv = __dadd_rn(bxv, __dmul_rn(x, v)); byv = __dadd_rn(byv, __dmul_rn(y, v));
        }
        double det = __dsub_rn(__dmul_rn(sxx, syy), __dmul_rn(sxy, sxy));
        double ref = fmax(__dmul_rn(sxx, syy), 1e-300);
        ok = det != 0.0 && isfinite(det) && !(fabs(det) < 1e-12 * ref);
        double idet = ok ? det : 1.0;
        m[0] = __ddiv_rn(__dsub_rn(__dmul_rn(bxu, syy), __dmul_rn(byu, sxy)), idet);
        m[1] = __ddiv_rn(__dsub_rn(__dmul_rn(byu, sxx), __dmul_rn(bxu, sxy)), idet);
        m[2] = __dsub_rn(__dsub_rn(mdx, __dmul_rn(m[0], msx)), __dmul_rn(m[1], msy));
        m[3] = __ddiv_rn(__dsub_rn(__dmul_rn(bxv, syy), __dmul_rn(byv, sxy)), idet);
        m[4] = __ddiv_rn(__dsub_rn(__dmul_rn(byv, sxx), __dmul_rn(bxv, sxy)), idet);
        m[5] = __dsub_rn(__dsub_rn(mdy, __dmul_rn(m[3], msx)), __dmul_rn(m[4], msy));
    }
    valid[i] = ok ? 1 : 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) M[i * 6 + k] = ok ? m[k] : __longlong_as_double(0x7ff8000000000000LL);
    // inverse map (cv2.warpAffine without WARP_INVERSE_MAP inverts M in double first)
    double D = __dsub_rn(__dmul_rn(m[0], m[4]), __dmul_rn(m[1], m[3]));
    D = D != 0.0 ? __ddiv_rn(1.0, D) : 0.0;
    double i00 = __dmul_rn(m[4], D), i11 = __dmul_rn(m[0], D);
    double i01 = __dmul_rn(m[1], -D), i10 = __dmul_rn(m[3], -D);
    double i02 = __dsub_rn(__dmul_rn(-i00, m[2]), __dmul_rn(i01, m[5]));
    double i12 = __dsub_rn(__dmul_rn(-i10, m[2]), __dmul_rn(i11, m[5]));
    double* o = inv + i * 6;
    o[0] = i00; o[1] = i01; o[2] = i02; o[3] = i10; o[4] = i11; o[5] = i12;
}

// inverse coefficients from caller-supplied matrices (test hook / landmarks given with precomputed M)
__global__ void invert_kernel(const double* __restrict__ M, int f, double* __restrict__ inv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= f) return;
    const double* m = M + i * 6;
    double D = __dsub_rn(__dmul_rn(m[0], m[4]), __dmul_rn(m[1], m[3]));
    D = D != 0.0 ? __ddiv_rn(1.0, D) : 0.0;
    double i00 = __dmul_rn(m[4], D), i11 = __dmul_rn(m[0], D);
    double i01 = __dmul_rn(m[1], -D), i10 = __dmul_rn(m[3], -D);
    double* o = inv + i * 6;
    o[0] = i00; o[1] = i01; o[2] = __dsub_rn(__dmul_rn(-i00, m[2]), __dmul_rn(i01, m[5]));
    o[3] = i10; o[4] = i11; o[5] = __dsub_rn(__dmul_rn(-i10, m[2]), __dmul_rn(i11, m[5]));
}

__device__ __forceinline__ int border_index(int p, int n, int mode) {
    // cv::borderInterpolate
    if ((unsigned)p < (unsigned)n) return p;
    if (mode == FCP_BORDER_REPLICATE) return p < 0 ? 0 : n - 1;
    if (mode == FCP_BORDER_WRAP) {
        int q = p % n;
        return q < 0 ? q + n : q;
    }
    if (n == 1) return 0;
    const int delta = mode == FCP_BORDER_REFLECT_101 ? 1 : 0;
    do {
        if (p < 0) p = -p - 1 + delta;
        else p = n - 1 - (p - n) - delta;
    } while ((unsigned)p >= (unsigned)n);
    return p;
}

struct WarpArgs {
    const uint8_t* images; int n, h, w;          // dense batch (or nullptr)
    const uint8_t* const* ptrs; const int* hs; const int* ws;   // ragged list (or nullptr)
    const int* paddings; const int* indices; const int* face_count;
    const double* inv; const unsigned char* valid;
    int out_w, out_h, border;
    uint8_t* out;
};

__global__ void __launch_bounds__(256) warp_kernel(const WarpArgs a) {
    const int f = blockIdx.z;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= a.out_w || y >= a.out_h) return;
    uint8_t* dst = a.out + (((size_t)f * a.out_h + y) * a.out_w + x) * 3;
    if ((a.face_count && f >= *a.face_count) || !a.valid[f]) {
        dst[0] = 0; dst[1] = 0; dst[2] = 0;
        return;
    }
    const int img = a.indices[f];
    const uint8_t* base;
    int H, W;
    if (a.ptrs) { base = a.ptrs[img]; H = a.hs[img]; W = a.ws[img]; }
    else { base = a.images + (size_t)img * a.h * a.w * 3; H = a.h; W = a.w; }
    size_t pitch = (size_t)W * 3;
    if (a.paddings) {                                  // image[t:H-b, l:W-r]  (cropper.py:536-539)
        int t = a.paddings[img * 4], b = a.paddings[img * 4 + 1], l = a.paddings[img * 4 + 2], r = a.paddings[img * 4 + 3];
        base += (size_t)t * pitch + (size_t)l * 3;
        H -= t + b; W -= l + r;
    }
    const double* iv = a.inv + f * 6;
    int adelta = __double2int_rn(__dmul_rn(__dmul_rn(iv[0], (double)x), 1024.0));
    int bdelta = __double2int_rn(__dmul_rn(__dmul_rn(iv[3], (double)x), 1024.0));
    int X0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(iv[1], (double)y), iv[2]), 1024.0)) + 16;
    int Y0 = __double2int_rn(__dmul_rn(__dadd_rn(__dmul_rn(iv[4], (double)y), iv[5]), 1024.0)) + 16;
    int X = (int)((unsigned)X0 + (unsigned)adelta) >> 5, Y = (int)((unsigned)Y0 + (unsigned)bdelta) >> 5;
    int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
    int fx = X & 31, fy = Y & 31;
    int w00 = 32 * (32 - fx) * (32 - fy), w01 = 32 * fx * (32 - fy), w10 = 32 * (32 - fx) * fy, w11 = 32 * fx * fy;
    int acc0 = 16384, acc1 = 16384, acc2 = 16384;
    if (H > 0 && W > 0) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            int yy = sy + (t >> 1), xx = sx + (t & 1);
            int wgt = t == 0 ? w00 : (t == 1 ? w01 : (t == 2 ? w10 : w11));
            bool inside = (unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)W;
            if (a.border == FCP_BORDER_CONSTANT) {
                if (!inside) continue;
            } else if (!inside) {
                yy = border_index(yy, H, a.border);
                xx = border_index(xx, W, a.border);
            }
            const uint8_t* p = base + (size_t)yy * pitch + (size_t)xx * 3;
            acc0 += wgt * p[0]; acc1 += wgt * p[1]; acc2 += wgt * p[2];
        }
    }
    dst[0] = (uint8_t)(acc0 >> 15); dst[1] = (uint8_t)(acc1 >> 15); dst[2] = (uint8_t)(acc2 >> 15);
}

}  // namespace

int launch_solve(fcp_ctx* ctx, const float* landmarks, const int32_t* face_count_dev, int f, const float* target,
                 int allow_skew, double* matrices, double* inv, uint8_t* valid) {
    if (f == 0) return FCP_OK;
    solve_kernel<<<(f + 127) / 128, 128, 0, ctx->stream>>>(landmarks, face_count_dev, f, target, allow_skew, matrices, inv, valid);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_invert(fcp_ctx* ctx, const double* matrices, int f, double* inv) {
    if (f == 0) return FCP_OK;
    invert_kernel<<<(f + 127) / 128, 128, 0, ctx->stream>>>(matrices, f, inv);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_warp(fcp_ctx* ctx, const uint8_t* images, int n, int h, int w, const uint8_t* const* image_ptrs,
                const int32_t* hs, const int32_t* ws, const int32_t* paddings, const int32_t* indices,
                const int32_t* face_count_dev, int f, const double* inv, const uint8_t* valid, int out_w, int out_h,
                int border, uint8_t* out) {
    if (f == 0) return FCP_OK;
    WarpArgs a{images, n, h, w, image_ptrs, hs, ws, paddings, indices, face_count_dev, inv, valid, out_w, out_h, border, out};
    dim3 grid((out_w + 31) / 32, (out_h + 7) / 8, f), block(32, 8);
    warp_kernel<<<grid, block, 0, ctx->stream>>>(a);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

// N-point -> 5-point landmark reduction (utils.py:90-168, cropper.py:828-831): point j of the result is the float32 mean of
// the source points [lo_j, hi_j) - summed in index order, then divided by the count, exactly what
// ``landmarks[:, s].mean(1)`` does on a float32 [F,K,2] array.
__global__ void reduce_landmarks_kernel(const float* __restrict__ lms, int f, int k, int lo0, int hi0, int lo1, int hi1, int lo2,
                                        int hi2, int lo3, int hi3, int lo4, int hi4, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= f * 10) return;
    const int face = idx / 10, r = idx - face * 10, j = r >> 1, c = r & 1;
    const int lo = j == 0 ? lo0 : j == 1 ? lo1 : j == 2 ? lo2 : j == 3 ? lo3 : lo4;
    const int hi = j == 0 ? hi0 : j == 1 ? hi1 : j == 2 ? hi2 : j == 3 ? hi3 : hi4;
    float acc = 0.f;
    for (int i = lo; i < hi; ++i) acc = __fadd_rn(acc, lms[((size_t)face * k + i) * 2 + c]);
    out[idx] = __fdiv_rn(acc, (float)(hi - lo));
}

int launch_reduce_landmarks(fcp_ctx* ctx, const float* lms, int f, int k, const int bounds[10], float* out) {
    if (f == 0) return FCP_OK;
    reduce_landmarks_kernel<<<(f * 10 + 127) / 128, 128, 0, ctx->stream>>>(lms, f, k, bounds[0], bounds[1], bounds[2], bounds[3], bounds[4],
                                                                         bounds[5], bounds[6], bounds[7], bounds[8], bounds[9], out);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

}  // namespace fcp
