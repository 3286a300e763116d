// misc.cu — the non-GEMM layers of the three graphs: stems (Cin=3), pooling, pooled-vector 1x1 convs,
// channel attention, layout conversion.  All HBM-bound except the stems (fp32 FMA).
#include "common.h"

namespace fcp {

namespace {

__device__ __forceinline__ float act_fn(float v, int act, float slope) {
    if (act == FCP_ACT_RELU) return fmaxf(v, 0.f);
    if (act == FCP_ACT_LRELU) return v > 0.f ? v : v * slope;
    if (act == FCP_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}

// ------------------------------------------------------------------------------------------------ stem 7x7/s2
// body.conv1+bn1+relu (torchvision resnet.py:268-270) and cp.resnet.conv1+bn1+relu (_layers.py:262-263).
// CTA = 4x128 output pixels x 64 channels, 256 threads; thread = 4 pixels (same row, 32 apart) x 32 channels, i.e. 128
// fp32 accumulators.  The input patch (13 x 261 x 3) lives in shared memory as three channel planes so that a warp's
// float2 loads (columns 2x, 2x+1 of consecutive pixels) are conflict-free; the 147x64 filter bank is read with
// warp-broadcast float4 loads.  Per (row tap, channel): 16 LDS.64 + 56 LDS.128 feed 896 FFMAs -> FMA-pipe bound.
constexpr int ST_TH = 4, ST_TW = 128, ST_PH = (ST_TH - 1) * 2 + 7, ST_PW = (ST_TW - 1) * 2 + 7, ST_LD = ST_PW + 1;

template <int MODE>
__global__ void __launch_bounds__(256, 1) stem7_kernel(const void* __restrict__ src, int N, int H, int W,
                                                       const float* __restrict__ wkn, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, float* __restrict__ out, int Ho,
                                                       int Wo, int out_cs, int out_co) {
    extern __shared__ __align__(16) float sm[];
    float* sw = sm;                       // [147][64]
    float* sp = sm + 147 * 64;            // [3][ST_PH][ST_LD]
    const int tid = threadIdx.x;
    const int n = blockIdx.z;
    const int ho0 = blockIdx.y * ST_TH, wo0 = blockIdx.x * ST_TW;
    for (int i = tid; i < 147 * 64 / 4; i += 256)
        reinterpret_cast<float4*>(sw)[i] = reinterpret_cast<const float4*>(wkn)[i];
    const int hi0 = ho0 * 2 - 3, wi0 = wo0 * 2 - 3;
    // Stage the raw input rows first (coalesced 32-bit words, 8 independent loads in flight per thread), then scatter them
    // into the three channel planes from shared memory: a per-element global load loop here was latency-bound and
    // cost ~40% of the kernel.
    uint32_t* raw = reinterpret_cast<uint32_t*>(sp + 3 * ST_PH * ST_LD);      // [ST_PH][RAW_WORDS]
    constexpr int ELT = MODE == 0 ? 1 : 4;                                     // bytes per input element
    constexpr int RAW_WORDS = (ST_PW * 3 * ELT + 3) / 4 + 2;                   // words per staged row (+ alignment slack)
    const char* base = static_cast<const char*>(src);
    const long long img_bytes = (long long)H * W * 3 * ELT, total_bytes = (long long)N * img_bytes;
    for (int i0 = tid; i0 < ST_PH * RAW_WORDS; i0 += 256 * 8) {
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * 256;
            v[u] = 0;
            if (i < ST_PH * RAW_WORDS) {
                const int py = i / RAW_WORDS, wd = i - py * RAW_WORDS;
                const int hi = hi0 + py;
                // byte address of the first element of the row segment, aligned down to 4
                const long long row0 = (long long)n * img_bytes + ((long long)hi * W + wi0) * 3 * ELT;
                const long long a = (row0 & ~3LL) + 4LL * wd;
                if (hi >= 0 && hi < H && a >= 0 && a + 4 <= total_bytes) v[u] = *reinterpret_cast<const uint32_t*>(base + a);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * 256;
            if (i < ST_PH * RAW_WORDS) raw[i] = v[u];
        }
    }
    __syncthreads();
    for (int i = tid; i < ST_PH * ST_PW * 3; i += 256) {
        const int c = i % 3;
        const int t = i / 3;
        const int px = t % ST_PW, py = t / ST_PW;
        const int hi = hi0 + py, wi = wi0 + px;
        float v = 0.f;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
            const long long row0 = (long long)n * img_bytes + ((long long)hi * W + wi0) * 3 * ELT;
            const int mis = (int)(row0 & 3);                                   // the staged row starts `mis` bytes early
            if (MODE == 0) {
                // RGB uint8 -> BGR, minus (104,117,123): retinaface.py:450-451
                const uint8_t* rb = reinterpret_cast<const uint8_t*>(raw + py * RAW_WORDS);
                const float mean = c == 0 ? 104.f : (c == 1 ? 117.f : 123.f);
                v = (float)rb[mis + px * 3 + (2 - c)] - mean;
            } else {
                v = reinterpret_cast<const float*>(raw + py * RAW_WORDS)[px * 3 + c];   // float rows are 4-byte aligned: mis == 0
            }
        }
        sp[(c * ST_PH + py) * ST_LD + px] = v;
    }
    __syncthreads();
    const int half = tid >> 7;            // output channels [32*half, 32*half+32)
    const int g = tid & 127;
    const int xi = g & 31, ty = g >> 5;   // pixels (ty, xi + 32*i), i = 0..3
    float acc[4][32];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[i][j] = 0.f;
    for (int r = 0; r < 7; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float* prow = sp + (c * ST_PH + ty * 2 + r) * ST_LD;
            float x[4][8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float2 v = *reinterpret_cast<const float2*>(prow + 2 * (xi + 32 * i) + 2 * q);
                    x[i][2 * q] = v.x;
                    x[i][2 * q + 1] = v.y;
                }
#pragma unroll
            for (int s = 0; s < 7; ++s) {
                const float* pw = sw + ((r * 7 + s) * 3 + c) * 64 + half * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(pw + j);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[i][j] = fmaf(x[i][s], w4.x, acc[i][j]);
                        acc[i][j + 1] = fmaf(x[i][s], w4.y, acc[i][j + 1]);
                        acc[i][j + 2] = fmaf(x[i][s], w4.z, acc[i][j + 2]);
                        acc[i][j + 3] = fmaf(x[i][s], w4.w, acc[i][j + 3]);
                    }
                }
            }
        }
    }
    const int ho = ho0 + ty;
    if (ho >= Ho) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int wo = wo0 + xi + 32 * i;
        if (wo >= Wo) continue;
        float* dst = out + (((size_t)n * Ho + ho) * Wo + wo) * out_cs + out_co + half * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 sc = *reinterpret_cast<const float4*>(scale + half * 32 + j);
            const float4 sh = *reinterpret_cast<const float4*>(shift + half * 32 + j);
            float4 o;
            o.x = fmaxf(acc[i][j] * sc.x + sh.x, 0.f);
            o.y = fmaxf(acc[i][j + 1] * sc.y + sh.y, 0.f);
            o.z = fmaxf(acc[i][j + 2] * sc.z + sh.z, 0.f);
            o.w = fmaxf(acc[i][j + 3] * sc.w + sh.w, 0.f);
            *reinterpret_cast<float4*>(dst + j) = o;
        }
    }
}

// ------------------------------------------------------------------------------- RRDB conv_first (3x3, Cin=3)
// rrdb.py:77 conv_first on images/255 (rrdb.py:142); input f32 NCHW, output NHWC 64 channels, bias only.
template <bool U8>
__global__ void __launch_bounds__(256) conv3_first_kernel(const void* __restrict__ src_v, float in_div, int N, int H,
                                                          int W, const float* __restrict__ wkn,
                                                          const float* __restrict__ shift, float* __restrict__ out,
                                                          int out_cs, int out_co) {
    __shared__ __align__(16) float sw[27 * 64];
    for (int i = threadIdx.x; i < 27 * 64; i += 256) sw[i] = wkn[i];
    __syncthreads();
    size_t pix = (size_t)blockIdx.x * 256 + threadIdx.x;
    size_t total = (size_t)N * H * W;
    if (pix >= total) return;
    int wq = pix % W;
    size_t t = pix / W;
    int hq = t % H;
    int n = t / H;
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = shift[j];
    for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
            int hi = hq + r - 1, wi = wq + s - 1;
            if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float raw = U8 ? (float)static_cast<const uint8_t*>(src_v)[(((size_t)n * H + hi) * W + wi) * 3 + c]
                                     : static_cast<const float*>(src_v)[(((size_t)n * 3 + c) * H + hi) * W + wi];
                float x = __fdiv_rn(raw, in_div);
                const float* pw = sw + ((r * 3 + s) * 3 + c) * 64;
#pragma unroll
                for (int j = 0; j < 64; ++j) acc[j] = fmaf(x, pw[j], acc[j]);
            }
        }
    float* dst = out + pix * out_cs + out_co;
#pragma unroll
    for (int j = 0; j < 64; j += 4)
        *reinterpret_cast<float4*>(dst + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
}

// ------------------------------------------------------------------------------- RRDB conv_last (3x3, 64 -> 3)
// rrdb.py:81 conv_last at the x4 resolution: 3 output channels, so an implicit GEMM would read every input tile for three
// columns of MMA (3.5 ms per 8.4 M pixels on the tensor-core kernel).  Direct fp32 convolution instead: a block stages the
// (8+2) x (32+2) halo of its 8 x 32 output pixels in shared memory (pixel pitch 68 floats: conflict-free float4 reads), one
// thread per pixel, the 1728 weights are FFMA constant-bank operands (no load instructions for them).
__constant__ float c_last_w[9 * 64 * 3];   // [tap][cin][cout]
__constant__ float c_last_b[4];
constexpr int CL_TW = 32, CL_TH = 8, CL_PITCH = 68;
constexpr int CL_SMEM = (CL_TH + 2) * (CL_TW + 2) * CL_PITCH * 4;

__global__ void __launch_bounds__(256) conv3_last_kernel(const float* __restrict__ in, int in_cs, int in_co, int H, int W,
                                                         float* __restrict__ out, int out_cs, int out_co) {
    extern __shared__ __align__(16) float cl_tile[];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int w0 = blockIdx.x * CL_TW, h0 = blockIdx.y * CL_TH, n = blockIdx.z;
    constexpr int TOTAL = (CL_TH + 2) * (CL_TW + 2) * 16, BATCH = 8;        // 16-byte pieces of the halo; loads in flight per thread
    for (int base = threadIdx.x; base < TOTAL; base += 256 * BATCH) {
        float4 v[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            const int i = base + k * 256, px = i >> 4, q = i & 15;
            const int hi = h0 + px / (CL_TW + 2) - 1, wi = w0 + px % (CL_TW + 2) - 1;
            v[k] = make_float4(0.f, 0.f, 0.f, 0.f);                        // zero padding
            if (i < TOTAL && hi >= 0 && hi < H && wi >= 0 && wi < W)
                v[k] = __ldg(reinterpret_cast<const float4*>(in + (((size_t)n * H + hi) * W + wi) * in_cs + in_co + q * 4));
        }
#pragma unroll
        for (int k = 0; k < BATCH; ++k) {
            const int i = base + k * 256;
            if (i < TOTAL) *reinterpret_cast<float4*>(cl_tile + (i >> 4) * CL_PITCH + (i & 15) * 4) = v[k];
        }
    }
    __syncthreads();
    const int ho = h0 + ty, wo = w0 + tx;
    if (ho >= H || wo >= W) return;
    float a0 = c_last_b[0], a1 = c_last_b[1], a2 = c_last_b[2];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const float* px = cl_tile + ((ty + t / 3) * (CL_TW + 2) + tx + t % 3) * CL_PITCH;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 x = *reinterpret_cast<const float4*>(px + q * 4);
            const float* w = c_last_w + (t * 64 + q * 4) * 3;
            a0 = fmaf(x.x, w[0], a0); a1 = fmaf(x.x, w[1], a1); a2 = fmaf(x.x, w[2], a2);
            a0 = fmaf(x.y, w[3], a0); a1 = fmaf(x.y, w[4], a1); a2 = fmaf(x.y, w[5], a2);
            a0 = fmaf(x.z, w[6], a0); a1 = fmaf(x.z, w[7], a1); a2 = fmaf(x.z, w[8], a2);
            a0 = fmaf(x.w, w[9], a0); a1 = fmaf(x.w, w[10], a1); a2 = fmaf(x.w, w[11], a2);
        }
    }
    float* dst = out + (((size_t)n * H + ho) * W + wo) * out_cs + out_co;
    dst[0] = a0; dst[1] = a1; dst[2] = a2;
}

// ------------------------------------------------------------------------------------------- maxpool 3x3/s2/p1
__global__ void maxpool3s2_kernel(const float* __restrict__ in, int N, int H, int W, int C, int in_cs, int in_co,
                                  float* __restrict__ out, int Ho, int Wo, int out_cs, int out_co) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int c4n = C / 4;
    size_t total = (size_t)N * Ho * Wo * c4n;
    if (idx >= total) return;
    int c4 = idx % c4n;
    size_t p = idx / c4n;
    int wo = p % Wo;
    size_t t = p / Wo;
    int ho = t % Ho;
    int n = t / Ho;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int r = 0; r < 3; ++r) {
        int hi = ho * 2 - 1 + r;
        if (hi < 0 || hi >= H) continue;
        for (int s = 0; s < 3; ++s) {
            int wi = wo * 2 - 1 + s;
            if (wi < 0 || wi >= W) continue;
            float4 v = *reinterpret_cast<const float4*>(in + (((size_t)n * H + hi) * W + wi) * in_cs + in_co + c4 * 4);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
    }
    *reinterpret_cast<float4*>(out + p * out_cs + out_co + c4 * 4) = m;
}

// ------------------------------------------------------------------------------------- global average pooling
// F.avg_pool2d(x, x.size()[2:]) (_layers.py:307,332,360).  grid (C/32, N), block 32x8.
__global__ void global_avgpool_kernel(const float* __restrict__ in, int HW, int cs, int co, int C,
                                      float* __restrict__ out) {
    __shared__ float red[8][33];
    int c = blockIdx.x * 32 + threadIdx.x;
    int n = blockIdx.y;
    float s = 0.f;
    const float* base = in + (size_t)n * HW * cs + co + c;
    for (int p = threadIdx.y; p < HW; p += 8) s += base[(size_t)p * cs];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0) {
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x];
        out[(size_t)n * C + c] = tot / (float)HW;
    }
}

// ------------------------------------------------------------- 1x1 conv on pooled vectors (+BN fold, activation)
__global__ void fc_kernel(const float* __restrict__ in, int cin, const float* __restrict__ wkn, int cout_pad,
                          int cout, const float* __restrict__ scale, const float* __restrict__ shift, int act,
                          float* __restrict__ out) {
    int co = blockIdx.x * blockDim.x + threadIdx.x;
    int n = blockIdx.y;
    if (co >= cout) return;
    float s = 0.f;
    const float* x = in + (size_t)n * cin;
    for (int ci = 0; ci < cin; ++ci) s = fmaf(x[ci], wkn[(size_t)ci * cout_pad + co], s);
    out[(size_t)n * cout + co] = act_fn(s * scale[co] + shift[co], act, 0.f);
}

// --------------------------------------------------- out = in * mul[n][c] (+ addvec[n][c]) (+ in), NHWC float4
__global__ void channel_affine_kernel(const float* __restrict__ in, int in_cs, int in_co, size_t HW, int C,
                                      const float* __restrict__ mul, const float* __restrict__ addv, int add_self,
                                      float* __restrict__ out, int out_cs, int out_co, size_t total) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int c4n = C / 4;
    int c = (idx % c4n) * 4;
    size_t p = idx / c4n;
    int n = p / HW;
    float4 v = *reinterpret_cast<const float4*>(in + p * in_cs + in_co + c);
    float4 m = *reinterpret_cast<const float4*>(mul + (size_t)n * C + c);
    float4 o = make_float4(v.x * m.x, v.y * m.y, v.z * m.z, v.w * m.w);
    if (addv) {
        float4 a = *reinterpret_cast<const float4*>(addv + (size_t)n * C + c);
        o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    }
    if (add_self) { o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w; }
    *reinterpret_cast<float4*>(out + p * out_cs + out_co + c) = o;
}

// ------------------------------------------------------------------------ nearest x2 upsample (F.interpolate, scale 2)
// Materialises the upsampled map so the following 3x3 conv can run on the tensor-core kernel (TMA boxes need a dense
// input): _layers.py:334-344 (BiSeNet context path), rrdb.py:78-79 (RRDBNet upconv1/2).  HBM-bound float4 copy.
__global__ void upsample2x_kernel(const float* __restrict__ in, int H, int W, int C, int in_cs, int in_co,
                                  float* __restrict__ out, int out_cs, int out_co, size_t total) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c4n = C / 4;
    const int c = (idx % c4n) * 4;
    size_t p = idx / c4n;
    const int wo = p % (2 * W);
    size_t t = p / (2 * W);
    const int ho = t % (2 * H);
    const size_t n = t / (2 * H);
    const float4 v = *reinterpret_cast<const float4*>(in + ((n * H + (ho >> 1)) * W + (wo >> 1)) * in_cs + in_co + c);
    *reinterpret_cast<float4*>(out + p * out_cs + out_co + c) = v;
}

// ----------------------------------------------------------------------------------------- layout conversions
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, int HW, int C, int cs, int co,
                                    float* __restrict__ out, size_t total) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // over n*c*hw, hw fastest (coalesced writes)
    if (idx >= total) return;
    int p = idx % HW;
    size_t t = idx / HW;
    int c = t % C;
    size_t n = t / C;
    out[idx] = in[(n * HW + p) * cs + co + c];
}

}  // namespace

int launch_stem7(fcp_ctx* ctx, const void* src, int mode, int n, int h, int w, const float* w_kn, const float* scale,
                 const float* shift, Tensor out) {
    // weights + 3 channel planes + raw row staging (mode 1 rows are floats: ST_PW*3 words)
    size_t smem = (147 * 64 + 3 * ST_PH * ST_LD + ST_PH * (ST_PW * 3 + 3)) * sizeof(float);
    static uint64_t configured = 0;                    // one bit per device: the attribute is per device
    if (!((configured >> (ctx->device & 63)) & 1)) {
        FCP_CUDA(ctx, cudaFuncSetAttribute(stem7_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FCP_CUDA(ctx, cudaFuncSetAttribute(stem7_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured |= (uint64_t)1 << (ctx->device & 63);
    }
    dim3 grid((out.w + ST_TW - 1) / ST_TW, (out.h + ST_TH - 1) / ST_TH, n);
    if (mode == 0)
        stem7_kernel<0><<<grid, 256, smem, ctx->stream>>>(src, n, h, w, w_kn, scale, shift, out.p, out.h, out.w, out.cs, out.co);
    else
        stem7_kernel<1><<<grid, 256, smem, ctx->stream>>>(src, n, h, w, w_kn, scale, shift, out.p, out.h, out.w, out.cs, out.co);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

// rows[n][hi][wo][32]: channel s*3 + c = input(hi, 2*wo + s - 3, conv channel c), 0 outside the image / for channels >= 21.
// One thread per (hi, wo): 128 contiguous output bytes; neighbouring threads re-read overlapping inputs through L1.
template <int MODE>
__global__ void stem_rows_kernel(const void* __restrict__ src, int N, int H, int W, int Wo, float* __restrict__ rows, int cs) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)N * H * Wo;
    if (idx >= total) return;
    const int wo = (int)(idx % Wo);
    const size_t t = idx / Wo;                          // n * H + hi
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
    for (int sx = 0; sx < 7; ++sx) {
        const int wi = 2 * wo + sx - 3;
        if (wi < 0 || wi >= W) continue;
        if (MODE == 0) {                                // RGB uint8 -> BGR, minus (104,117,123): retinaface.py:450-451
            const uint8_t* px = static_cast<const uint8_t*>(src) + (t * W + wi) * 3;
            v[sx * 3 + 0] = (float)px[2] - 104.f;
            v[sx * 3 + 1] = (float)px[1] - 117.f;
            v[sx * 3 + 2] = (float)px[0] - 123.f;
        } else {
            const float* px = static_cast<const float*>(src) + (t * W + wi) * 3;
            v[sx * 3 + 0] = px[0]; v[sx * 3 + 1] = px[1]; v[sx * 3 + 2] = px[2];
        }
    }
    float4* dst = reinterpret_cast<float4*>(rows + idx * cs);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

int launch_stem_rows(fcp_ctx* ctx, const void* src, int mode, int n, int h, int w, Tensor rows) {
    if (rows.c != 32 || rows.co != 0 || rows.cs % 4 != 0 || rows.h != h) return fail(ctx, FCP_ERR_INVALID, "stem rows: bad tensor");
    const size_t total = (size_t)n * h * rows.w;
    const unsigned blocks = (unsigned)((total + 255) / 256);
    if (mode == 0) stem_rows_kernel<0><<<blocks, 256, 0, ctx->stream>>>(src, n, h, w, rows.w, rows.p, rows.cs);
    else stem_rows_kernel<1><<<blocks, 256, 0, ctx->stream>>>(src, n, h, w, rows.w, rows.p, rows.cs);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_conv3_first(fcp_ctx* ctx, const float* src_nchw, float in_div, int n, int h, int w, const float* w_kn,
                       const float* shift, Tensor out) {
    size_t total = (size_t)n * h * w;
    conv3_first_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(src_nchw, in_div, n, h, w, w_kn, shift,
                                                                                        out.p, out.cs, out.co);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_conv3_first_u8(fcp_ctx* ctx, const uint8_t* src_nhwc, int n, int h, int w, const float* w_kn, const float* shift, Tensor out) {
    size_t total = (size_t)n * h * w;
    conv3_first_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(src_nhwc, 255.f, n, h, w, w_kn, shift,
                                                                                       out.p, out.cs, out.co);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

// `packed` = host [9*64*3 weights (tap, cin, cout)] + [3 bias, 1 pad] (finalize_rrdbnet); re-uploaded per launch, stream-ordered,
// because the constant bank is per device and several contexts may hold different weights
int launch_conv3_last(fcp_ctx* ctx, Tensor in, const float* packed, Tensor out) {
    if (in.c != 64 || out.c != 3 || ((in.cs | in.co) & 3)) return fail(ctx, FCP_ERR_INVALID, "conv3_last: expects a 64 -> 3 convolution");
    static uint64_t configured = 0;                    // one bit per device: the attribute is per device
    if (!((configured >> (ctx->device & 63)) & 1)) {
        FCP_CUDA(ctx, cudaFuncSetAttribute(conv3_last_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CL_SMEM));
        configured |= (uint64_t)1 << (ctx->device & 63);
    }
    FCP_CUDA(ctx, cudaMemcpyToSymbolAsync(c_last_w, packed, sizeof(float) * 9 * 64 * 3, 0, cudaMemcpyHostToDevice, ctx->stream));
    FCP_CUDA(ctx, cudaMemcpyToSymbolAsync(c_last_b, packed + 9 * 64 * 3, sizeof(float) * 4, 0, cudaMemcpyHostToDevice, ctx->stream));
    dim3 grid((in.w + CL_TW - 1) / CL_TW, (in.h + CL_TH - 1) / CL_TH, in.n);
    conv3_last_kernel<<<grid, 256, CL_SMEM, ctx->stream>>>(in.p, in.cs, in.co, in.h, in.w, out.p, out.cs, out.co);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_maxpool3s2(fcp_ctx* ctx, Tensor in, Tensor out) {
    size_t total = out.pixels() * (in.c / 4);
    maxpool3s2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(in.p, in.n, in.h, in.w, in.c, in.cs, in.co,
                                                                               out.p, out.h, out.w, out.cs, out.co);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_global_avgpool(fcp_ctx* ctx, Tensor in, float* out_nc) {
    if (in.c % 32) return fail(ctx, FCP_ERR_INVALID, "avgpool: C must be a multiple of 32");
    dim3 grid(in.c / 32, in.n), block(32, 8);
    global_avgpool_kernel<<<grid, block, 0, ctx->stream>>>(in.p, in.h * in.w, in.cs, in.co, in.c, out_nc);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_fc(fcp_ctx* ctx, const float* in_nc, int n, int cin, const ConvWeights* wt, int act, float* out_nc) {
    dim3 grid((wt->cout + 127) / 128, n);
    fc_kernel<<<grid, 128, 0, ctx->stream>>>(in_nc, cin, wt->w_kn, wt->cout_pad, wt->cout, wt->scale, wt->shift, act, out_nc);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_channel_affine(fcp_ctx* ctx, Tensor in, const float* mul_nc, const float* addvec_nc, int add_self,
                          Tensor out) {
    size_t total = in.pixels() * (in.c / 4);
    channel_affine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
        in.p, in.cs, in.co, (size_t)in.h * in.w, in.c, mul_nc, addvec_nc, add_self, out.p, out.cs, out.co, total);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_upsample2x(fcp_ctx* ctx, Tensor in, Tensor out) {
    size_t total = out.pixels() * (in.c / 4);
    upsample2x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(in.p, in.h, in.w, in.c, in.cs, in.co, out.p,
                                                                               out.cs, out.co, total);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

int launch_nhwc_to_nchw(fcp_ctx* ctx, const float* in, int n, int h, int w, int c, int cs, float* out) {
    size_t total = (size_t)n * c * h * w;
    nhwc_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(in, h * w, c, cs, 0, out, total);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

}  // namespace fcp
