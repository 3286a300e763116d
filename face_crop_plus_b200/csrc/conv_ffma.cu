// conv_ffma.cu — exact-fp32 implicit-GEMM convolution on the CUDA cores (NHWC activations, [K][Cout] weights).
//
// Replaces nn.Conv2d (+ BatchNorm2d eval + activation + residual adds) of the reference graphs
// (_layers.py:70-125,179-239,279-283; torchvision resnet.py:143-163).  Never materialises im2col: each K-step
// stages one (tap, 32-channel) slab of the shifted input window with cp.async (zero-filled outside the image).
//
// Tile: 128 output pixels x BN output channels x 32 input channels per step, 256 threads, 8 x (BN/16) register
// micro-tile, 3-stage cp.async pipeline.  Bound: fp32 FMA pipe (this is the strict-precision kernel; the
// tensor-core kernel lives in conv_tc.cu).
#include "common.h"

namespace fcp {

namespace {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int A_LD = BK + 4;   // padded row (floats): keeps 16-byte alignment, kills bank conflicts on the stores
constexpr int STAGES = 3;
constexpr int THREADS = 256;

struct ConvArgs {
    const float* in; int N, H, W, in_cs, in_co, Cin, up_in, Hp, Wp;   // H,W logical; Hp,Wp physical
    const float* w; int cout_pad;
    float* out; int Ho, Wo, out_cs, out_co, Cout;
    int KH, KW, stride, pad;
    const float* scale; const float* shift;
    const float* res1; int res1_cs, res1_co;
    int act; float slope;
    float post_scale; const float* res2; int res2_cs, res2_co, res2_h, res2_w;
    float post_scale2; const float* res3; int res3_cs, res3_co;
    int M;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == FCP_ACT_RELU) return fmaxf(v, 0.f);
    if (act == FCP_ACT_LRELU) return v > 0.f ? v : v * slope;
    if (act == FCP_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}

template <int BN>
__global__ void __launch_bounds__(THREADS, (BN == 128) ? 2 : 2)
conv_ffma_kernel(const ConvArgs a) {
    constexpr int TN = BN / 16;                 // channels per thread: 2, 4 or 8
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                           // [STAGES][BM][A_LD]
    float* Bs = smem + STAGES * BM * A_LD;      // [STAGES][BK][BN]

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // ---- per-thread A-load assignment: 4 (pixel, channel-quad) pairs, fixed over the K loop
    const float* a_base[4];
    int a_hi0[4], a_wi0[4];
    bool a_ok[4];
    const int q = tid & 7;                      // which 4-channel group of the 32-channel slab
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        int pix = (tid >> 3) + t * 32;
        int m = m0 + pix;
        a_ok[t] = m < a.M;
        int mm = a_ok[t] ? m : 0;
        int wo = mm % a.Wo;
        int tmp = mm / a.Wo;
        int ho = tmp % a.Ho;
        int n = tmp / a.Ho;
        a_hi0[t] = ho * a.stride - a.pad;
        a_wi0[t] = wo * a.stride - a.pad;
        a_base[t] = a.in + (size_t)n * a.Hp * a.Wp * a.in_cs + a.in_co + q * 4;
    }
    const int cchunks = a.Cin / BK;
    const int ksteps = a.KH * a.KW * cchunks;

    auto load_stage = [&](int stage, int ks) {
        int tap = ks / cchunks;
        int c0 = (ks - tap * cchunks) * BK;
        int r = tap / a.KW, s = tap - r * a.KW;
        float* As_s = As + stage * BM * A_LD;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            int pix = (tid >> 3) + t * 32;
            int hi = a_hi0[t] + r, wi = a_wi0[t] + s;
            bool ok = a_ok[t] && hi >= 0 && hi < a.H && wi >= 0 && wi < a.W;
            int hp = a.up_in ? (hi >> 1) : hi, wp = a.up_in ? (wi >> 1) : wi;
            const float* src = ok ? a_base[t] + ((size_t)hp * a.Wp + wp) * a.in_cs + c0 : a.in;
            cp_async16(As_s + pix * A_LD + q * 4, src, ok);
        }
        float* Bs_s = Bs + stage * BK * BN;
        const float* wsrc = a.w + (size_t)(tap * a.Cin + c0) * a.cout_pad + n0;
        constexpr int B_VEC = BK * BN / 4;      // float4 count
#pragma unroll
        for (int i = tid; i < B_VEC; i += THREADS) {
            int row = i / (BN / 4), col = i - row * (BN / 4);
            cp_async16(Bs_s + row * BN + col * 4, wsrc + (size_t)row * a.cout_pad + col * 4, true);
        }
    };

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < ksteps) load_stage(s, s);
        cp_async_commit();
    }

    for (int ks = 0; ks < ksteps; ++ks) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = ks + STAGES - 1;
            if (nk < ksteps) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const float* As_s = As + (ks % STAGES) * BM * A_LD;
        const float* Bs_s = Bs + (ks % STAGES) * BK * BN;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            float4 av[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                av[i] = *reinterpret_cast<const float4*>(As_s + (ty * 4 + i) * A_LD + kk);
                av[i + 4] = *reinterpret_cast<const float4*>(As_s + (64 + ty * 4 + i) * A_LD + kk);
            }
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                float bv[TN];
                const float* brow = Bs_s + (kk + k4) * BN;
                if constexpr (TN == 8) {
                    float4 b0 = *reinterpret_cast<const float4*>(brow + tx * 4);
                    float4 b1 = *reinterpret_cast<const float4*>(brow + 64 + tx * 4);
                    bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
                    bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
                } else if constexpr (TN == 4) {
                    float4 b0 = *reinterpret_cast<const float4*>(brow + tx * 4);
                    bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
                } else {
                    float2 b0 = *reinterpret_cast<const float2*>(brow + tx * 2);
                    bv[0] = b0.x; bv[1] = b0.y;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float aval = k4 == 0 ? av[i].x : k4 == 1 ? av[i].y : k4 == 2 ? av[i].z : av[i].w;
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(aval, bv[j], acc[i][j]);
                }
            }
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: scale/shift (+res1) -> act -> *post_scale (+res2) -> *post_scale2 (+res3)
    constexpr int NG = (TN == 8) ? 2 : 1;       // column groups per thread
    constexpr int GW = (TN == 2) ? 2 : 4;       // group width
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= a.M) continue;
        size_t res2_pix = 0;
        if (a.res2) {
            if (a.res2_h) {
                // F.interpolate(mode="nearest"): src = min(floor(dst * (float)in/out), in-1)   (_layers.py:137-142)
                int wo = m % a.Wo;
                int tmp = m / a.Wo;
                int ho = tmp % a.Ho;
                int n = tmp / a.Ho;
                int hs = min((int)floorf(ho * ((float)a.res2_h / (float)a.Ho)), a.res2_h - 1);
                int ws = min((int)floorf(wo * ((float)a.res2_w / (float)a.Wo)), a.res2_w - 1);
                res2_pix = ((size_t)n * a.res2_h + hs) * a.res2_w + ws;
            } else {
                res2_pix = m;
            }
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            int nb = n0 + (TN == 2 ? tx * 2 : tx * 4) + g * 64;
#pragma unroll
            for (int j = 0; j < GW; ++j) {
                int n = nb + j;
                if (n >= a.Cout) continue;
                float v = acc[i][g * 4 + j] * a.scale[n] + a.shift[n];
                if (a.res1) v += a.res1[(size_t)m * a.res1_cs + a.res1_co + n];
                v = apply_act(v, a.act, a.slope);
                if (a.post_scale != 1.f) v *= a.post_scale;
                if (a.res2) v += a.res2[res2_pix * a.res2_cs + a.res2_co + n];
                if (a.post_scale2 != 1.f) v *= a.post_scale2;
                if (a.res3) v += a.res3[(size_t)m * a.res3_cs + a.res3_co + n];
                acc[i][g * 4 + j] = v;
            }
            float* dst = a.out + (size_t)m * a.out_cs + a.out_co + nb;
            if (GW == 4 && nb + 3 < a.Cout) {
                *reinterpret_cast<float4*>(dst) =
                    make_float4(acc[i][g * 4], acc[i][g * 4 + 1], acc[i][g * 4 + 2], acc[i][g * 4 + 3]);
            } else if (GW == 2 && nb + 1 < a.Cout) {
                *reinterpret_cast<float2*>(dst) = make_float2(acc[i][0], acc[i][1]);
            } else {
                for (int j = 0; j < GW; ++j)
                    if (nb + j < a.Cout) dst[j] = acc[i][g * 4 + j];
            }
        }
    }
}

template <int BN>
int launch(fcp_ctx* ctx, const ConvArgs& a) {
    size_t smem = (size_t)STAGES * (BM * A_LD + BK * BN) * sizeof(float);
    static uint64_t configured = 0;                    // one bit per device: the attribute is per device
    if (!((configured >> (ctx->device & 63)) & 1)) {
        FCP_CUDA(ctx, cudaFuncSetAttribute(conv_ffma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured |= (uint64_t)1 << (ctx->device & 63);
    }
    dim3 grid((a.M + BM - 1) / BM, (a.cout_pad + BN - 1) / BN);
    conv_ffma_kernel<BN><<<grid, THREADS, smem, ctx->stream>>>(a);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

}  // namespace

int launch_conv_ffma(fcp_ctx* ctx, const ConvOp& op) {
    const ConvWeights& wt = *op.wt;
    if (op.in.c != wt.cin || op.out.c != wt.cout) return fail(ctx, FCP_ERR_INVALID, "conv: channel mismatch");
    if (wt.cin % BK != 0) return fail(ctx, FCP_ERR_INVALID, "conv_ffma: Cin must be a multiple of 32");
    if ((op.in.cs | op.in.co | op.out.cs | op.out.co) & 3) return fail(ctx, FCP_ERR_INVALID, "conv: channel strides/offsets must be multiples of 4");
    ConvArgs a{};
    a.in = op.in.p; a.N = op.in.n; a.Hp = op.in.h; a.Wp = op.in.w; a.up_in = op.up_in;
    a.H = op.up_in ? op.in.h * 2 : op.in.h; a.W = op.up_in ? op.in.w * 2 : op.in.w;
    a.in_cs = op.in.cs; a.in_co = op.in.co; a.Cin = wt.cin;
    a.w = wt.w_kn; a.cout_pad = wt.cout_pad;
    a.out = op.out.p; a.Ho = op.out.h; a.Wo = op.out.w; a.out_cs = op.out.cs; a.out_co = op.out.co; a.Cout = wt.cout;
    a.KH = a.KW = wt.k; a.stride = op.stride; a.pad = op.pad;
    int eh = (a.H + 2 * op.pad - wt.k) / op.stride + 1, ew = (a.W + 2 * op.pad - wt.k) / op.stride + 1;
    if (eh != a.Ho || ew != a.Wo || op.out.n != op.in.n) return fail(ctx, FCP_ERR_INVALID, "conv: output shape mismatch");
    a.scale = wt.scale; a.shift = wt.shift;
    a.res1 = op.res1; a.res1_cs = op.res1_cs; a.res1_co = op.res1_co;
    a.act = op.act; a.slope = op.slope;
    a.post_scale = op.post_scale; a.res2 = op.res2; a.res2_cs = op.res2_cs; a.res2_co = op.res2_co; a.res2_h = op.res2_h; a.res2_w = op.res2_w;
    a.post_scale2 = op.post_scale2; a.res3 = op.res3; a.res3_cs = op.res3_cs; a.res3_co = op.res3_co;
    a.M = op.out.n * op.out.h * op.out.w;
    if (wt.cout_pad % 128 == 0) return launch<128>(ctx, a);
    if (wt.cout_pad % 64 == 0) return launch<64>(ctx, a);
    return launch<32>(ctx, a);
}

}  // namespace fcp
