// conv_tc.cu — implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM), operands
// staged by TMA, fp32-equivalent through a 3xTF32 split.  sm_100a only.
//
// Replaces the same nn.Conv2d(+BN+act+residual) sites as conv_ffma.cu; same ConvOp contract and epilogue.
//
// GEMM view: D[M = N*Ho*Wo pixels, Cout] = A[M, K = KH*KW*Cin] * W[Cout, K]^T, NHWC activations (K-major rows of 32
// channels = 128 B), weights pre-packed K-major.  No im2col buffer exists anywhere: for every (tap, 32-channel) K-block
// the TMA engine loads the *shifted* BHxBW spatial box of the input straight into the 128B-swizzled K-major tile the
// MMA descriptor expects; out-of-image rows/columns are zero-filled by TMA (== the conv's zero padding).  Stride-2
// convolutions address one of four parity views of the input (x = 2i+p), so their boxes are dense as well.
//
// Precision (parity bar: landmarks <= 1e-3 px, bit-exact argmax => single-pass TF32/BF16 is not enough):
//   a = a_hi + a_lo, w = w_hi + w_lo with *_hi = value truncated to TF32;  D += a_hi*w_hi + a_lo*w_hi + a_hi*w_lo.
//   w_hi/w_lo are split on the host and stay in shared memory (B operand, smem descriptors); a_hi/a_lo are split on the
//   fly by 4 converter warps that read the TMA-landed fp32 tile (one 128-byte pixel row per thread, conflict-free thanks
//   to the swizzle) and write both parts into TENSOR MEMORY (tcgen05.st), from where the MMA reads its A operand
//   (tcgen05.mma with A in TMEM).  That halves the shared-memory traffic of the MMAs, removes the converters' smem
//   stores, and shrinks a pipeline stage to 48 KiB so that 4 stages fit (the K loop is latency-bound, not bandwidth-
//   bound: measured cycle per stage ~ T_tma + T_convert + T_mma, throughput = stages / cycle).  Every operand has its
//   low 13 mantissa bits zeroed, hence the result does not depend on how the tensor core rounds fp32 -> tf32.
//
//   Measured on B200: the TMEM accumulator add rounds toward zero (~3e-8 relative loss per accumulate, 5e-5 over a
//   K=4608 chain) - a systematic shrink that a 60-layer network amplifies.  Therefore a TMEM accumulation chain never
//   spans more than ONE K-block: the 8 small-term MMAs (a_lo*w_hi, a_hi*w_lo) go first, the 4 a_hi*w_hi MMAs last, then
//   the block's partial sum is drained to fp32 registers and added there with round-to-nearest.
//
// CTA = 15 warps, persistent over output tiles (128 pixels x BN channels):
//   warp 0      TMA producer          full[s]  <- TMA bytes          (waits empty[s])
//   warps 2-5   hi/lo converters      conv[s]  <- 4 arrivals (1/warp) (wait full[s])
//   warps 1,14  MMA issuers (1 lane each, alternating K-blocks)  empty[s], d_full[b] <- tcgen05.commit (wait conv[s], d_empty[b])
//   warps 6-13  accumulate+epilogue   d_empty[b] <- 8 arrivals (1/warp) (wait d_full[b]); per K-block TMEM -> regs (+=),
//               after the last K-block: fused epilogue -> HBM.  Warp w owns TMEM lanes 32*(w%4).. and half of the columns.
// Two TMEM partial-sum buffers let the drain of K-block i overlap the MMAs of K-block i+1.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.h"

namespace fcp {

namespace {

constexpr int TILE_M = 128;
constexpr int KB = 32;                         // channels per K-block (128 bytes of fp32)
constexpr int A_TILE_BYTES = TILE_M * KB * 4;  // 16 KiB
// MMA-issuing warps (warp 1 and warp 14).  Must stay 2: each partial-sum buffer (g & 1) then has exactly one issuer, so a
// parity wait on d_empty can never be more than one phase ahead (with 3 issuers the same buffer is touched out of
// order and mbarrier parity waits alias -> corrupted sums and a deadlock; measured the hard way).
constexpr int NUM_ISSUERS = 2;
constexpr int NUM_THREADS = 32 * (14 + NUM_ISSUERS - 1);

struct alignas(64) TcParams {
    CUtensorMap tmA[4];                        // input, one per (row parity, column parity); stride 1 uses [0]
    CUtensorMap tmBhi, tmBlo;                  // weights [cout_pad][K], K-major
    int N, Ho, Wo, Cout, Cin, KH, KW, stride, pad;
    int tiles_x, tiles_y, tiles_n, num_tiles, bw_log2, BH;
    float* out; int out_cs, out_co;
    const float* scale; const float* shift;
    const float* res1; int res1_cs, res1_co;
    int act; float slope;
    float post_scale; const float* res2; int res2_cs, res2_co, res2_h, res2_w;
    float post_scale2; const float* res3; int res3_cs, res3_co;
    int exp_nolo;                              // experiment: skip the w_lo loads (wrong results; L2-traffic probe)
    long long* dbg;                            // FCP_EXP_TIMELINE: clock64 stamps of CTA 0, [g][16]
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <bool BACKOFF = false>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    const uint32_t addr = smem_u32(bar);
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        // non-critical roles back off: 13 warps spinning on try_wait starve the MMA-issuing threads of issue slots
        if (BACKOFF && !done) __nanosleep(32);
    } while (!done);
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
// explicit shared-space float4 accesses: through a generic pointer the compiler emits LD.E/ST.E (generic path) for
// the epilogue slab, which measured ~4x slower than LDS/STS
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f4(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f1(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
#ifdef FCP_EXP_TIMELINE
#define TL(g, ev) do { if (p.dbg && blockIdx.x == 0 && (g) < 512) p.dbg[(g) * 16 + (ev)] = clock64(); } while (0)
#else
#define TL(g, ev) do { } while (0)
#endif
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, one CTA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// same with the A operand in tensor memory (rows = lanes, one 32-bit column per tf32 element)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 columns: thread i of the warp writes its 32 registers to lane (base_lane + i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
          "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
          "r"(r[30]), "r"(r[31])
        : "memory");
}
// arrives on `bar` when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x N consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i)
template <int N> __device__ __forceinline__ void tmem_ld(uint32_t taddr, float* v);
template <> __device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row atoms 1024 bytes apart (SBO), LBO unused (=1),
// descriptor version 1 (Blackwell), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}

__device__ __forceinline__ float act_fn(float v, int act, float slope) {
    if (act == FCP_ACT_RELU) return fmaxf(v, 0.f);
    if (act == FCP_ACT_LRELU) return v > 0.f ? v : v * slope;
    if (act == FCP_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
    return v;
}

// One shared-memory slab per epilogue warp, S[32 pixels][BN/2 channels (+4 pad)], serves three purposes in turn: the
// tile's residual operand (bottleneck shortcut / FPN top-down / RRDB skip) is prefetched into it with cp.async while the
// K loop runs; the fused epilogue math then runs in place, pixel-per-thread (the TMEM lane layout); finally the slab is
// read back transposed (channel-contiguous float4s) so that the output stores are full 128-byte lines.
template <int BN> struct Cfg {
    static constexpr int B_TILE_BYTES = BN * KB * 4;
    static constexpr int STAGE_BYTES = A_TILE_BYTES + 2 * B_TILE_BYTES;   // fp32 A tile + w_hi + w_lo
    static constexpr int STAGES = BN >= 128 ? 3 : (BN == 64 ? 5 : 7);
    // tensor memory: [0, 2*BN) two partial-sum buffers; then STAGES x (a_hi 32 cols | a_lo 32 cols)
    static constexpr int TMEM_A0 = 2 * BN;
    static constexpr int TMEM_COLS = 512;
    static_assert(TMEM_A0 + STAGES * 64 <= TMEM_COLS, "tensor memory budget");
    static constexpr int HALF = BN / 2;                                   // channels owned by one epilogue warp
    static constexpr int EPI_CW = HALF >= 32 ? 32 : 16;                   // channels per store chunk (32 -> full 128-byte lines)
    static constexpr int S_LD = HALF + 4;                                 // floats per slab row (+4 pad: conflict-free float4s)
    static constexpr int S_BYTES = 8 * 32 * S_LD * 4;                     // 8 epilogue warps x 32 pixels
    static constexpr int PAR_BYTES = 8 * 2 * HALF * 4;                    // per warp: scale[HALF] | shift[HALF]
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + S_BYTES + PAR_BYTES;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
    uint64_t* full = bars;                         // [STAGES]
    uint64_t* conv = bars + C::STAGES;             // [STAGES]
    uint64_t* empty = bars + 2 * C::STAGES;        // [STAGES]
    uint64_t* d_full = bars + 3 * C::STAGES;       // [2]  partial sum of one K-block is complete in TMEM buffer b
    uint64_t* d_empty = d_full + 2;                // [2]  buffer b has been drained to registers
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(d_empty + 2);
    float* slab_all = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 256);   // [8 warps][32 rows][S_LD]
    float* par_all = slab_all + C::S_BYTES / 4;                                            // [8 warps][2][HALF]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto stage_a = [&](int s) { return smem + s * C::STAGE_BYTES; };
    auto stage_b_hi = [&](int s) { return smem + s * C::STAGE_BYTES + A_TILE_BYTES; };
    auto stage_b_lo = [&](int s) { return smem + s * C::STAGE_BYTES + A_TILE_BYTES + C::B_TILE_BYTES; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], 4); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&d_full[a], 1); mbar_init(&d_empty[a], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(tmem_base_smem, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    const int cchunks = p.Cin / KB;
    const int kblocks = p.KH * p.KW * cchunks;
    const int BW = 1 << p.bw_log2;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        // ================================================================================== TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0; uint32_t gp = 0; (void)gp;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.tiles_n, m_tile = tile / p.tiles_n;
                const int img = m_tile / tiles_per_img, rem = m_tile - img * tiles_per_img;
                const int ho0 = (rem / p.tiles_x) * p.BH, wo0 = (rem % p.tiles_x) * BW;
                int tap = 0, cc = 0, r = 0, sx = 0;                          // K-block = (tap (r, sx), 32-channel chunk cc)
                for (int kb = 0; kb < kblocks; ++kb) {
                    int dy = r - p.pad, dx = sx - p.pad, map = 0;
                    if (p.stride == 2) {   // input row 2*ho + dy lives in parity view (dy & 1) at row ho + (dy - (dy & 1)) / 2
                        const int py = dy & 1, px = dx & 1;
                        map = py * 2 + px;
                        dy = (dy - py) >> 1;
                        dx = (dx - px) >> 1;
                    }
                    const int c0 = cc * KB, kcol = tap * p.Cin + c0;
                    mbar_wait<true>(&empty[stage], phase ^ 1);
                    TL(gp, 0);
                    mbar_expect_tx(&full[stage], A_TILE_BYTES + ((p.exp_nolo & 1) ? 1 : 2) * C::B_TILE_BYTES);
                    tma_load_4d(stage_a(stage), &p.tmA[map], &full[stage], c0, wo0 + dx, ho0 + dy, img);
                    tma_load_2d(stage_b_hi(stage), &p.tmBhi, &full[stage], kcol, n_tile * BN);
                    if (!(p.exp_nolo & 1)) tma_load_2d(stage_b_lo(stage), &p.tmBlo, &full[stage], kcol, n_tile * BN);
                    if (++cc == cchunks) { cc = 0; ++tap; if (++sx == p.KW) { sx = 0; ++r; } }
                    TL(gp, 1); ++gp;
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 || warp >= 14) {
        // ============================================================== MMA issuers (warp 1: even K-blocks, warp 14: odd)
        if (lane == 0) {
            // instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), both K-major, N>>3 at bit 17, M>>4 at bit 24
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
            // Issuing a tcgen05.mma blocks the thread for about its execution time, and every barrier wait costs a few
            // hundred cycles; with a single issuer the tensor pipe idled ~40% of each K-block.  NUM_ISSUERS threads take
            // K-blocks round-robin, so one thread's waits overlap the others' MMAs.
            const uint32_t me = warp == 1 ? 0u : (uint32_t)(warp - 13);
            uint32_t g = 0;                                                   // K-blocks so far (all tiles)
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < kblocks; ++kb, ++g) {
                    if (g % NUM_ISSUERS == me) {
                        const uint32_t buf = g & 1;
                        if (p.exp_nolo & 16) mbar_wait<true>(&d_empty[buf], ((g >> 1) & 1) ^ 1);
                        else mbar_wait(&d_empty[buf], ((g >> 1) & 1) ^ 1);    // partial-sum buffer drained
                        TL(g, 2);
                        if (p.exp_nolo & 16) mbar_wait<true>(&conv[stage], phase);
                        else mbar_wait(&conv[stage], phase);                  // operands (hi/lo) ready
                        TL(g, 3);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + buf * BN;
                        const uint32_t a_hi = tmem_base + C::TMEM_A0 + stage * 64, a_lo = a_hi + 32;   // A operand: tensor memory
                        const uint64_t b_hi = umma_desc(smem_u32(stage_b_hi(stage))), b_lo = umma_desc(smem_u32(stage_b_lo(stage)));
                        // one K-step = 8 tf32: +8 TMEM columns for A, +32 bytes (+2 in the addr>>4 field) inside B's swizzle span.
                        // Small terms first: while the accumulator is tiny its round-toward-zero losses are negligible.
#pragma unroll
                        for (int k = 0; k < KB / 8; ++k) {
                            umma_tf32_ts(d_tmem, a_lo + 8 * k, b_hi + 2 * k, idesc, k != 0);
                            umma_tf32_ts(d_tmem, a_hi + 8 * k, b_lo + 2 * k, idesc, 1);
                        }
#pragma unroll
                        for (int k = 0; k < KB / 8; ++k) umma_tf32_ts(d_tmem, a_hi + 8 * k, b_hi + 2 * k, idesc, 1);
                        umma_commit(&empty[stage]);                           // smem slot + TMEM A slot reusable once these MMAs retire
                        umma_commit(&d_full[buf]);                            // partial sum of this K-block complete
                        TL(g, 4);
                    }
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp < 6) {
        // ================================================================= hi/lo converters (warps 2..5, 128 threads)
        // thread <-> one pixel row of the tile == one TMEM lane (a warp may only touch lanes 32*(warp%4)..+31)
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
        int stage = 0; uint32_t phase = 0; uint32_t gc = 0; (void)gc;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait<true>(&full[stage], phase);
                if (warp == 2 && lane == 0) TL(gc, 5);
                const uint32_t src = smem_u32(stage_a(stage)) + row * 128;
                uint4 v[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = lds128(src + (uint32_t)((c ^ (row & 7)) << 4));   // undo the 128B swizzle: logical chunk c
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t w[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        hi[4 * c + e] = w[e] & 0xFFFFE000u;
                        lo[4 * c + e] = __float_as_uint(__uint_as_float(w[e]) - __uint_as_float(hi[4 * c + e])) & 0xFFFFE000u;
                    }
                }
                const uint32_t dst = tmem_base + C::TMEM_A0 + stage * 64 + lane_addr;
                tmem_st32(dst, hi);
                tmem_st32(dst + 32, lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&conv[stage]);                     // one arrival per warp
                if (warp == 2 && lane == 0) TL(gc, 6);
                ++gc;
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp < 14) {
        // ============================================================ accumulate + epilogue (warps 6..13, 256 threads)
        constexpr int HALF = C::HALF;                                         // columns owned by this warp
        constexpr int CH = HALF >= 32 ? 32 : 16;                              // columns per tcgen05.ld
        constexpr int NCH = HALF / CH;
        constexpr int LD = C::S_LD;
        const int quarter = warp & 3;                                         // TMEM lanes [32*quarter, 32*quarter+32)
        const int half = (warp - 6) >> 2;
        const uint32_t lane_col = ((uint32_t)(quarter * 32) << 16) + half * HALF;
        const uint32_t S = smem_u32(slab_all + (warp - 6) * 32 * LD);         // this warp's slab [32][LD] (shared-space address)
        const uint32_t par = smem_u32(par_all + (warp - 6) * 2 * HALF);       // scale[HALF] | shift[HALF]
        // store mapping: LPR lanes cover the CW contiguous channels of one pixel (CW=32: a full 128-byte line per pixel and
        // store instruction), RPI pixels per instruction, NST instructions per 32-pixel chunk
        constexpr int CW = C::EPI_CW, LPR = CW / 4, RPI = 32 / LPR, NST = 32 / RPI;
        const int sub = lane % LPR, rbase = lane / LPR;
        // residual prefetch mapping: RL lanes cover the HALF contiguous channels of one pixel, RR pixels per instruction
        constexpr int RL = HALF / 4, RR = 32 / RL;
        const int rcol = (lane % RL) * 4, rrow = lane / RL;
        const float* __restrict__ rsrc = p.res1 ? p.res1 : p.res2;            // the graphs never use res1 and res2 together
        const int r_cs = p.res1 ? p.res1_cs : p.res2_cs, r_co = p.res1 ? p.res1_co : p.res2_co;
        const bool r_pre = p.res1 != nullptr, r_post = !r_pre && rsrc != nullptr;
        const bool r_resize = !r_pre && p.res2_h != 0;                        // nearest resize of the added map (_layers.py:137-142)
        const float rs_y = r_resize ? (float)p.res2_h / (float)p.Ho : 1.f, rs_x = r_resize ? (float)p.res2_w / (float)p.Wo : 1.f;
        const float neg_slope = p.act == FCP_ACT_NONE ? 1.f : (p.act == FCP_ACT_RELU ? 0.f : p.slope);
        const float post_scale = p.post_scale;
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int n_tile = tile % p.tiles_n, m_tile = tile / p.tiles_n;
            const int img = m_tile / tiles_per_img, rem = m_tile - img * tiles_per_img;
            const int ho0 = (rem / p.tiles_x) * p.BH, wo0 = (rem % p.tiles_x) * BW;
            const int n0 = n_tile * BN + half * HALF;                         // first channel of this warp
            const int img_pix0 = img * p.Ho * p.Wo;
            // ---- this warp's folded-BN scale/shift into shared memory (arrays are padded to cout_pad >= n0 + HALF)
            for (int j = lane; j < HALF; j += 32) {
                sts_f1(par + 4 * j, __ldg(p.scale + n0 + j));
                sts_f1(par + 4 * (HALF + j), __ldg(p.shift + n0 + j));
            }
            // ---- residual sub-tile (32 pixels x HALF channels) streams into the slab while the K loop runs
            if (rsrc && !(p.exp_nolo & 8)) {
#pragma unroll 2
                for (int it = 0; it < 32 / RR; ++it) {
                    const int row = it * RR + rrow, prow = quarter * 32 + row;
                    const int ho = ho0 + (prow >> p.bw_log2), wo = wo0 + (prow & (BW - 1));
                    if (ho >= p.Ho || wo >= p.Wo || n0 + rcol + 3 >= p.Cout) continue;
                    int rp = img_pix0 + ho * p.Wo + wo;
                    if (r_resize) {
                        const int hs = min((int)floorf(ho * rs_y), p.res2_h - 1), ws = min((int)floorf(wo * rs_x), p.res2_w - 1);
                        rp = (img * p.res2_h + hs) * p.res2_w + ws;
                    }
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(S + 4 * (row * LD + rcol)),
                                 "l"(rsrc + (size_t)rp * r_cs + r_co + n0 + rcol) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            float acc[HALF];
#pragma unroll
            for (int j = 0; j < HALF; ++j) acc[j] = 0.f;
            if (warp == 6 && lane == 0) TL(g, 13);                            // tile prologue (params, residual prefetch) issued
            for (int kb = 0; kb < kblocks; ++kb, ++g) {
                const uint32_t buf = g & 1;
                mbar_wait<true>(&d_full[buf], (g >> 1) & 1);
                if (warp == 6 && lane == 0) TL(g, 7);
                tc_fence_after();
#pragma unroll
                for (int cb = 0; cb < NCH; ++cb) {
                    float v[CH];
                    tmem_ld<CH>(tmem_base + buf * BN + lane_col + cb * CH, v);
#pragma unroll
                    for (int j = 0; j < CH; ++j) acc[cb * CH + j] += v[j];     // round-to-nearest fp32 running sum
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&d_empty[buf]);
                if (warp == 6 && lane == 0) TL(g, 8);
            }
            // ---- fused epilogue, phase 1 (pixel per thread == TMEM lane, in place in the slab):
            //      y = post_scale * act(acc * scale + shift [+ res1]) [+ res2]
            if (rsrc) asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            if (warp == 6 && lane == 0) TL(g - 1, 10);
            if (!(p.exp_nolo & 4)) {
                const uint32_t row = S + 4 * lane * LD;
#pragma unroll
                for (int j = 0; j < HALF; j += 4) {
                    const float4 sc = lds_f4(par + 4 * j);
                    const float4 sh = lds_f4(par + 4 * (HALF + j));
                    float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rsrc) rv = lds_f4(row + 4 * j);
                    float4 x;
                    x.x = acc[j] * sc.x + sh.x; x.y = acc[j + 1] * sc.y + sh.y; x.z = acc[j + 2] * sc.z + sh.z; x.w = acc[j + 3] * sc.w + sh.w;
                    if (r_pre) { x.x += rv.x; x.y += rv.y; x.z += rv.z; x.w += rv.w; }
                    x.x = x.x > 0.f ? x.x : x.x * neg_slope; x.y = x.y > 0.f ? x.y : x.y * neg_slope;
                    x.z = x.z > 0.f ? x.z : x.z * neg_slope; x.w = x.w > 0.f ? x.w : x.w * neg_slope;
                    x.x *= post_scale; x.y *= post_scale; x.z *= post_scale; x.w *= post_scale;
                    if (r_post) { x.x += rv.x; x.y += rv.y; x.z += rv.z; x.w += rv.w; }
                    sts_f4(row + 4 * j, x);
                }
            }
            __syncwarp();
            if (warp == 6 && lane == 0) TL(g - 1, 11);
            // ---- phase 2: read the slab back channel-contiguous; coalesced float4 stores (+ the RRDB second residual)
            auto out_pixel = [&](int st) -> int {                             // output pixel of staged row rbase + RPI*st, -1 = outside
                const int prow = quarter * 32 + rbase + RPI * st;
                const int ho = ho0 + (prow >> p.bw_log2), wo = wo0 + (prow & (BW - 1));
                return (ho < p.Ho && wo < p.Wo) ? img_pix0 + ho * p.Wo + wo : -1;
            };
            if (warp == 6 && lane == 0) TL(g - 1, 12);
            // rolled loops on purpose: the epilogue runs once per tile, and straight-line code that does not fit the
            // instruction caches is paced by instruction fetch (measured), not by the memory system
            const bool has3 = p.res3 != nullptr;
#pragma unroll 1
            for (int cc = 0; cc < HALF / CW; ++cc) {
                const int cl = cc * CW + sub * 4, n = n0 + cl;                // channel within the slab / within the tensor
                if (n + 3 < p.Cout) {
                    float* const outp = p.out + p.out_co + n;
                    const float* const r3p = p.res3 + p.res3_co + n;
#pragma unroll 2
                    for (int st = 0; st < NST; ++st) {
                        const int m = out_pixel(st);
                        if (m < 0) continue;
                        float4 x = lds_f4(S + 4 * ((rbase + RPI * st) * LD + cl));
                        if (has3) {
                            const float4 r3 = __ldg(reinterpret_cast<const float4*>(r3p + (size_t)m * p.res3_cs));
                            x.x = x.x * p.post_scale2 + r3.x; x.y = x.y * p.post_scale2 + r3.y;
                            x.z = x.z * p.post_scale2 + r3.z; x.w = x.w * p.post_scale2 + r3.w;
                        }
                        if (!(p.exp_nolo & 2)) *reinterpret_cast<float4*>(outp + (size_t)m * p.out_cs) = x;
                    }
                } else if (n < p.Cout) {
                    // ragged tail (Cout not a multiple of 4, e.g. the 19-class logits; never carries a residual): scalar stores
#pragma unroll 1
                    for (int st = 0; st < NST; ++st) {
                        const int m = out_pixel(st);
                        if (m < 0) continue;
#pragma unroll 1
                        for (int e = 0; e < 4 && n + e < p.Cout; ++e) {
                            float x = lds_f1(S + 4 * ((rbase + RPI * st) * LD + cl + e));
                            if (p.res3) x = x * p.post_scale2 + p.res3[(size_t)m * p.res3_cs + p.res3_co + n + e];
                            p.out[(size_t)m * p.out_cs + p.out_co + n + e] = x;
                        }
                    }
                }
            }
            __syncwarp();                                                     // slab + params are rewritten by the next tile
            if (warp == 6 && lane == 0) TL(g - 1, 9);                        // epilogue of this tile finished
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// ----------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

bool make_map(CUtensorMap* map, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, base, dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN>
int launch(fcp_ctx* ctx, const TcParams& p) {
    using C = Cfg<BN>;
    static bool configured = false;
    if (!configured) {
        FCP_CUDA(ctx, cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    int grid = p.num_tiles < ctx->sm_count ? p.num_tiles : ctx->sm_count;
    conv_tc_kernel<BN><<<grid, NUM_THREADS, C::SMEM_BYTES, ctx->stream>>>(p);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

}  // namespace

bool conv_tc_supported(const ConvOp& op) {
    const ConvWeights& wt = *op.wt;
    if (wt.cin % KB != 0 || op.up_in) return false;
    if (op.stride != 1 && op.stride != 2) return false;
    if (op.stride == 2 && !(wt.k == 1 || wt.k == 3)) return false;
    if ((op.in.cs | op.in.co) & 3) return false;
    if ((size_t)op.out.h * op.out.w < 64) return false;       // pooled 1x1 maps etc. stay on the CUDA-core kernel
    if (op.res1 && op.res2) return false;                     // the epilogue prefetches one residual source (no graph uses both)
    if ((op.res1 || op.res2) && wt.cout % 4 != 0) return false;  // residual rows are prefetched in 16-byte pieces
    if ((size_t)op.out.n * op.out.h * op.out.w >= ((size_t)1 << 31)) return false;   // 32-bit pixel indices in the kernel
    if (op.act == FCP_ACT_SIGMOID) return false;              // one activation form (leaky with slope 0 / 1 / s) in the compact epilogue                     // the epilogue prefetches one residual source (no graph uses both)
    return encode_fn() != nullptr;
}

int launch_conv_tc(fcp_ctx* ctx, const ConvOp& op) {
    const ConvWeights& wt = *op.wt;
    if (!conv_tc_supported(op)) return fail(ctx, FCP_ERR_INVALID, "conv_tc: unsupported shape");
    if (op.in.c != wt.cin || op.out.c != wt.cout) return fail(ctx, FCP_ERR_INVALID, "conv: channel mismatch");
    TcParams p{};
    const int H = op.in.h, W = op.in.w, cs = op.in.cs;
    p.N = op.in.n; p.Ho = op.out.h; p.Wo = op.out.w; p.Cout = wt.cout; p.Cin = wt.cin;
    p.KH = p.KW = wt.k; p.stride = op.stride; p.pad = op.pad;
    if ((H + 2 * op.pad - wt.k) / op.stride + 1 != p.Ho || (W + 2 * op.pad - wt.k) / op.stride + 1 != p.Wo)
        return fail(ctx, FCP_ERR_INVALID, "conv: output shape mismatch");
    // spatial box of 128 output pixels: widest power-of-two width that does not overshoot the row by more than 2x
    int bw_log2 = 7;
    while (bw_log2 > 0 && (1 << bw_log2) >= 2 * p.Wo) --bw_log2;
    if (bw_log2 < 3) bw_log2 = 3;
    int best = -1; size_t best_tiles = 0;
    for (int l = 3; l <= 7; ++l) {                            // pick the box with the fewest tiles (ties: squarer)
        int bw = 1 << l, bh = TILE_M / bw;
        size_t t = (size_t)((p.Wo + bw - 1) / bw) * ((p.Ho + bh - 1) / bh);
        if (best < 0 || t < best_tiles || (t == best_tiles && abs(l - 4) < abs(best - 4))) { best = l; best_tiles = t; }
    }
    bw_log2 = best;
    const int BW = 1 << bw_log2, BH = TILE_M / BW;
    p.bw_log2 = bw_log2; p.BH = BH;
    p.tiles_x = (p.Wo + BW - 1) / BW; p.tiles_y = (p.Ho + BH - 1) / BH;
    const int BN = wt.cout_pad % 128 == 0 ? 128 : (wt.cout_pad % 64 == 0 ? 64 : 32);
    p.tiles_n = wt.cout_pad / BN;
    p.num_tiles = p.N * p.tiles_x * p.tiles_y * p.tiles_n;
    // ---- tensor maps
    float* base = op.in.p + op.in.co;
    const int nviews = op.stride == 2 ? 4 : 1;
    for (int v = 0; v < nviews; ++v) {
        const int py = v >> 1, px = v & 1, st = op.stride;
        cuuint64_t dims[4] = {(cuuint64_t)wt.cin, (cuuint64_t)((W - px + st - 1) / st), (cuuint64_t)((H - py + st - 1) / st), (cuuint64_t)p.N};
        cuuint64_t strides[3] = {(cuuint64_t)st * cs * 4, (cuuint64_t)st * W * cs * 4, (cuuint64_t)H * W * cs * 4};
        cuuint32_t box[4] = {KB, (cuuint32_t)BW, (cuuint32_t)BH, 1};
        if (dims[1] == 0 || dims[2] == 0) { dims[1] = dims[1] ? dims[1] : 1; dims[2] = dims[2] ? dims[2] : 1; }
        if (!make_map(&p.tmA[v], base + ((size_t)py * W + px) * cs, 4, dims, strides, box))
            return fail(ctx, FCP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the activation tensor");
    }
    const cuuint64_t K = (cuuint64_t)wt.k * wt.k * wt.cin;
    cuuint64_t bdims[2] = {K, (cuuint64_t)wt.cout_pad};
    cuuint64_t bstr[1] = {K * 4};
    cuuint32_t bbox[2] = {KB, (cuuint32_t)BN};
    if (!make_map(&p.tmBhi, wt.w_hi, 2, bdims, bstr, bbox) || !make_map(&p.tmBlo, wt.w_lo, 2, bdims, bstr, bbox))
        return fail(ctx, FCP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the weights");
    p.out = op.out.p; p.out_cs = op.out.cs; p.out_co = op.out.co;
    p.scale = wt.scale; p.shift = wt.shift;
    p.res1 = op.res1; p.res1_cs = op.res1_cs; p.res1_co = op.res1_co;
    p.act = op.act; p.slope = op.slope;
    p.post_scale = op.post_scale; p.res2 = op.res2; p.res2_cs = op.res2_cs; p.res2_co = op.res2_co; p.res2_h = op.res2_h; p.res2_w = op.res2_w;
    p.post_scale2 = op.post_scale2; p.res3 = op.res3; p.res3_cs = op.res3_cs; p.res3_co = op.res3_co;
    static const int exp_nolo = getenv("FCP_EXP_NOLO") ? atoi(getenv("FCP_EXP_NOLO")) : 0;
    p.exp_nolo = exp_nolo;
    auto do_launch = [&]() -> int {
        if (BN == 128) return launch<128>(ctx, p);
        if (BN == 64) return launch<64>(ctx, p);
        return launch<32>(ctx, p);
    };
#ifdef FCP_EXP_TIMELINE
    // clock64 stamps of CTA 0 for one launch of the shape FCP_TL_SHAPE="k,cin,cout" (default 3,256,256), K-blocks
    // [FCP_TL_G0, FCP_TL_G0+FCP_TL_N) of launch number FCP_TL_SHOT of that shape
    static int shots = 0;
    auto envi = [](const char* n, int d) { const char* v = getenv(n); return v ? atoi(v) : d; };
    int tk = 3, tcin = 256, tcout = 256;
    if (const char* sh = getenv("FCP_TL_SHAPE")) sscanf(sh, "%d,%d,%d", &tk, &tcin, &tcout);
    const int shot = envi("FCP_TL_SHOT", 0), g0 = envi("FCP_TL_G0", 100), gn = envi("FCP_TL_N", 24);
    const bool shoot = getenv("FCP_TC_TIMELINE") && wt.k == tk && wt.cin == tcin && wt.cout == tcout && shots <= shot &&
                       p.N * p.Ho * p.Wo >= 16384;
    long long* dbg = nullptr;
    if (shoot) { cudaMalloc(&dbg, 512 * 16 * 8); cudaMemset(dbg, 0, 512 * 16 * 8); p.dbg = dbg; ++shots; }
    const bool print_it = shoot && shots == shot + 1;
    int rc = do_launch();
    if (shoot && !print_it) { cudaStreamSynchronize(ctx->stream); cudaFree(dbg); }
    if (print_it) {
        cudaStreamSynchronize(ctx->stream);
        std::vector<long long> h(512 * 16);
        cudaMemcpy(h.data(), dbg, 512 * 16 * 8, cudaMemcpyDeviceToHost);
        long long t0 = h[g0 * 16 + 0];
        fprintf(stderr, "[timeline] k%d cin%d cout%d BN=%d res=%d kblocks/tile=%d tiles=%d\n", tk, tcin, tcout, BN, (int)(op.res1 || op.res2),
                wt.k * wt.k * wt.cin / KB, p.num_tiles);
        fprintf(stderr, "[timeline] g: prod_wait_empty prod_issued | mma_dempty mma_conv mma_issued | conv_full conv_done | drain_dfull drain_done | epi_done res_landed phase1_done om_done | tile_prologue_done\n");
        for (int g = g0; g < g0 + gn && g < 512; ++g) {
            fprintf(stderr, "[timeline] %3d:", g);
            for (int e = 0; e < 14; ++e) fprintf(stderr, " %7lld", h[g * 16 + e] ? h[g * 16 + e] - t0 : -1);
            fprintf(stderr, "\n");
        }
        cudaFree(dbg);
    }
    return rc;
#else
    return do_launch();
#endif
}

}  // namespace fcp
