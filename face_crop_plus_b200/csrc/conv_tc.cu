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
//   warp 0      TMA producer          full_a[l] <- A bytes (waits a_free[l]);  full_b[s] <- weight bytes (waits empty[s])
//   warps 2-5   hi/lo converters      a_free[l], conv[s] <- 4 arrivals (1/warp) (wait full_a[l], then empty[s] for the TMEM slot)
//   warps 1,14,(15: BN<=64)  MMA issuers (1 lane each, K-blocks round-robin, one partial-sum buffer each)  empty[s], d_full[b] <- tcgen05.commit (wait conv[s], full_b[s], d_empty[b])
//   warps 6-13  accumulate+epilogue   d_empty[b] <- 8 arrivals (1/warp) (wait d_full[b]); per K-block TMEM -> regs (+=),
//               after the last K-block: fused epilogue -> HBM.  Warp w owns TMEM lanes 32*(w%4).. and half of the columns.
// Two TMEM partial-sum buffers let the drain of K-block i overlap the MMAs of K-block i+1.
//
// MODE 1 (default, "f16x3"): the same split scheme on the kind::f16 pipe, which runs at twice the TF32 rate.  fp16 has the
//   same 11-bit significand as TF32 but only a 5-bit exponent, so the operands are BLOCK-SCALED by exact powers of two:
//   weights once per layer on the host (max |w*bn_scale| -> [2^13, 2^14)), activations per (pixel row, K-block) by the
//   converter warps (row max -> [2^14, 2^15)); hi = rn_f16(x*s), lo = rn_f16(x*s - hi) (round-to-nearest split: the
//   representation error is <= 2^-23 relative, 4x smaller than the truncating TF32 split).  The inverse scale of the row
//   travels through a 16-slot ring in tensor memory to the drain warps, which fold it into the per-K-block accumulate
//   (acc = fma(partial, 2^-k, acc): exact scaling, one rounding - the same rounding the TF32 mode's add performs).
//   A K-block is 64 channels (two 32-channel TMA boxes -> one 128-byte fp16 row per operand) = 4 K=16 steps x 3 MMAs, the
//   same 12 MMAs per drain as MODE 0 for twice the K.  Accuracy vs fp64: tests/test_gpu_parity.py (f16x3 <= tf32x3 <= fp32 FMA).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <type_traits>
#include <vector>

#include "common.h"

namespace fcp {

namespace {

constexpr int TILE_M = 128;
constexpr int KB = 32;                         // channels per K-block (128 bytes of fp32)
constexpr int A_TILE_BYTES = TILE_M * KB * 4;  // 16 KiB
// MMA-issuing warps: warp 1, warp 14 and (BN <= 64 only) warp 15, one lane each, K-blocks round-robin.  Invariant: there
// are exactly as many partial-sum buffers in tensor memory as issuers and issuer i only ever writes buffer i, so a parity
// wait on d_empty can never be more than one phase ahead (3 issuers sharing 2 buffers touch a buffer out of order and
// the mbarrier parity waits alias -> corrupted sums and a deadlock; measured the hard way).
constexpr int MAX_ISSUERS = 3;
constexpr int NUM_THREADS = 32 * (14 + MAX_ISSUERS - 1);

struct alignas(64) TcParams {
    CUtensorMap tmA[4];                        // input, one per (row parity, column parity); stride 1 uses [0]
    CUtensorMap tmBhi, tmBlo;                  // weights [cout_pad][K], K-major
    CUtensorMap tmOut, tmRes;                  // output / residual tensor, box = the tile's 128 pixels x 32 channels
    int out_tma, res_tma;                      // epilogue data paths: bulk tensor store / load usable for this launch
    int N, Ho, Wo, Cout, Cin, KH, KW, stride, pad, stride_w, pad_w;   // stride / pad: vertical; *_w: horizontal
    int cin_p;                                 // channels per tap in the packed weight matrix (MODE 1: Cin rounded up to 32)
    int w_exp;                                 // MODE 1: the packed weights are w * 2^w_exp
    // stem mode (MODE 1 only): the 7x7 / stride 2 / pad 3 detector stem straight from the uint8 RGB image.  The producer lands
    // one uint8 halo tile per output tile (2*BH+5 rows x 6*BW+15 bytes + alignment slack, TMA, zero filled outside the image); the converters
    // cut each pixel's 7 horizontal taps x 3 channels (21 contiguous bytes per tap row) out of it, subtract the channel
    // means and write exact fp16 integers into tensor memory.  K-block b = tap rows 2b, 2b+1 (32 K slots each, 21 used).
    CUtensorMap tmStem;
    int stem, stem_H, stem_W, stem_box_w, stem_box_h;
    int src2_units;                            // != 0: 32-channel units >= src2_units come from the second source (tmA[1], 1x1 stride-1 convs)
    int out_add;                               // TMA epilogue: the result is ADDED to the output tensor (cp.reduce.async.bulk .add)
    int act_cols;                              // channels >= act_cols skip the activation (general epilogue form only)
    int single;                                // opt-in fast mode: hi x hi only (one pass, 11 significant bits per operand)
    int a_exact;                               // the activations are exactly representable in 11 significant bits (u8 - mean):
                                               //   a_lo == 0, the a_lo*w_hi MMAs are skipped
    int tiles_x, tiles_y, tiles_n, num_tiles, bw_log2, BH;
    int pair_tiles;                            // PAIR kernels: tiles per CTA pair = tiles_n * ceil(m_tiles / 2)
    float* out; int out_cs, out_co;
    const float* scale; const float* shift;
    const float* res1; int res1_cs, res1_co;
    int act; float slope;
    float post_scale; const float* res2; int res2_cs, res2_co, res2_h, res2_w;
    float post_scale2; const float* res3; int res3_cs, res3_co;
    int ablate;                                // FCP_TC_ABLATE bit mask, measurement only (results are WRONG with bits 1/2):
                                               //   1 skip the w_lo loads (L2->SM / smem-write traffic probe), 2 skip the bulk stores,
                                               //   4 converters skip the fp16 split arithmetic (MODE 1), 8 no early tile decode, 32 early residual prefetch (racy), 64 late shift fetch,
                                               //   16 back-off in the drain warps' d_full wait
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
// arrive that cannot be issued before `dep` has been computed: used to hand a TMA landing buffer back to the producer only
// after the values loaded from it have really arrived in registers (an ld.shared is asynchronous; an arrive that merely
// follows it in program order can overtake it, and the refill then races the load - seen as a few wrong 16-byte pieces of
// single pixel rows, a few times per thousand launches)
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, uint32_t dep) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)), "r"(dep) : "memory");
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
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// smem (128B-swizzled box) -> global tensor; completion tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// ... the same box ADDED to the tensor (fp32 add performed in L2): out += slab
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2, int c3) {
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk stores of this thread have finished READING shared memory (the slab may be overwritten)
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
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
#ifdef FCP_EXP_TIMELINE
#define TL(g, ev) do { if (p.dbg && blockIdx.x == 0 && (g) < 512) p.dbg[(g) * 16 + (ev)] = clock64(); } while (0)
#else
#define TL(g, ev) do { } while (0)
#endif
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (TMA store source)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc], kind::tf32, one CTA: A operand in tensor memory (rows = lanes, one 32-bit column
// per tf32 element), B through a shared-memory descriptor
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// same, kind::f16: A = packed fp16 pairs in tensor memory (one 32-bit column = K elements 2j (low half), 2j+1), K = 16
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st1(uint32_t taddr, uint32_t r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
__device__ __forceinline__ void tmem_ld1_issue(uint32_t taddr, uint32_t& r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
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
// PAIR kernels (cta_group::2: two CTAs of a cluster run one 256-row MMA, each holds half of the weight tile).
// The weight load of either CTA completes its bytes on the LEADER's barrier (shared::cluster address with the peer bit cleared).
__device__ __forceinline__ uint32_t leader_addr(const void* smem_ptr) { return smem_u32(smem_ptr) & 0xFEFFFFFFu; }
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(leader_addr(bar)), "r"(c0), "r"(c1) : "memory");
}
// arrive on the leader CTA's copy of `bar` (converters / drain warps of both CTAs report to the one MMA issuer)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(raddr) : "r"(smem_u32(bar)));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
// MMA-retired signal to the same-offset barrier of both CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_f16_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t cols) {     // same warp id in both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrives on `bar` when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp receives lane (base_lane + i).
// issue only (no wait): the caller overlaps the load with arithmetic on the previous piece and then calls tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row atoms 1024 bytes apart (SBO), LBO unused (=1),
// descriptor version 1 (Blackwell), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}


// Epilogue data path.  The 4 epilogue warps that share a channel half (one warp per TMEM lane quarter = 32 pixels) own a
// shared-memory slab of CHUNKS x [128 pixels][32 channels] fp32 in the 128B-swizzled layout TMA produces (16-byte piece c
// of pixel row r lives at r*128 + ((c ^ (r & 7)) << 4)); chunk q of the group is driven by lane 0 of its quarter-q warp:
//   1. it bulk-loads the tile's residual chunk (bottleneck shortcut / RRDB skip) into the slab while the K loop runs,
//   2. the fused epilogue math runs in place, pixel per thread (the TMEM lane layout; conflict-free thanks to the swizzle),
//   3. it bulk-stores the chunk to the output tensor (TMA clips rows / channels outside the tensor).
// One TMA operation costs the issuing thread 75-150 cycles of TMA-unit time whatever its size (measured), hence whole
// 16 KiB chunks per operation and a 128-thread named barrier per group instead of one operation per warp.
// The warps never touch global memory with per-thread instructions on this path: measured, the per-thread version spent
// ~2300 instructions per warp and tile on address arithmetic and was the bound of every layer with a short K loop.
// Fallbacks (rolled loops): resized residual (FPN top-down add) via cp.async, per-thread stores when the output is
// not 16-byte addressable (19-class logits) or carries the RRDB second residual.
template <int BN, int MODE> struct Cfg {
    static constexpr int KBLK = MODE ? 64 : 32;                           // channels per K-block (one 128-byte operand row)
    static constexpr int HALVES = MODE ? 2 : 1;                           // 32-channel fp32 landing units (TMA boxes) per K-block
    static constexpr int B_TILE_BYTES = BN * 128;                         // BN rows x (32 tf32 | 64 fp16)
    static constexpr int B_SLOT_BYTES = 2 * B_TILE_BYTES;                 // w_hi + w_lo of one K-block
    // One trip round the pipeline (TMA issue + landing + convert + MMA issue + retire) takes ~3300 cycles (measured), so
    // its depth sets the K-block period until the co-limiters of DESIGN.md 5.3 take over (TMA ingest of 64 KiB per K-block
    // at ~42 B/cycle/SM, ~1200 busy converter cycles, 144 KiB through the shared-memory port).  The fp32 A tile only lives from its
    // landing to its conversion, so it gets its own short ring (LANDINGS units of 16 KiB); a pipeline STAGE is a weight
    // slot in shared memory plus an (a_hi | a_lo) slot in tensor memory, both held until the K-block's MMAs retire.
    static constexpr int STAGES = MODE ? (BN >= 128 ? 3 : (BN == 64 ? 4 : 6)) : (BN >= 128 ? 4 : (BN == 64 ? 5 : 6));
    static constexpr int LANDINGS = MODE ? (BN >= 128 ? 4 : 6) : (BN >= 128 ? 2 : 3);
    // One tcgen05.mma costs its issuing thread ~85 cycles whatever N is, so the narrow BN = 64 tiles (32-cycle MMAs) are
    // issue-bound: BN <= 64 gets a third issuer + accumulator (tensor memory is 512 columns).
    static constexpr int ISSUERS = BN <= 64 ? 3 : 2;
    // tensor memory: [0, ISSUERS*BN) partial-sum buffers; then STAGES x (a_hi 32 cols | a_lo 32 cols); MODE 1: then the
    // ring of per-row inverse scales (one column per K-block in flight).  A scale slot is rewritten SCALE_SLOTS K-blocks
    // later; the converter of K-block g' waits for the retirement of K-block g' - STAGES, whose MMAs were issued after the
    // drain of K-block g' - STAGES - ISSUERS (and, the drain being in order, of every earlier one) had finished:
    // SCALE_SLOTS >= STAGES + ISSUERS makes the slot of K-block g safe to overwrite.
    static constexpr int TMEM_A0 = ISSUERS * BN;
    static constexpr int TMEM_SC0 = TMEM_A0 + STAGES * 64;
    static constexpr int SCALE_SLOTS = 16;
    static constexpr int TMEM_COLS = 512;
    static_assert(TMEM_SC0 + (MODE ? SCALE_SLOTS : 0) <= TMEM_COLS, "tensor memory budget");
    static_assert(SCALE_SLOTS >= STAGES + ISSUERS, "scale ring too short");
    static constexpr int EPI_WARPS = BN >= 64 ? 8 : 4;                    // BN=32: one warp per TMEM lane quarter
    static constexpr int HALF = BN >= 64 ? BN / 2 : BN;                   // channels owned by one epilogue warp
    static constexpr int GROUPS = EPI_WARPS / 4;                          // 4 warps (all 128 pixels) share HALF channels
    static constexpr int CHUNKS = HALF / 32;                              // 32-channel (128-byte) chunks per group
    static constexpr int CHUNK_BYTES = TILE_M * 128;                      // one chunk = the tile's 128 pixels x 32 channels
    static constexpr int S_BYTES = GROUPS * CHUNKS * CHUNK_BYTES;
    static constexpr int PAR_BYTES = GROUPS * 2 * HALF * 4;               // per group: shift[HALF] (+ spare)
    static constexpr int PIPE_BYTES = STAGES * B_SLOT_BYTES + LANDINGS * A_TILE_BYTES;
    static constexpr int SMEM_BYTES = PIPE_BYTES + 1024 /*align slack*/ + 512 /*barriers*/ + S_BYTES + PAR_BYTES;
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    static_assert(B_SLOT_BYTES % 1024 == 0 && B_TILE_BYTES % 1024 == 0, "operand tiles must stay 1024-byte aligned (swizzle atoms)");
};

// PAIR = 1: launched as clusters of two CTAs (cta_group::2).  A pair works on two pixel tiles of the same output-channel
// tile in lockstep: each CTA stages its own 128 pixel rows (activations -> converters -> its tensor memory) and HALF of
// every weight tile; the leader CTA issues one 256 x BN MMA for both, the partial sums land in each CTA's tensor memory and
// are drained and stored per CTA.  Per K-block an SM takes in 25 % fewer bytes (its weight tile is half as large) and the
// tensor core reads half the weight bytes from its own shared memory.
template <int BN, int MODE, int PAIR = 0>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
    using C = Cfg<BN, MODE>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // [STAGES x (w_hi | w_lo)] [LANDINGS x fp32 A tile] [slab: GROUPS x CHUNKS x 128 rows x 128 B] [params] [barriers]
    uint8_t* slab_all = smem + C::PIPE_BYTES;                                              // 1024-byte aligned
    float* par_all = reinterpret_cast<float*>(slab_all + C::S_BYTES);                      // [GROUPS][2][HALF]
    uint64_t* bars = reinterpret_cast<uint64_t*>(slab_all + C::S_BYTES + C::PAR_BYTES);
    uint64_t* full_b = bars;                       // [STAGES]   weights of the K-block have landed
    uint64_t* conv = bars + C::STAGES;             // [STAGES]   a_hi | a_lo of the K-block are in the stage's TMEM slot
    uint64_t* empty = bars + 2 * C::STAGES;        // [STAGES]   the K-block's MMAs have retired: weight slot + TMEM slot reusable
    uint64_t* full_a = bars + 3 * C::STAGES;       // [LANDINGS] fp32 A tile has landed
    uint64_t* a_free = full_a + C::LANDINGS;       // [LANDINGS] the converters have read the landing buffer
    uint64_t* d_full = a_free + C::LANDINGS;       // [ISSUERS]  partial sum of one K-block is complete in TMEM buffer b
    uint64_t* d_empty = d_full + C::ISSUERS;       // [ISSUERS]  buffer b has been drained to registers
    uint64_t* res_bar = d_empty + C::ISSUERS;      // [2]  (first GROUPS used) residual chunks of the group's tile have landed
    uint64_t* par_bar = res_bar + 2;               // [2]  (first GROUPS used) folded-BN shift of the group's tile has landed
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(par_bar + 2);
    static_assert((3 * C::STAGES + 2 * C::LANDINGS + 2 * C::ISSUERS + 4) * 8 + 4 <= 512, "barrier area");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto stage_b_hi = [&](int s) { return smem + s * C::B_SLOT_BYTES; };
    auto stage_b_lo = [&](int s) { return smem + s * C::B_SLOT_BYTES + C::B_TILE_BYTES; };
    auto landing = [&](int l) { return smem + C::STAGES * C::B_SLOT_BYTES + l * A_TILE_BYTES; };

    if (threadIdx.x == 0) {
        // PAIR: full_b / conv / d_empty are only waited on in the leader CTA, where both CTAs' producers, converters and drain
        // warps report; empty / d_full receive the leader's multicast commits in both CTAs
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full_b[s], 1); mbar_init(&conv[s], PAIR ? 8 : 4); mbar_init(&empty[s], 1); }
        for (int l = 0; l < C::LANDINGS; ++l) { mbar_init(&full_a[l], 1); mbar_init(&a_free[l], 4); }
        for (int a = 0; a < C::ISSUERS; ++a) { mbar_init(&d_full[a], 1); mbar_init(&d_empty[a], (PAIR ? 2 : 1) * C::EPI_WARPS); }
        for (int a = 0; a < 2; ++a) { mbar_init(&res_bar[a], C::CHUNKS); mbar_init(&par_bar[a], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    uint32_t cta_rank = 0;
    if constexpr (PAIR) {
        asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
        cluster_sync_all();                                               // both CTAs are resident, their barriers initialised
        if (warp == 1) tmem_alloc_pair(tmem_base_smem, C::TMEM_COLS);
    } else {
        if (warp == 1) tmem_alloc(tmem_base_smem, C::TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    if constexpr (PAIR) cluster_sync_all();
    // tile walk: a CTA (PAIR: a CTA pair) takes every tile_step-th tile; PAIR tile t = (n_tile, pixel tiles 2m, 2m+1)
    const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int tile_end = PAIR ? p.pair_tiles : p.num_tiles;
    auto m_tile_of = [&](int tile) { return PAIR ? 2 * (tile / p.tiles_n) + (int)cta_rank : tile / p.tiles_n; };

    // K runs over 32-channel units in (tap, channel) order; a K-block is HALVES consecutive units, so with Cin = 32, 96,
    // 160 ... a MODE 1 K-block pairs the last unit of one tap with the first of the next instead of carrying an empty half
    // (the packed weight matrix has the same order: pack_f16).  Only the very last K-block of a tile can be half empty.
    const int upt = (p.Cin + 31) >> 5;                                    // units per tap
    const int units = p.KH * p.KW * upt;
    const int kblocks = (units + C::HALVES - 1) / C::HALVES;
    const int BW = 1 << p.bw_log2;
    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        // ================================================================================== TMA producer
        if (lane == 0) {
            int stage = 0, land = 0; uint32_t phase = 0, lphase = 0; uint32_t gp = 0; (void)gp;
            for (int tile = tile0; tile < tile_end; tile += tile_step) {
                const int n_tile = tile % p.tiles_n, m_tile = m_tile_of(tile);
                const int img = m_tile / tiles_per_img, rem = m_tile - img * tiles_per_img;
                const int ho0 = (rem / p.tiles_x) * p.BH, wo0 = (rem % p.tiles_x) * BW;
                int tap = 0, ch = 0, r = 0, sx = 0;                          // next unit = (tap (r, sx), 32-channel chunk ch)
                for (int kb = 0; kb < kblocks; ++kb) {
                    const int kcol = tap * p.cin_p + ch * 32;
                    // the activation tile first (it has the longer way to go: landing -> converters -> tensor memory);
                    // MODE 1: two 32-channel boxes per K-block (the second one is missing in a trailing half block)
                    if (MODE && p.stem && kb == 0) {
                        mbar_wait<true>(&a_free[land], lphase ^ 1);
                        mbar_expect_tx(&full_a[land], (uint32_t)(p.stem_box_w * p.stem_box_h));
                        tma_load_3d(landing(land), &p.tmStem, &full_a[land], (6 * wo0 - 9) & ~15, 2 * ho0 - 3, img);   // TMA: 16-byte aligned innermost start
                        if (++land == C::LANDINGS) { land = 0; lphase ^= 1; }
                    }
#pragma unroll
                    for (int hf = 0; hf < C::HALVES; ++hf) {
                        if (tap >= p.KH * p.KW) break;
                        if (!(MODE && p.stem)) {
                            int dy = r - p.pad, dx = sx - p.pad_w, map = 0;
                            if (p.stride == 2) {   // input row 2*ho + dy lives in parity view (dy & 1) at row ho + (dy - (dy & 1)) / 2
                                const int py = dy & 1;
                                map = py * 2;
                                dy = (dy - py) >> 1;
                            }
                            if (p.stride_w == 2) {
                                const int px = dx & 1;
                                map += px;
                                dx = (dx - px) >> 1;
                            }
                            mbar_wait<true>(&a_free[land], lphase ^ 1);
                            if (hf == 0) TL(gp, 0);
                            mbar_expect_tx(&full_a[land], A_TILE_BYTES);
                            const CUtensorMap* am = &p.tmA[map];
                            int chc = ch * 32;
                            if (p.src2_units && ch >= p.src2_units) { am = &p.tmA[1]; chc -= p.src2_units * 32; }   // folded shortcut conv
                            tma_load_4d(landing(land), am, &full_a[land], chc, wo0 + dx, ho0 + dy, img);
                            if (++land == C::LANDINGS) { land = 0; lphase ^= 1; }
                        }
                        if (++ch == upt) { ch = 0; ++tap; if (++sx == p.KW) { sx = 0; ++r; } }
                    }
                    mbar_wait<true>(&empty[stage], phase ^ 1);
                    if constexpr (PAIR) {   // my half of the output channels (BN/2 rows of w_hi and w_lo), bytes counted by the leader
                        if (cta_rank == 0) mbar_expect_tx(&full_b[stage], 2 * C::B_TILE_BYTES);
                        const int row = n_tile * BN + (int)cta_rank * (BN / 2);
                        tma_load_2d_pair(stage_b_hi(stage), &p.tmBhi, &full_b[stage], kcol, row);
                        tma_load_2d_pair(stage_b_lo(stage), &p.tmBlo, &full_b[stage], kcol, row);
                    } else {
                        mbar_expect_tx(&full_b[stage], ((p.ablate & 1) ? 1 : 2) * C::B_TILE_BYTES);
                        tma_load_2d(stage_b_hi(stage), &p.tmBhi, &full_b[stage], kcol, n_tile * BN);
                        if (!(p.ablate & 1)) tma_load_2d(stage_b_lo(stage), &p.tmBlo, &full_b[stage], kcol, n_tile * BN);
                    }
                    TL(gp, 1); ++gp;
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 || warp >= 14) {
        // ============================================================== MMA issuers (warps 1, 14, 15: K-blocks round-robin)
        const uint32_t me = warp == 1 ? 0u : (uint32_t)(warp - 13);
        if (lane == 0 && me < (uint32_t)C::ISSUERS && cta_rank == 0) {     // PAIR: the leader CTA issues for both
            // instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10) | F16 (0, 0), both K-major, N>>3 at bit 17, M>>4 at bit 24
            const uint32_t idesc = (1u << 4) | (MODE ? 0u : ((2u << 7) | (2u << 10))) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(((PAIR ? 2 : 1) * TILE_M) >> 4) << 24);
            auto mma = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t acc) {
                if constexpr (PAIR) umma_f16_ts_pair(d, a, b, idesc, acc);
                else if constexpr (MODE) umma_f16_ts(d, a, b, idesc, acc); else umma_tf32_ts(d, a, b, idesc, acc);
            };
            // Issuing a tcgen05.mma blocks the thread for about its execution time, and every barrier wait costs a few
            // hundred cycles; with a single issuer the tensor pipe idled ~40% of each K-block.  ISSUERS threads take
            // K-blocks round-robin, so one thread's waits overlap the others' MMAs.  Issuer i owns partial-sum buffer i.
            uint32_t g = 0;                                                   // K-blocks so far (all tiles)
            uint32_t turn = 0, mine = 0;                                      // g % ISSUERS; K-blocks this thread has issued
            int stage = 0; uint32_t phase = 0;
            const uint32_t buf = me;
            const bool a_exact = p.a_exact != 0, single = p.single != 0;
            for (int tile = tile0; tile < tile_end; tile += tile_step) {
                for (int kb = 0; kb < kblocks; ++kb, ++g) {
                    if (turn == me) {
                        // operands first (normally long complete), the partial-sum buffer last: its release by the drain
                        // warps is the critical dependency of this thread's K-block chain
                        mbar_wait(&full_b[stage], phase);                     // weights landed
                        mbar_wait(&conv[stage], phase);                       // a_hi | a_lo converted into the stage's TMEM slot
                        TL(g, 3);
                        mbar_wait(&d_empty[buf], (mine & 1) ^ 1);             // partial-sum buffer drained
                        TL(g, 2);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + buf * BN;
                        const uint32_t a_hi = tmem_base + C::TMEM_A0 + stage * 64, a_lo = a_hi + 32;   // A operand: tensor memory
                        const uint64_t b_hi = umma_desc(smem_u32(stage_b_hi(stage))), b_lo = umma_desc(smem_u32(stage_b_lo(stage)));
                        // one K-step = 8 tf32 | 16 fp16: +8 TMEM columns for A, +32 bytes (+2 in the addr>>4 field) inside B's
                        // swizzle span.  MODE 1: a trailing half block (odd number of 32-channel units) only has 2 K-steps.
                        // Small terms first: while the accumulator is tiny its round-toward-zero losses are negligible.
                        const int ksteps = (MODE && 2 * kb + 1 >= units) ? 2 : 4;
                        if (single) {                                         // fast mode: no correction terms
                        } else if (!a_exact) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (k < ksteps) {
                                    mma(d_tmem, a_lo + 8 * k, b_hi + 2 * k, k != 0);
                                    mma(d_tmem, a_hi + 8 * k, b_lo + 2 * k, 1);
                                }
                            }
                        } else {                                              // a_lo == 0 (integer-valued activations)
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (k < ksteps) mma(d_tmem, a_hi + 8 * k, b_lo + 2 * k, k != 0);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k < ksteps) mma(d_tmem, a_hi + 8 * k, b_hi + 2 * k, (k != 0) | !single);
                        if constexpr (PAIR) {                                 // both CTAs' producers / converters / drain warps
                            umma_commit_pair(&empty[stage]);
                            umma_commit_pair(&d_full[buf]);
                        } else {
                            umma_commit(&empty[stage]);                       // smem slot + TMEM A slot reusable once these MMAs retire
                            umma_commit(&d_full[buf]);                        // partial sum of this K-block complete
                        }
                        TL(g, 4);
                        ++mine;
                    }
                    if (++turn == (uint32_t)C::ISSUERS) turn = 0;
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
        int stage = 0, land = 0; uint32_t phase = 0, lphase = 0; uint32_t gc = 0;
        for (int tile = tile0; tile < tile_end; tile += tile_step) {
            for (int kb = 0; kb < kblocks; ++kb) {
                const uint32_t dst = tmem_base + C::TMEM_A0 + stage * 64 + lane_addr;
                if constexpr (MODE == 0) {
                    mbar_wait<true>(&full_a[land], lphase);
                    if (warp == 2 && lane == 0) TL(gc, 5);
                    const uint32_t src = smem_u32(landing(land)) + row * 128;
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
                    uint32_t dep = 0;                                         // depends on all eight loads of every lane
#pragma unroll
                    for (int c = 0; c < 8; ++c) dep |= hi[4 * c];
                    dep = __reduce_or_sync(0xffffffffu, dep);
                    if (lane == 0) mbar_arrive_after(&a_free[land], dep);     // the landing buffer is in registers: refill it
                    if (++land == C::LANDINGS) { land = 0; lphase ^= 1; }
                    mbar_wait<true>(&empty[stage], phase ^ 1);                // the stage's TMEM slot: MMAs of K-block g - STAGES retired
                    tc_fence_after();
                    tmem_st32(dst, hi);
                    tmem_st32(dst + 32, lo);
                } else if (p.stem) {
                    // ---- stem: this thread's output pixel (ty, tx) of the tile; K-block kb = tap rows 2kb, 2kb+1.  The 7 taps x 3
                    //      channels of a tap row are 21 contiguous bytes of the uint8 halo tile; byte b becomes the fp16 integer
                    //      b - mean exactly: 0x6400 | b is fp16(1024 + b), minus fp16(1024 + mean).  Memory order R,G,B with
                    //      means 123,117,104 (the BGR flip of retinaface.py:450 lives in the packed weights).
                    if (kb == 0) {
                        mbar_wait<true>(&full_a[land], lphase);
                        if (warp == 2 && lane == 0) TL(gc, 5);
                    }
                    const int m_tile = m_tile_of(tile), rem = m_tile % tiles_per_img;
                    const int ho0 = (rem / p.tiles_x) * p.BH, wo0 = (rem % p.tiles_x) * BW;
                    const int ty = row >> p.bw_log2, tx = row & (BW - 1);
                    const int ho = ho0 + ty, wo = wo0 + tx;
                    const bool edge_x = 2 * wo0 - 3 < 0 || 2 * (wo0 + BW - 1) + 3 >= p.stem_W;        // tile-uniform
                    uint32_t vb = 0x1FFFFFu;                                                           // valid window bytes (21 bits)
                    if (edge_x) {
                        vb = 0;
#pragma unroll
                        for (int sx = 0; sx < 7; ++sx) {
                            const int xx = 2 * wo - 3 + sx;
                            if (xx >= 0 && xx < p.stem_W) vb |= 7u << (3 * sx);
                        }
                    }
                    const uint32_t tile_smem = smem_u32(landing(land));
                    const __half2 c0 = __floats2half2_rn(1024.f + 123.f, 1024.f + 117.f), c1 = __floats2half2_rn(1024.f + 104.f, 1024.f + 123.f),
                                  c2 = __floats2half2_rn(1024.f + 117.f, 1024.f + 104.f);
                    uint32_t hi[32];
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int r = 2 * kb + rr, y = 2 * ho + r - 3;
                        const bool row_ok = r < 7 && y >= 0 && y < p.stem_H;
                        const uint32_t a = tile_smem + (uint32_t)((2 * ty + r) * p.stem_box_w + ((6 * wo0 - 9) & 15) + 6 * tx);   // box starts at the aligned byte
                        const uint32_t a4 = a & ~3u, sh = (a & 3u) * 8;
                        uint32_t w[7];
#pragma unroll
                        for (int i = 0; i < 7; ++i) w[i] = r < 7 ? lds32(a4 + 4 * i) : 0u;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            uint32_t v = 0;
                            if (j < 11) {
                                const uint32_t q = __funnelshift_r(w[j >> 1], w[(j >> 1) + 1], sh);  // window bytes 4*(j>>1) .. +3
                                const uint32_t pair = __byte_perm(q, 0x64646464u, (j & 1) ? 0x4342u : 0x4140u);   // (1024 + b_odd) << 16 | (1024 + b_even)
                                const __half2 d = __hsub2(*reinterpret_cast<const __half2*>(&pair), j % 3 == 0 ? c0 : (j % 3 == 1 ? c1 : c2));
                                v = *reinterpret_cast<const uint32_t*>(&d);
                                uint32_t keep = ((vb >> (2 * j)) & 1u ? 0x0000FFFFu : 0u) | ((vb >> (2 * j + 1)) & 1u ? 0xFFFF0000u : 0u);
                                if (!row_ok) keep = 0;
                                v &= keep;
                            }
                            hi[16 * rr + j] = v;
                        }
                    }
                    if (kb == kblocks - 1) {                                  // the halo tile has been read for the last time: refill it
                        uint32_t dep = 0;                                     // (after the loads have really completed, see mbar_arrive_after)
#pragma unroll
                        for (int j = 0; j < 11; ++j) dep ^= hi[j] ^ hi[16 + j];
                        dep = __reduce_or_sync(0xffffffffu, dep);
                        if (lane == 0) mbar_arrive_after(&a_free[land], dep);
                        if (++land == C::LANDINGS) { land = 0; lphase ^= 1; }
                    }
                    mbar_wait<true>(&empty[stage], phase ^ 1);                // the stage's TMEM slot: MMAs of K-block g - STAGES retired
                    tc_fence_after();
                    tmem_st16(dst, hi);
                    tmem_st16(dst + 16, hi + 16);
                    int ie = 127 - p.w_exp;                                   // the activations are unscaled integers: 1 / 2^w_exp
                    tmem_st1(tmem_base + C::TMEM_SC0 + (gc & (C::SCALE_SLOTS - 1)) + lane_addr, (uint32_t)ie << 23);
                } else {
                    // ---- block-scaled fp16 split of this thread's pixel row (64 channels, or 32 in a trailing half block)
                    const bool two = 2 * kb + 1 < units;
                    float x[64];
                    const int land0 = land;
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        if (hf == 0 || two) {
                            mbar_wait<true>(&full_a[land], lphase);
                            if (hf == 0 && warp == 2 && lane == 0) TL(gc, 5);
                            const uint32_t src = smem_u32(landing(land)) + row * 128;
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const uint4 q = lds128(src + (uint32_t)((c ^ (row & 7)) << 4));           // logical 16-byte chunk c
                                x[32 * hf + 4 * c] = __uint_as_float(q.x); x[32 * hf + 4 * c + 1] = __uint_as_float(q.y);
                                x[32 * hf + 4 * c + 2] = __uint_as_float(q.z); x[32 * hf + 4 * c + 3] = __uint_as_float(q.w);
                            }
                            if (++land == C::LANDINGS) { land = 0; lphase ^= 1; }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) x[32 + j] = 0.f;
                        }
                    }
                    // row maximum -> power-of-two scale that puts it into [2^14, 2^15) (fp16 max 65504); rows of zeros and
                    // magnitudes below 2^-63 use the scale of 2^-63 (their contribution is below fp32 resolution anyway)
                    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
#pragma unroll
                    for (int j = 0; j < 64; j += 4) {
                        m0 = fmaxf(m0, fabsf(x[j])); m1 = fmaxf(m1, fabsf(x[j + 1]));
                        m2 = fmaxf(m2, fabsf(x[j + 2])); m3 = fmaxf(m3, fabsf(x[j + 3]));
                    }
                    const uint32_t mbits = __float_as_uint(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
                    // the landing units are in registers now (the maximum depends on every loaded value): refill them
                    const uint32_t all_loaded = __reduce_max_sync(0xffffffffu, mbits);
                    if (lane == 0) {
                        mbar_arrive_after(&a_free[land0], all_loaded);
                        if (two) mbar_arrive_after(&a_free[land0 + 1 == C::LANDINGS ? 0 : land0 + 1], all_loaded);
                    }
                    int e = (int)(mbits >> 23);                               // biased exponent of the row maximum (<= 254 for finite data)
                    e = e < 64 ? 64 : (e > 254 ? 254 : e);
                    const float sc = __uint_as_float((uint32_t)(268 - e) << 23);             // 2^(14 - (e - 127))
                    int ie = e - 14 - p.w_exp;                                               // 1 / (sc * 2^w_exp)
                    ie = ie < 1 ? 1 : (ie > 254 ? 254 : ie);
                    mbar_wait<true>(&empty[stage], phase ^ 1);                // the stage's TMEM slot: MMAs of K-block g - STAGES retired
                    tc_fence_after();
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        if (hf == 0 || two) {
                            uint32_t hi[16], lo[16];
                            if (p.ablate & 4) {                           // measurement only: no split arithmetic
#pragma unroll
                                for (int j = 0; j < 16; ++j) { hi[j] = __float_as_uint(x[32 * hf + 2 * j]) & 0x3FFF3FFFu; lo[j] = 0; }
                            } else
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float a0 = x[32 * hf + 2 * j] * sc, a1 = x[32 * hf + 2 * j + 1] * sc;
                                const __half2 h = __floats2half2_rn(a0, a1);  // low half = even channel (the K order of a TMEM column)
                                const float2 hf2 = __half22float2(h);
                                const __half2 l = __floats2half2_rn(a0 - hf2.x, a1 - hf2.y);
                                hi[j] = *reinterpret_cast<const uint32_t*>(&h);
                                lo[j] = *reinterpret_cast<const uint32_t*>(&l);
                            }
                            tmem_st16(dst + 16 * hf, hi);
                            tmem_st16(dst + 32 + 16 * hf, lo);
                        }
                    }
                    tmem_st1(tmem_base + C::TMEM_SC0 + (gc & (C::SCALE_SLOTS - 1)) + lane_addr, (uint32_t)ie << 23);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if constexpr (PAIR) mbar_arrive_leader(&conv[stage]); else mbar_arrive(&conv[stage]); }   // one arrival per warp
                if (warp == 2 && lane == 0) TL(gc, 6);
                ++gc;
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp < 14) {
        // ============================================================ accumulate + epilogue (warps 6..13, 256 threads)
        constexpr int HALF = C::HALF;                                         // columns owned by this warp
        constexpr int CHUNKS = C::CHUNKS;
        const int ew = warp - 6;                                              // epilogue warp index
        if (ew < C::EPI_WARPS) {
        const int quarter = warp & 3;                                         // TMEM lanes [32*quarter, 32*quarter+32)
        const int half = ew >> 2;
        const uint32_t lane_col = ((uint32_t)(quarter * 32) << 16) + half * HALF;
        const uint32_t S = smem_u32(slab_all + half * CHUNKS * C::CHUNK_BYTES);   // the group's slab: CHUNKS x [128 rows][128 B], swizzled
        const uint32_t par = smem_u32(par_all + half * 2 * HALF);             // the group's scale[HALF] | shift[HALF]
        uint64_t* const rbar = &res_bar[half];
        uint64_t* const pbar = &par_bar[half];
        // swizzled address of 16-byte piece c (4 channels) of this warp's pixel row r (0..31) in chunk q
        auto slab_addr = [&](int q, int r, int c) -> uint32_t {
            return S + q * C::CHUNK_BYTES + (quarter * 32 + r) * 128 + ((c ^ (r & 7)) << 4);
        };
        auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory"); };   // the 4 warps of this half
        const bool has_res = p.res1 != nullptr || p.res2 != nullptr;          // the graphs never use res1 and res2 together
        const bool r_pre = p.res1 != nullptr, r_post = !r_pre && has_res;
        const float neg_slope = p.act == FCP_ACT_NONE ? 1.f : (p.act == FCP_ACT_RELU ? 0.f : p.slope);
        const float post_scale = p.post_scale;
        const bool dma = lane == 0 && quarter < CHUNKS;                       // this thread drives the TMA traffic of chunk `quarter`
        uint32_t g = 0, tile_par = 0;
        uint32_t buf = 0, dpar = 0;                                           // partial-sum buffer of K-block g (g % ISSUERS) and its phase
        for (int tile = tile0; tile < tile_end; tile += tile_step, tile_par ^= 1) {
            // The tensor pipe is stalled at a tile boundary until this tile's first partial sums are drained, so nothing is
            // computed up front: the tile coordinates are decoded after the first K-block's drain.
            int img = 0, ho0 = 0, wo0 = 0, n0 = 0;
            auto out_pixel = [&](int r) -> int {                              // output pixel of this warp's row r, -1 = outside the image
                const int prow = quarter * 32 + r;
                const int ho = ho0 + (prow >> p.bw_log2), wo = wo0 + (prow & (BW - 1));
                return (ho < p.Ho && wo < p.Wo && img < p.N) ? (img * p.Ho + ho) * p.Wo + wo : -1;   // (img >= N: the odd tile out of a PAIR)
            };
            // ---- after the first K-block (tiles with a long K loop): decode the tile and fetch the group's folded-BN shift (the
            //      scale is folded into the weights; arrays are padded to cout_pad >= n0 + HALF) - safe to overwrite: every
            //      warp has passed the pre-store barrier of the previous tile, and with it its add_shift.  Then: the previous
            //      tile's bulk stores have released the slab (waited for by the threads that issued them, published by one
            //      group barrier), and the residual chunks (128 pixels x 32 channels each) stream into the slab while the K
            //      loop runs.
            // Tiles with a short K loop (<= 4 K-blocks, e.g. the bottleneck 1x1 convs) decode and fetch their folded-BN shift
            // at the TOP of the tile (with one K-block there is no second drain to hide the fetch behind).  The residual
            // prefetch stays behind the first drain's group barrier: issuing it at the top of the tile as well (right after
            // the issuing thread's own bulk_wait_read) measured 2-3 % faster, but made the network output depend on what
            // had run on the device before (1e-2 drift of the detector heads after a 13 GB RRDBNet run in another context,
            // profiles/diag_noise.py) - a hazard on the slab that only the group barrier closes.  FCP_TC_ABLATE=32 re-enables
            // it for measurements.
            const bool early = kblocks <= 4 && (!has_res || p.res_tma) && !(p.ablate & 8);
            // one thread per group copies the tile's HALF shift values (a bulk copy that completes on par_bar: single writer,
            // every reader waits for the barrier phase - the previous scheme, identical cp.async copies by all four warps,
            // was a benign but real write-write race that compute-sanitizer racecheck reports)
            auto fetch_params = [&]() {
                if (lane == 0 && quarter == 0) {
                    mbar_expect_tx(pbar, HALF * 4);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(par), "l"(p.shift + n0), "r"(HALF * 4), "r"(smem_u32(pbar)) : "memory");
                }
            };
            auto fetch_residual = [&]() {
                if (has_res && p.res_tma && dma) {
                    bulk_wait_read();
                    mbar_expect_tx(rbar, C::CHUNK_BYTES);
                    tma_load_4d(slab_all + (half * CHUNKS + quarter) * C::CHUNK_BYTES, &p.tmRes, rbar, n0 + quarter * 32, wo0, ho0, img);
                }
            };
            auto decode = [&]() {
                const int n_tile = tile % p.tiles_n, m_tile = m_tile_of(tile);
                img = m_tile / tiles_per_img;
                const int rem = m_tile - img * tiles_per_img;
                ho0 = (rem / p.tiles_x) * p.BH; wo0 = (rem % p.tiles_x) * BW;
                n0 = n_tile * BN + half * HALF;                               // first channel of this group
            };
            const bool early_par = early && !(p.ablate & 64), early_res = early && (p.ablate & 32);
            if (early) decode();
            if (early_par) fetch_params();
            if (early_res) fetch_residual();
            // The slab is released by the previous tile's bulk stores (they read it for ~1 500 cycles after being issued).  The
            // drain warps must not sit in that wait right after the first K-block - the MMA issuers are at most ISSUERS K-blocks
            // ahead and stall with them - so the wait (+ the group barrier that publishes it, + the residual prefetch into the
            // slab) moves to the drain of K-block `slab_kb`, by when the stores are long done; without a residual nobody
            // touches the slab before phase 1 and the wait moves there.
            const int slab_kb = (p.ablate & 128) ? 0 : (kblocks > 2 ? 2 : kblocks - 1);
            auto after_first_kblock = [&]() {
                if (!early) decode();
                if (!early_par) fetch_params();
            };
            auto slab_ready = [&]() {
                if (dma) bulk_wait_read();
                group_sync();
                if (!early_res) fetch_residual();
                if (has_res && !p.res_tma) {
                    // cp.async fallback (resized or unaligned residual): HALF/4 lanes cover one pixel, RR pixels per instruction
                    constexpr int RL = HALF / 4, RR = 32 / RL;
                    const int pc = lane % RL, rrow = lane / RL;               // 16-byte piece within the row
                    const float* rsrc = p.res1 ? p.res1 : p.res2;
                    const int r_cs = p.res1 ? p.res1_cs : p.res2_cs, r_co = p.res1 ? p.res1_co : p.res2_co;
                    const bool r_resize = !r_pre && p.res2_h != 0;            // nearest resize of the added map (_layers.py:137-142)
                    const float rs_y = (float)p.res2_h / (float)p.Ho, rs_x = (float)p.res2_w / (float)p.Wo;
#pragma unroll 1
                    for (int it = 0; it < 32 / RR; ++it) {
                        const int row = it * RR + rrow, prow = quarter * 32 + row;
                        const int ho = ho0 + (prow >> p.bw_log2), wo = wo0 + (prow & (BW - 1));
                        if (ho >= p.Ho || wo >= p.Wo || img >= p.N || n0 + pc * 4 + 3 >= p.Cout) continue;
                        int rp = (img * p.Ho + ho) * p.Wo + wo;
                        if (r_resize) {
                            const int hs = min((int)floorf(ho * rs_y), p.res2_h - 1), ws = min((int)floorf(wo * rs_x), p.res2_w - 1);
                            rp = (img * p.res2_h + hs) * p.res2_w + ws;
                        }
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(slab_addr(pc >> 3, row, pc & 7)),
                                     "l"(rsrc + (size_t)rp * r_cs + r_co + n0 + pc * 4) : "memory");
                    }
                }
            };
            float acc[HALF];
#pragma unroll
            for (int j = 0; j < HALF; ++j) acc[j] = 0.f;
            // acc += shift (per channel; the weights carry the folded-BN scale): needs the params fetched after the first
            // K-block, so it rides on the second K-block's drain
            auto add_shift = [&]() {
                mbar_wait(pbar, tile_par);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < HALF; j += 4) {
                    const float4 sh = lds_f4(par + 4 * j);
                    acc[j] += sh.x; acc[j + 1] += sh.y; acc[j + 2] += sh.z; acc[j + 3] += sh.w;
                }
            };
            if (warp == 6 && lane == 0) TL(g, 13);                            // tile prologue (params, residual prefetch) issued
            for (int kb = 0; kb < kblocks; ++kb, ++g) {
                if (p.ablate & 16) mbar_wait<true>(&d_full[buf], dpar);
                else mbar_wait(&d_full[buf], dpar);
                if (warp == 6 && lane == 0) TL(g, 7);
                tc_fence_after();
                // software-pipelined drain in 16-column pieces: the load of piece i+1 is in flight while piece i is added;
                // the buffer is handed back to the MMA issuer as soon as the last load has landed (before the last adds)
                {
                    const uint32_t src = tmem_base + buf * BN + lane_col;
                    uint32_t va[16], vb[16];
                    uint32_t scale_bits = 0x3F800000u;
                    if constexpr (MODE) tmem_ld1_issue(tmem_base + C::TMEM_SC0 + (g & (C::SCALE_SLOTS - 1)) + ((uint32_t)(quarter * 32) << 16), scale_bits);
                    tmem_ld16_issue(src, va);
                    tmem_ld_wait();
                    // MODE 1: partial * 2^-k (exact) added with one rounding - the fma is the TF32 mode's add
                    const float inv = __uint_as_float(scale_bits);
                    auto accum = [&](float& a, uint32_t v) {
                        if constexpr (MODE) a = fmaf(__uint_as_float(v), inv, a); else a += __uint_as_float(v);
                    };
#pragma unroll
                    for (int pc = 0; pc < HALF / 16; pc += 2) {
                        if (pc + 1 < HALF / 16) tmem_ld16_issue(src + (pc + 1) * 16, vb);
#pragma unroll
                        for (int j = 0; j < 16; ++j) accum(acc[pc * 16 + j], va[j]);                 // round-to-nearest fp32 running sum
                        if (pc + 1 < HALF / 16) {
                            tmem_ld_wait();
                            if (pc + 2 < HALF / 16) tmem_ld16_issue(src + (pc + 2) * 16, va);
                            else {
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) { if constexpr (PAIR) mbar_arrive_leader(&d_empty[buf]); else mbar_arrive(&d_empty[buf]); }
                            }
#pragma unroll
                            for (int j = 0; j < 16; ++j) accum(acc[(pc + 1) * 16 + j], vb[j]);
                            if (pc + 2 < HALF / 16) tmem_ld_wait();
                        }
                    }
                }
                if (warp == 6 && lane == 0) TL(g, 8);
                if (++buf == (uint32_t)C::ISSUERS) { buf = 0; dpar ^= 1; }
                if (kb == 0) after_first_kblock();
                else if (kb == 1) add_shift();                                // off the tile-boundary critical path
                if (kb == slab_kb && (has_res || (p.ablate & 128))) slab_ready();
            }
            if (kblocks == 1) add_shift();
            if (!has_res && !(p.ablate & 128)) slab_ready();
            // ---- fused epilogue (pixel per thread == TMEM lane, in place in the slab):
            //      y = post_scale * act(acc [+ res1]) [+ res2]      (acc already = conv*scale + shift)
            asm volatile("cp.async.wait_all;" ::: "memory");                  // params (+ this warp's rows of a cp.async residual)
            if (has_res && p.res_tma) mbar_wait(rbar, tile_par);
            __syncwarp();
            if (warp == 6 && lane == 0) TL(g - 1, 10);
            // two 4-channel groups per step: their six shared-memory loads are in flight together (in-order issue, only two
            // epilogue warps per scheduler to hide latency).  SIMPLE = the ResNet/FPN/SSH form max(acc*scale+shift [+res], lo):
            // half the instructions of the general form (RRDB: leaky, post-scale, post-activation residual).
            auto phase1 = [&](auto simple_tag) {
                constexpr bool SIMPLE = decltype(simple_tag)::value;
                const float lo_clamp = neg_slope == 0.f ? 0.f : -INFINITY;    // relu | none
                auto fuse = [&](float x, const float rv, const float ns) -> float {   // x = conv*scale + shift already
                    if constexpr (SIMPLE) {
                        if (has_res) x += rv;
                        return fmaxf(x, lo_clamp);
                    } else {
                        if (r_pre) x += rv;
                        x = fmaxf(x, x * ns) * post_scale;                    // act in {none, relu, leaky}: 0 <= ns <= 1
                        if (r_post) x += rv;
                        return x;
                    }
                };
                // software-pipelined by hand (the shared-memory accesses are volatile asm, the compiler keeps their order):
                // the residual loads of step i+1 are issued before the arithmetic of step i
                struct Step { float4 r0, r1; };
                auto load_step = [&](int j) -> Step {
                    Step t;
                    t.r0 = make_float4(0.f, 0.f, 0.f, 0.f); t.r1 = t.r0;
                    if (has_res) {
                        t.r0 = lds_f4(slab_addr(j >> 5, lane, (j >> 2) & 7));
                        t.r1 = lds_f4(slab_addr(j >> 5, lane, ((j >> 2) & 7) + 1));
                    }
                    return t;
                };
                Step cur = load_step(0);
#pragma unroll
                for (int j = 0; j < HALF; j += 8) {
                    const uint32_t a0 = slab_addr(j >> 5, lane, (j >> 2) & 7), a1 = slab_addr(j >> 5, lane, ((j >> 2) & 7) + 1);
                    Step nxt = cur;
                    if (j + 8 < HALF) nxt = load_step(j + 8);
                    float4 x0, x1;
                    const float ns = n0 + j < p.act_cols ? neg_slope : 1.f;   // channels past act_cols: no activation
                    x0.x = fuse(acc[j], cur.r0.x, ns); x0.y = fuse(acc[j + 1], cur.r0.y, ns);
                    x0.z = fuse(acc[j + 2], cur.r0.z, ns); x0.w = fuse(acc[j + 3], cur.r0.w, ns);
                    x1.x = fuse(acc[j + 4], cur.r1.x, ns); x1.y = fuse(acc[j + 5], cur.r1.y, ns);
                    x1.z = fuse(acc[j + 6], cur.r1.z, ns); x1.w = fuse(acc[j + 7], cur.r1.w, ns);
                    sts_f4(a0, x0);
                    sts_f4(a1, x1);
                    cur = nxt;
                }
            };
            if (post_scale == 1.f && !r_post && (neg_slope == 0.f || neg_slope == 1.f) && p.act_cols >= p.Cout) phase1(std::true_type{});
            else phase1(std::false_type{});
            if (warp == 6 && lane == 0) TL(g - 1, 11);
            if (p.out_tma) {
                // ---- bulk tensor store of the slab (rows / channels outside the tensor are clipped by TMA)
                fence_async_smem();
                group_sync();                                                 // all 128 pixel rows of the group's chunks are final
                if (warp == 6 && lane == 0) TL(g - 1, 12);
                if (dma && !(p.ablate & 2)) {
                    if (n0 + quarter * 32 < p.Cout) {
                        if (p.out_add) tma_reduce_add_4d(&p.tmOut, S + quarter * C::CHUNK_BYTES, n0 + quarter * 32, wo0, ho0, img);
                        else tma_store_4d(&p.tmOut, S + quarter * C::CHUNK_BYTES, n0 + quarter * 32, wo0, ho0, img);
                    }
                    bulk_commit();
                }
            } else {
                // ---- per-thread fallback: slab read back channel-contiguous (8 lanes cover the 128 bytes of one pixel and
                //      chunk), + the RRDB second residual; scalar stores where Cout is not a multiple of 4 (19-class logits)
                __syncwarp();                                                 // (only this warp's own rows are read back)
                const int sub = lane & 7, rbase = lane >> 3;
                const bool has3 = p.res3 != nullptr;
#pragma unroll 1
                for (int q = 0; q < CHUNKS; ++q) {
                    const int n = n0 + q * 32 + sub * 4;
                    if (n >= p.Cout) continue;
                    const bool vec = n + 3 < p.Cout && ((p.out_cs | p.out_co) & 3) == 0 && (!has3 || ((p.res3_cs | p.res3_co) & 3) == 0);
#pragma unroll 1
                    for (int st = 0; st < 8; ++st) {
                        const int r = rbase + 4 * st, m = out_pixel(r);
                        if (m < 0) continue;
                        float4 x = lds_f4(slab_addr(q, r, sub));
                        float* const op = p.out + (size_t)m * p.out_cs + p.out_co + n;
                        const float* const r3 = has3 ? p.res3 + (size_t)m * p.res3_cs + p.res3_co + n : nullptr;
                        if (vec) {
                            if (has3) {
                                const float4 t = __ldg(reinterpret_cast<const float4*>(r3));
                                x.x = x.x * p.post_scale2 + t.x; x.y = x.y * p.post_scale2 + t.y;
                                x.z = x.z * p.post_scale2 + t.z; x.w = x.w * p.post_scale2 + t.w;
                            }
                            *reinterpret_cast<float4*>(op) = x;
                        } else {
                            const float xe[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (n + e < p.Cout) op[e] = has3 ? xe[e] * p.post_scale2 + r3[e] : xe[e];
                        }
                    }
                }
                group_sync();                                                 // slab + params are rewritten by the next tile
            }
            if (warp == 6 && lane == 0) TL(g - 1, 9);                        // epilogue of this tile finished
        }
        if (dma) bulk_wait_all();                                       // global writes of the last tile complete before exit
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) {
        cluster_sync_all();                                              // no remote signal may target a CTA that has exited
        if (warp == 1) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    } else if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
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

bool make_map(CUtensorMap* map, void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box,
              CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    return fn(map, dtype, rank, base, dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int MODE>
int launch_pair(fcp_ctx* ctx, const TcParams& p) {
    using C = Cfg<BN, MODE>;
    static uint64_t configured = 0;
    if (!((configured >> (ctx->device & 63)) & 1)) {
        FCP_CUDA(ctx, (cudaFuncSetAttribute(conv_tc_kernel<BN, MODE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES)));
        configured |= (uint64_t)1 << (ctx->device & 63);
    }
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    // persistent kernel: exactly as many pairs as can be resident at once (GPCs with an odd SM count leave an SM unpaired)
    static int resident[64] = {0};
    if (!resident[ctx->device & 63]) {
        cfg.gridDim = dim3(2 * (ctx->sm_count / 2));
        int n = 0;
        FCP_CUDA(ctx, (cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<BN, MODE, 1>, &cfg)));
        resident[ctx->device & 63] = std::max(1, std::min(n, ctx->sm_count / 2));
        if (getenv("FCP_LOG_CONV")) fprintf(stderr, "[conv_tc] %d CTA pairs resident on %d SMs\n", n, ctx->sm_count);
    }
    const int pairs = std::min(p.pair_tiles, resident[ctx->device & 63]);
    cfg.gridDim = dim3(2 * pairs);
    FCP_CUDA(ctx, (cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, MODE, 1>, p)));
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

template <int BN, int MODE>
int launch(fcp_ctx* ctx, const TcParams& p) {
    using C = Cfg<BN, MODE>;
    static uint64_t configured = 0;                    // one bit per device: the attribute is per device
    if (!((configured >> (ctx->device & 63)) & 1)) {
        FCP_CUDA(ctx, (cudaFuncSetAttribute(conv_tc_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES)));
        configured |= (uint64_t)1 << (ctx->device & 63);
    }
    int grid = p.num_tiles < ctx->sm_count ? p.num_tiles : ctx->sm_count;
    conv_tc_kernel<BN, MODE><<<grid, NUM_THREADS, C::SMEM_BYTES, ctx->stream>>>(p);
    FCP_KERNEL_CHECK(ctx);
    return FCP_OK;
}

}  // namespace

// the uint8 stem route needs TMA-addressable image rows: 16-byte aligned base and row pitch
bool conv_tc_stem_supported(const void* images, int h, int w) {
    return encode_fn() != nullptr && (reinterpret_cast<uintptr_t>(images) & 15) == 0 && (w * 3) % 16 == 0 && h >= 8 && w >= 16 && h % 2 == 0 && w % 2 == 0;
}

bool conv_tc_supported(const ConvOp& op) {
    const ConvWeights& wt = *op.wt;
    if (op.stem_src) return op.impl >= 2 && wt.h_hi && wt.cout_pad == 64 && conv_tc_stem_supported(op.stem_src, op.stem_h, op.stem_w);
    if (wt.cin % KB != 0 || op.up_in) return false;
    if (op.in2.p && (wt.k != 1 || wt.kh || op.stride != 1 || op.stride_w > 1 || op.in.c % KB != 0 || op.in.c + op.in2.c != wt.cin ||
                     ((op.in2.cs | op.in2.co) & 3) || (op.in2_stride != 1 && op.in2_stride != 2)))
        return false;
    const int stride_w = op.stride_w ? op.stride_w : op.stride;
    if ((op.stride != 1 && op.stride != 2) || (stride_w != 1 && stride_w != 2)) return false;
    if (op.stride == 2 && !(wt.k == 1 || wt.k == 3 || wt.kh)) return false;
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
    const bool stem = op.stem_src != nullptr;
    if (!stem && (op.in.c + (op.in2.p ? op.in2.c : 0) != wt.cin || op.out.c != wt.cout)) return fail(ctx, FCP_ERR_INVALID, "conv: channel mismatch");
    TcParams p{};
    const int H = stem ? op.stem_h : op.in.h, W = stem ? op.stem_w : op.in.w, cs = op.in.cs;
    p.N = op.out.n; p.Ho = op.out.h; p.Wo = op.out.w; p.Cout = wt.cout; p.Cin = wt.cin;
    p.KH = wt.kh ? wt.kh : wt.k; p.KW = wt.kw ? wt.kw : wt.k;
    p.stride = op.stride; p.pad = op.pad;
    p.stride_w = op.stride_w ? op.stride_w : op.stride; p.pad_w = op.pad_w >= 0 ? op.pad_w : op.pad;
    if (stem) {
        // K-block b of the packed weights = tap rows 2b, 2b+1 of the 7x7 kernel (4 K-blocks of 64): the kernel sees a 4x1 "conv"
        if ((H + 6 - 7) / 2 + 1 != p.Ho || (W + 6 - 7) / 2 + 1 != p.Wo || op.out.c != 64) return fail(ctx, FCP_ERR_INVALID, "stem: output shape mismatch");
        p.KH = 4; p.KW = 1; p.Cin = 64; p.stride = p.stride_w = 1; p.pad = p.pad_w = 0;
        p.stem = 1; p.stem_H = H; p.stem_W = W;
    } else if (op.out_rs) {
        // parity class of an upsampling conv: 2x2 taps, `pad` rows/columns of zero padding BEFORE and 1 - pad after -> same size
        if (p.KH != 2 || p.KW != 2 || p.stride != 1 || p.stride_w != 1 || p.Ho != H || p.Wo != W || (unsigned)p.pad > 1u || (unsigned)p.pad_w > 1u)
            return fail(ctx, FCP_ERR_INVALID, "conv: a strided output view takes a 2x2 stride-1 same-size conv");
    } else if ((H + 2 * p.pad - p.KH) / p.stride + 1 != p.Ho || (W + 2 * p.pad_w - p.KW) / p.stride_w + 1 != p.Wo)
        return fail(ctx, FCP_ERR_INVALID, "conv: output shape mismatch");
    // spatial box of 128 output pixels: widest power-of-two width that does not overshoot the row by more than 2x
    int bw_log2 = 7;
    while (bw_log2 > 0 && (1 << bw_log2) >= 2 * p.Wo) --bw_log2;
    if (bw_log2 < 3) bw_log2 = 3;
    int best = -1; size_t best_tiles = 0;
    for (int l = 3; l <= 7; ++l) {                            // pick the box with the fewest tiles (ties: squarer)
        int bw = 1 << l, bh = TILE_M / bw;
        if (stem && l > 5) continue;                          // the uint8 halo box is 6*BW + 15 bytes wide (TMA: <= 256)
        size_t t = (size_t)((p.Wo + bw - 1) / bw) * ((p.Ho + bh - 1) / bh);
        if (best < 0 || t < best_tiles || (t == best_tiles && abs(l - 4) < abs(best - 4))) { best = l; best_tiles = t; }
    }
    bw_log2 = best;
    const int BW = 1 << bw_log2, BH = TILE_M / BW;
    p.bw_log2 = bw_log2; p.BH = BH;
    p.tiles_x = (p.Wo + BW - 1) / BW; p.tiles_y = (p.Ho + BH - 1) / BH;
    // 96 output channels: one padded 128-wide tile (the activation tile is read once) instead of three 32-wide ones
    const int BN = wt.cout_pad % 128 == 0 || wt.cout_pad == 96 ? 128 : (wt.cout_pad % 64 == 0 ? 64 : 32);
    p.tiles_n = (wt.cout_pad + BN - 1) / BN;
    p.num_tiles = p.N * p.tiles_x * p.tiles_y * p.tiles_n;
    // ---- tensor maps
    float* base = op.in.p + op.in.co;
    if (stem) {
        p.stem_box_w = (15 + 6 * BW + 15 + 15) / 16 * 16;        // up to 15 bytes of alignment slack in front of the 6*BW + 15 window bytes
        p.stem_box_h = 2 * BH + 5;
        cuuint64_t dims[3] = {(cuuint64_t)W * 3, (cuuint64_t)H, (cuuint64_t)p.N};
        cuuint64_t strides[2] = {(cuuint64_t)W * 3, (cuuint64_t)H * W * 3};
        cuuint32_t box[3] = {(cuuint32_t)p.stem_box_w, (cuuint32_t)p.stem_box_h, 1};
        if (p.stem_box_w > 256 || p.stem_box_h > 256 || p.stem_box_w * p.stem_box_h > A_TILE_BYTES ||
            !make_map(&p.tmStem, const_cast<uint8_t*>(op.stem_src), 3, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_UINT8, CU_TENSOR_MAP_SWIZZLE_NONE))
            return fail(ctx, FCP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the uint8 image");
    } else
    for (int py = 0; py < p.stride; ++py)
        for (int px = 0; px < p.stride_w; ++px) {
            const int sh = p.stride, sw = p.stride_w;
            cuuint64_t dims[4] = {(cuuint64_t)op.in.c, (cuuint64_t)((W - px + sw - 1) / sw), (cuuint64_t)((H - py + sh - 1) / sh), (cuuint64_t)p.N};
            cuuint64_t strides[3] = {(cuuint64_t)sw * cs * 4, (cuuint64_t)sh * W * cs * 4, (cuuint64_t)H * W * cs * 4};
            cuuint32_t box[4] = {KB, (cuuint32_t)BW, (cuuint32_t)BH, 1};
            if (dims[1] == 0 || dims[2] == 0) { dims[1] = dims[1] ? dims[1] : 1; dims[2] = dims[2] ? dims[2] : 1; }
            if (!make_map(&p.tmA[py * 2 + px], base + ((size_t)py * W + px) * cs, 4, dims, strides, box))
                return fail(ctx, FCP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the activation tensor");
        }
    if (!stem && op.in2.p) {   // second K source: the (0, 0) sampling of in2 at the output resolution
        const int s2 = op.in2_stride, H2 = op.in2.h, W2 = op.in2.w, cs2 = op.in2.cs;
        if ((H2 + s2 - 1) / s2 != p.Ho || (W2 + s2 - 1) / s2 != p.Wo || op.in2.n != p.N) return fail(ctx, FCP_ERR_INVALID, "conv_tc: second source does not match the output grid");
        cuuint64_t dims[4] = {(cuuint64_t)op.in2.c, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)p.N};
        cuuint64_t strides[3] = {(cuuint64_t)s2 * cs2 * 4, (cuuint64_t)s2 * W2 * cs2 * 4, (cuuint64_t)H2 * W2 * cs2 * 4};
        cuuint32_t box[4] = {KB, (cuuint32_t)BW, (cuuint32_t)BH, 1};
        if (!make_map(&p.tmA[1], op.in2.p + op.in2.co, 4, dims, strides, box))
            return fail(ctx, FCP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the second activation tensor");
        p.src2_units = op.in.c / 32;
    }
    const bool f16 = op.impl >= 2;
    if (f16 && !wt.h_hi) return fail(ctx, FCP_ERR_INVALID, "conv_tc: this convolution has no fp16 packing");
    p.cin_p = f16 ? wt.cin_p : wt.cin;
    p.w_exp = wt.w_exp;
    p.a_exact = op.a_exact || stem;
    p.single = op.impl == 3;
    p.act_cols = op.act_cols;
    p.out_add = op.out_add;
    const cuuint64_t K = f16 ? ((cuuint64_t)p.KH * p.KW * p.cin_p + 63) / 64 * 64 : (cuuint64_t)p.KH * p.KW * p.cin_p;
    cuuint64_t bdims[2] = {K, (cuuint64_t)wt.cout_pad};
    cuuint64_t bstr[1] = {K * (f16 ? 2 : 4)};
    // CTA pairs (cta_group::2): opt-in, measured SLOWER than one CTA per tile - FCP_TC_PAIR=1 takes the wide
    // f16 tiles of launches that give every pair work, FCP_TC_PAIR=2 every wide f16 tile (tests)
    const char* pair_s = getenv("FCP_TC_PAIR");
    const int pair_env = pair_s ? atoi(pair_s) : 0;
    const int m_tiles = p.N * p.tiles_x * p.tiles_y;
    p.pair_tiles = p.tiles_n * ((m_tiles + 1) / 2);
    const bool pair = pair_env && f16 && !stem && BN == 128 && (p.pair_tiles >= ctx->sm_count / 2 || pair_env > 1);
    cuuint32_t bbox[2] = {(cuuint32_t)(f16 ? 64 : 32), (cuuint32_t)(pair ? BN / 2 : BN)};
    const CUtensorMapDataType bt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    if (!make_map(&p.tmBhi, f16 ? wt.h_hi : (void*)wt.w_hi, 2, bdims, bstr, bbox, bt) ||
        !make_map(&p.tmBlo, f16 ? wt.h_lo : (void*)wt.w_lo, 2, bdims, bstr, bbox, bt))
        return fail(ctx, FCP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the weights");
    // ---- epilogue tensor maps: one box = the tile's 128 pixels x 32 channels (one chunk of an epilogue group)
    {
        const int bws = BW, bhs = BH;
        auto tensor_ok = [](const float* base, int cs) { return cs % 4 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0; };
        auto epi_map = [&](CUtensorMap* map, const float* base, int cs, size_t rs = 0, size_t is = 0) {
            cuuint64_t dims[4] = {(cuuint64_t)wt.cout, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)p.N};
            cuuint64_t strides[3] = {(cuuint64_t)cs * 4, (cuuint64_t)(rs ? rs : (size_t)p.Wo * cs) * 4,
                                     (cuuint64_t)(is ? is : (size_t)p.Ho * p.Wo * cs) * 4};
            cuuint32_t box[4] = {32, (cuuint32_t)bws, (cuuint32_t)bhs, 1};
            return make_map(map, const_cast<float*>(base), 4, dims, strides, box);
        };
        p.out_tma = !op.res3 && tensor_ok(op.out.p + op.out.co, op.out.cs) && getenv("FCP_TC_NO_TMA_EPI") == nullptr;
        if (p.out_tma && !epi_map(&p.tmOut, op.out.p + op.out.co, op.out.cs, (size_t)op.out_rs, op.out_is))
            return fail(ctx, FCP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the output tensor");
        if ((op.out_add || op.out_rs || op.out_is) && !p.out_tma)
            return fail(ctx, FCP_ERR_INVALID, "conv_tc: out_add / a strided output view need the TMA epilogue");
        const float* rsrc = op.res1 ? op.res1 : op.res2;
        const int r_cs = op.res1 ? op.res1_cs : op.res2_cs, r_co = op.res1 ? op.res1_co : op.res2_co;
        p.res_tma = rsrc && !(op.res2 && op.res2_h) && tensor_ok(rsrc + r_co, r_cs) && getenv("FCP_TC_NO_TMA_EPI") == nullptr;
        if (p.res_tma && !epi_map(&p.tmRes, rsrc + r_co, r_cs))
            return fail(ctx, FCP_ERR_CUDA, "cuTensorMapEncodeTiled failed for the residual tensor");
    }
    p.out = op.out.p; p.out_cs = op.out.cs; p.out_co = op.out.co;
    p.scale = wt.scale; p.shift = wt.shift;
    p.res1 = op.res1; p.res1_cs = op.res1_cs; p.res1_co = op.res1_co;
    p.act = op.act; p.slope = op.slope;
    p.post_scale = op.post_scale; p.res2 = op.res2; p.res2_cs = op.res2_cs; p.res2_co = op.res2_co; p.res2_h = op.res2_h; p.res2_w = op.res2_w;
    p.post_scale2 = op.post_scale2; p.res3 = op.res3; p.res3_cs = op.res3_cs; p.res3_co = op.res3_co;
    static const int ablate = getenv("FCP_TC_ABLATE") ? atoi(getenv("FCP_TC_ABLATE")) : 0;
    p.ablate = ablate;
    auto do_launch = [&]() -> int {
        if (f16) {
            if (pair) return launch_pair<128, 1>(ctx, p);
            if (BN == 128) return launch<128, 1>(ctx, p);
            if (BN == 64) return launch<64, 1>(ctx, p);
            return launch<32, 1>(ctx, p);
        }
        if (BN == 128) return launch<128, 0>(ctx, p);
        if (BN == 64) return launch<64, 0>(ctx, p);
        return launch<32, 0>(ctx, p);
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
                p.KH * p.KW * ((wt.cin + (f16 ? 63 : 31)) / (f16 ? 64 : 32)), p.num_tiles);
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
